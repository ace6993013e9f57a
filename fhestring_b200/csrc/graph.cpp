// graph.cpp -- see graph.h.  Recipes follow SURVEY.md 2.4/2.5 (tfhe-rs 0.5 radix ops over 2_2 blocks) in
// VALUE: every op returns a clean char (blocks in [0,3]; flags: block 0 in {0,1}, blocks 1..3 trivial 0)
// holding exactly the u8 the reference's op holds.  Where the reference's recipe does work that cannot
// change the plaintext (PBS on trivial blocks, message_extract of an already clean sum, the full 8-PBS
// subtract behind flip) the value-set analysis folds it away.
#include "graph.h"

#include <algorithm>
#include <cstring>

namespace fhestr {

static inline int mod32(int v) { return ((v % 32) + 32) % 32; }
static inline int popcount32(uint32_t v) { return __builtin_popcount(v); }

int Graph::lut_id(const std::array<uint8_t, 16>& t) {
    auto it = lut_ids.find(t);
    if (it != lut_ids.end()) return it->second;
    const int id = (int)lut_tables.size();
    lut_tables.push_back(t);
    lut_ids[t] = id;
    return id;
}

BlockId Graph::trivial_block(int value) {
    value = mod32(value);
    if (!triv_cache_init) {
        nodes.reserve(1 << 16);
        cse_tab.assign(1 << 14, ~0ull);
        for (int v = 0; v < 32; v++) {
            BlockNode n;
            n.kind = BKind::Trivial;
            n.vset = 1u << v;
            n.cst = v;
            triv_cache[v] = (BlockId)nodes.size();
            nodes.push_back(n);
        }
        triv_cache_init = true;
    }
    return triv_cache[value];
}

int Graph::trivial_value(BlockId b) const { return nodes[b].cst; }

BlockId Graph::input_block() {
    BlockNode n;
    n.kind = BKind::Input;
    n.vset = 0xF;  // clean message block
    n.noise2 = 1.f;
    n.slot = (int32_t)next_slot++;  // callers upload the ciphertext to this arena slot before running
    n.done = true;
    n_input_nodes++;
    nodes.push_back(n);
    return (BlockId)nodes.size() - 1;
}

static inline uint32_t rotl32(uint32_t x, int k) { k &= 31; return k ? (x << k) | (x >> (32 - k)) : x; }

// set of values sum_i coeff_i * v_i + cst (mod 32) over v_i in vset_i: adding coeff * v shifts the whole partial set
// by coeff * v, i.e. rotates its 32-bit mask
uint32_t Graph::vset_of(const std::vector<std::pair<BlockId, int>>& ops, int cst) const {
    uint32_t S = 1u << mod32(cst);
    for (auto& op : ops) {
        uint32_t V = nodes[op.first].vset, T = 0;
        while (V) {
            const int v = __builtin_ctz(V);
            V &= V - 1;
            const int sv = v >= 16 ? v - 32 : v;  // bit 4 set = negative value (padding bit)
            T |= rotl32(S, mod32(op.second * sv));
        }
        S = T;
        if (S == 0xFFFFFFFFu) break;
    }
    return S;
}

void Graph::flatten(const std::vector<std::pair<BlockId, int>>& ops, int cst, std::vector<Term>& terms, int& out_cst) {
    for (;;) {
        std::vector<std::pair<BlockId, int>>& acc = flat_scratch;
        acc.clear();
        auto add = [&](BlockId b, int k) {
            for (auto& kv : acc) if (kv.first == b) { kv.second += k; return; }
            acc.push_back({b, k});
        };
        int c = cst;
        BlockId widest = 0;
        size_t widest_n = 0;
        for (auto& op : ops) {
            const BlockNode& n = nodes[op.first];
            if (n.kind == BKind::Trivial) {
                const int v = n.cst >= 16 ? n.cst - 32 : n.cst;
                c += op.second * v;
            } else if (n.kind == BKind::Linear && !n.materialized) {
                for (auto& t : n.terms) add(t.blk, op.second * t.coeff);
                c += op.second * n.cst;
                if (n.terms.size() > widest_n) { widest_n = n.terms.size(); widest = op.first; }
            } else {
                add(op.first, op.second);
            }
        }
        std::sort(acc.begin(), acc.end());     // the CSE key is order-sensitive: canonical order by block id
        terms.clear();
        for (auto& kv : acc) if (kv.second != 0) terms.push_back(Term{kv.first, kv.second});
        out_cst = mod32(c);
        if (terms.size() <= FHESTR_MAX_TERMS) return;
        if (widest_n <= 1) { fail("linear combination needs more than FHESTR_MAX_TERMS atoms"); terms.resize(FHESTR_MAX_TERMS); return; }
        nodes[widest].materialized = true;  // becomes an atom with its own leveled job
    }
}

float Graph::noise2_of(const std::vector<Term>& terms) const {
    float s = 0;
    for (auto& t : terms) s += (float)t.coeff * (float)t.coeff * nodes[t.blk].noise2;
    return s;
}

// a materialised Linear atom is written by the leveled section of its level, so a Linear consumer
// (same section) has to wait one level; PBS consumers are one level later anyway
int Graph::level_of(const std::vector<Term>& terms, bool for_linear) const {
    int l = 0;
    for (auto& t : terms) {
        const BlockNode& n = nodes[t.blk];
        l = std::max(l, (int)n.level + ((for_linear && n.kind == BKind::Linear) ? 1 : 0));
    }
    return l;
}

BlockId Graph::lin(const std::vector<std::pair<BlockId, int>>& ops, int cst, uint32_t declared_vset) {
    std::vector<Term> terms;
    int c;
    flatten(ops, cst, terms, c);
    if (terms.empty()) return trivial_block(c);
    if (terms.size() == 1 && terms[0].coeff == 1 && c == 0) return terms[0].blk;
    BlockNode n;
    n.kind = BKind::Linear;
    n.vset = declared_vset ? declared_vset : vset_of(ops, cst);
    n.terms = terms;
    n.cst = c;
    n.noise2 = noise2_of(terms);
    n.level = level_of(terms, true);
    nodes.push_back(n);
    return (BlockId)nodes.size() - 1;
}

static std::array<uint8_t, 16> table_of(const std::function<int(int)>& f) {
    std::array<uint8_t, 16> t;
    for (int v = 0; v < 16; v++) t[v] = (uint8_t)(f(v) & 15);
    return t;
}

BlockId Graph::refresh(BlockId b) {
    return pbs({{b, 1}}, 0, table_of([](int v) { return v; }));
}

// scope in which PBS inputs may carry the padding bit (negative values): the comparison recipe
struct SignedScope {
    bool& flag;
    bool prev;
    explicit SignedScope(bool& f) : flag(f), prev(f) { flag = true; }
    ~SignedScope() { flag = prev; }
};

BlockId Graph::pbs(const std::vector<std::pair<BlockId, int>>& ops_in, int cst, const std::array<uint8_t, 16>& table) {
    std::vector<std::pair<BlockId, int>> ops = ops_in;
    std::vector<Term> terms;
    int c;
    flatten(ops, cst, terms, c);
    // value set of the input: the operands as given (lazy sums carry the range their recipe declared) AND the merged
    // atoms (x - x vanishes, a block that occurs twice counts once with coefficient 2): both over-approximate the
    // true set, so their intersection does too
    auto atom_ops = [&]() {
        std::vector<std::pair<BlockId, int>> a;
        for (auto& t : terms) a.push_back({t.blk, t.coeff});
        return a;
    };
    uint32_t S = vset_of(ops, cst) & vset_of(atom_ops(), c);
    if (!S) S = vset_of(atom_ops(), c);
    // half-step tables (entries 0x80 | e, fhestr_lut_register): f(v) = e[v] below 16, 1 - e[v - 16] from 16 on; they
    // are MEANT to be read on [16, 32) (threshold of a sum), so the padding-bit checks do not apply to them
    const bool half = (table[0] & 0x80) != 0;
    if (!half && (S >> 16 & 1)) fail("PBS input can reach the ambiguous value 16");
    if (!half && !signed_ok && (S >> 16)) {
        std::string m = "PBS input overflows into the padding bit:";
        for (auto& op : ops) m += " (" + std::to_string(op.second) + " x vset " + std::to_string(nodes[op.first].vset) + ")";
        fail(m + " + " + std::to_string(cst));
    }
    // image of the value set; padding-bit inputs give -f(v-16)
    uint32_t O = 0;
    for (int v = 0; v < 32; v++) {
        if (!(S >> v & 1)) continue;
        if (half) O |= 1u << (v < 16 ? (table[v] & 0x7f) : mod32(1 - (int)(table[v - 16] & 0x7f)));
        else O |= 1u << (v < 16 ? table[v] : mod32(-(int)table[v - 16]));
    }
    if (popcount32(O) == 1) return trivial_block(__builtin_ctz(O));

    // keep the input inside the variance the reference's own recipes use (graph.h).  First refresh the noisiest
    // lazily-summed operand; when none qualifies, the excess sits in the COEFFICIENTS of the merged atoms (the same
    // block summed m times -- identical windows over trivial chars, a string compared with itself, repeat_clear --
    // counts m^2): such a term is replaced by ONE PBS that computes m * v from the block alone (input noise 1,
    // output noise 1).  A refresh could not do that: the identity PBS of a block is one node however often it is asked for.
    for (int guard = 0; noise2_of(terms) > kNoise2Limit && guard < 24; guard++) {
        int worst = -1;
        float worst_c = 0;
        for (size_t i = 0; i < ops.size(); i++) {
            const BlockNode& n = nodes[ops[i].first];
            if (n.kind != BKind::Linear || n.noise2 <= 1.f || n.noise2 > kNoise2Limit || (n.vset >> 16)) continue;
            const float contrib = (float)ops[i].second * ops[i].second * n.noise2;
            if (contrib > worst_c) { worst_c = contrib; worst = (int)i; }
        }
        if (worst >= 0) {
            ops[worst].first = refresh(ops[worst].first);
            flatten(ops, cst, terms, c);
            continue;
        }
        int wt = -1;
        worst_c = 0;
        for (size_t i = 0; i < terms.size(); i++) {
            const BlockNode& n = nodes[terms[i].blk];
            const int m = std::abs(terms[i].coeff);
            if (m < 2 || (n.vset >> 16) || n.noise2 > kNoise2Limit) continue;
            bool fits = true;
            for (int v = 0; v < 16; v++) if ((n.vset >> v & 1) && m * v > 15) fits = false;
            const float contrib = (float)m * m * n.noise2;
            if (fits && contrib > worst_c) { worst_c = contrib; wt = (int)i; }
        }
        if (wt < 0) {
            std::string m = "noise budget exceeded and nothing to refresh:";
            for (auto& t : terms) {
                const BlockNode& n = nodes[t.blk];
                m += " (" + std::to_string(t.coeff) + " x kind " + std::to_string((int)n.kind) + " noise2 " + std::to_string(n.noise2) + ")";
            }
            fail(m);
            break;
        }
        const int m = std::abs(terms[wt].coeff);
        const BlockId scaled = pbs({{terms[wt].blk, 1}}, 0, table_of([m](int v) { return m * v; }));
        ops = atom_ops();
        ops[wt] = {scaled, terms[wt].coeff < 0 ? -1 : 1};
        cst = c;
        flatten(ops, cst, terms, c);
    }
    if (terms.empty()) return trivial_block(half ? (c < 16 ? (table[c] & 0x7f) : 1 - (int)(table[c - 16] & 0x7f)) : table[c & 15] * (c < 16 ? 1 : -1));

    const int lid = lut_id(table);
    const uint64_t key = cse_hash(lid, c, terms);
    const BlockId hit = cse_find(key, lid, c, terms);
    if (hit != kNoBlock) return hit;

    BlockNode n;
    n.kind = BKind::Pbs;
    n.vset = O;
    n.noise2 = 1.f;
    n.lut = lid;
    n.terms = terms;
    n.cst = c;
    n.level = 1 + level_of(terms, false);
    nodes.push_back(n);
    n_pbs_nodes++;
    const BlockId id = (BlockId)nodes.size() - 1;
    cse_insert(key, id);
    return id;
}

uint64_t Graph::cse_hash(int lut, int cst, const std::vector<Term>& terms) {
    uint64_t h = 0x9e3779b97f4a7c15ull ^ ((uint64_t)(uint32_t)lut << 32 | (uint32_t)cst);
    for (const Term& t : terms) {
        h ^= ((uint64_t)t.blk << 32) | (uint32_t)t.coeff;
        h *= 0xff51afd7ed558ccdull;
        h ^= h >> 29;
    }
    h *= 0xc4ceb9fe1a85ec53ull;
    return h ^ (h >> 32);
}

static constexpr uint64_t kCseEmpty = ~0ull;

BlockId Graph::cse_find(uint64_t h, int lut, int cst, const std::vector<Term>& terms) const {
    if (cse_tab.empty()) return kNoBlock;
    const size_t mask = cse_tab.size() - 1;
    const uint64_t tag = h & 0xffffffffull;
    for (size_t i = (size_t)h & mask;; i = (i + 1) & mask) {
        const uint64_t e = cse_tab[i];
        if (e == kCseEmpty) return kNoBlock;
        if ((e >> 32) != tag) continue;                 // the node is only touched when the stored hash half matches
        const BlockNode& n = nodes[(BlockId)e];
        if (n.lut == lut && n.cst == cst && n.terms.size() == terms.size() &&
            std::equal(terms.begin(), terms.end(), n.terms.begin(), [](const Term& a, const Term& b) { return a.blk == b.blk && a.coeff == b.coeff; }))
            return (BlockId)e;
    }
}

void Graph::cse_insert(uint64_t h, BlockId id) {
    if (cse_tab.empty()) cse_tab.assign(1 << 14, kCseEmpty);
    if (2 * (cse_count + 1) > cse_tab.size()) {   // keep the load below one half; an entry keeps the LOWER hash half,
        std::vector<uint64_t> old;                 // which is all a slot index needs: growing touches no node
        old.swap(cse_tab);
        cse_tab.assign(old.size() * 4, kCseEmpty);
        const size_t mask = cse_tab.size() - 1;
        for (uint64_t e : old) {
            if (e == kCseEmpty) continue;
            size_t i = (size_t)(e >> 32) & mask;
            while (cse_tab[i] != kCseEmpty) i = (i + 1) & mask;
            cse_tab[i] = e;
        }
    }
    const size_t mask = cse_tab.size() - 1;
    size_t i = (size_t)h & mask;
    while (cse_tab[i] != kCseEmpty) i = (i + 1) & mask;
    cse_tab[i] = (h << 32) | id;
    cse_count++;
}

// f(x, y) over clean 2-bit block values; one operand trivial -> univariate LUT on the other
BlockId Graph::bivar(BlockId x, BlockId y, const std::function<int(int, int)>& f) {
    const bool tx = is_trivial(x), ty = is_trivial(y);
    if (tx && ty) return trivial_block(f(trivial_value(x), trivial_value(y)));
    if (((tx ? 0u : nodes[x].vset) | (ty ? 0u : nodes[y].vset)) & ~0xFu) fail("bivariate operand is not a clean 2-bit block");
    if (tx) {
        const int cx = trivial_value(x);
        return pbs({{y, 1}}, 0, table_of([&](int v) { return f(cx, v & 3); }));
    }
    if (ty) {
        const int cy = trivial_value(y);
        return pbs({{x, 1}}, 0, table_of([&](int v) { return f(v & 3, cy); }));
    }
    return pbs({{x, 4}, {y, 1}}, 0, table_of([&](int v) { return f(v >> 2, v & 3); }));
}

// ------------------------------------------------------------------------------------ char level
Char Graph::input_char() { return Char{input_block(), input_block(), input_block(), input_block()}; }

Char Graph::trivial_char(uint8_t v) {
    return Char{trivial_block(v & 3), trivial_block((v >> 2) & 3), trivial_block((v >> 4) & 3), trivial_block((v >> 6) & 3)};
}

Char Graph::flag_char(BlockId b) { return Char{b, trivial_block(0), trivial_block(0), trivial_block(0)}; }

std::pair<BlockId, BlockId> Graph::differs_and_strict(const Char& a, const Char& b, bool greater) {
    BlockId p[2];
    {
        SignedScope sc(signed_ok);
        for (int j = 0; j < 2; j++)
            p[j] = pbs({{a[2 * j], 1}, {a[2 * j + 1], 4}, {b[2 * j], -1}, {b[2 * j + 1], -4}}, 0,
                       table_of([](int v) { return v != 0; }));
    }
    // 4 (hi + 1) + (lo + 1) with hi, lo in {0 smaller, 1 equal, 2 greater}; the high nibble decides unless it is equal
    auto order = [](int v) { const int hi = v >> 2, lo = v & 3; return (hi == 1) ? lo : hi; };
    const BlockId d = pbs({{p[1], 4}, {p[0], 1}}, 5, table_of([order](int v) { return (int)(order(v) != 1); }));
    const int want = greater ? 2 : 0;
    const BlockId s = pbs({{p[1], 4}, {p[0], 1}}, 5, table_of([order, want](int v) { return (int)(order(v) == want); }));
    return {d, s};
}

Char Graph::eq(const Char& a, const Char& b) {
    std::vector<std::pair<BlockId, int>> s;
    for (int i = 0; i < 4; i++) s.push_back({bivar(a[i], b[i], [](int x, int y) { return x == y; }), 1});
    return flag_char(pbs(s, 0, table_of([](int v) { return v == 4; })));
}

Char Graph::ne(const Char& a, const Char& b) {
    std::vector<std::pair<BlockId, int>> s;
    for (int i = 0; i < 4; i++) s.push_back({bivar(a[i], b[i], [](int x, int y) { return x != y; }), 1});
    return flag_char(pbs(s, 0, table_of([](int v) { return v != 0; })));
}

// op: 0 lt, 1 le, 2 gt, 3 ge.  Pairs of blocks are packed lo + 4 hi, subtracted raw (the difference
// lives in [-15, 15]: negative values carry the padding bit, so the nz LUT returns -1 for them),
// sign = nz(diff) + 1 in {0:<, 1:=, 2:>}; the two signs are merged by one more LUT.
Char Graph::cmp(const Char& a, const Char& b, int op) {
    BlockId p[2];
    {
        SignedScope sc(signed_ok);
        for (int j = 0; j < 2; j++)
            p[j] = pbs({{a[2 * j], 1}, {a[2 * j + 1], 4}, {b[2 * j], -1}, {b[2 * j + 1], -4}}, 0,
                       table_of([](int v) { return v != 0; }));
    }
    auto fin = table_of([op](int v) {
        const int hi = v >> 2, lo = v & 3;
        const int r = (hi == 1) ? lo : hi;
        switch (op) {
            case 0: return (int)(r == 0);
            case 1: return (int)(r != 2);
            case 2: return (int)(r == 2);
            default: return (int)(r != 0);
        }
    });
    return flag_char(pbs({{p[1], 4}, {p[0], 1}}, 5, fin));
}
Char Graph::lt(const Char& a, const Char& b) { return cmp(a, b, 0); }
Char Graph::le(const Char& a, const Char& b) { return cmp(a, b, 1); }
Char Graph::gt(const Char& a, const Char& b) { return cmp(a, b, 2); }
Char Graph::ge(const Char& a, const Char& b) { return cmp(a, b, 3); }

Char Graph::bitand_(const Char& a, const Char& b) {
    Char r;
    for (int i = 0; i < 4; i++) r[i] = bivar(a[i], b[i], [](int x, int y) { return x & y; });
    return r;
}
Char Graph::bitor_(const Char& a, const Char& b) {
    Char r;
    for (int i = 0; i < 4; i++) r[i] = bivar(a[i], b[i], [](int x, int y) { return x | y; });
    return r;
}

// sequential carry propagation over 4 blocks (the path tfhe-rs takes for <= 4 blocks, A.10); a block
// whose carry is the same for every reachable value needs no PBS at all
std::array<BlockId, 4> Graph::propagate(std::array<BlockId, 4> s) {
    std::array<BlockId, 4> m;
    BlockId carry = trivial_block(0);
    for (int i = 0; i < 4; i++) {
        const BlockId cur = lin({{s[i], 1}, {carry, 1}}, 0);
        const uint32_t S = nodes[cur].vset;
        if (S >> 16) { fail("carry propagation input overflows the block"); m[i] = cur; continue; }
        uint32_t carries = 0, msgs = 0;
        for (int v = 0; v < 16; v++) if (S >> v & 1) { carries |= 1u << (v >> 2); msgs |= 1u << (v & 3); }
        if (popcount32(carries) == 1) {
            const int cv = __builtin_ctz(carries);
            m[i] = lin({{cur, 1}}, -4 * cv, msgs);
            carry = trivial_block(cv);
        } else {
            m[i] = pbs({{cur, 1}}, 0, table_of([](int v) { return v & 3; }));
            carry = (i < 3) ? pbs({{cur, 1}}, 0, table_of([](int v) { return v >> 2; })) : trivial_block(0);
        }
    }
    return m;
}

Char Graph::add(const Char& a, const Char& b) {
    std::array<BlockId, 4> s;
    for (int i = 0; i < 4; i++) s[i] = lin({{a[i], 1}, {b[i], 1}}, 0);
    return propagate(s);
}

// a - b mod 256: blocks a_i + (4 - b_i) with the borrowed 4 taken back from the next block
Char Graph::sub(const Char& a, const Char& b) {
    std::array<BlockId, 4> s;
    for (int i = 0; i < 4; i++) s[i] = lin({{a[i], 1}, {b[i], -1}}, i == 0 ? 4 : 3);
    return propagate(s);
}

BlockId Graph::cond_bit(const Char& c) {
    std::vector<std::pair<BlockId, int>> live;
    for (int i = 0; i < 4; i++) {
        if (is_trivial(c[i])) { if (trivial_value(c[i]) != 0) return trivial_block(1); }
        else live.push_back({c[i], 1});
    }
    if (live.empty()) return trivial_block(0);
    if (live.size() == 1 && !(nodes[live[0].first].vset & ~0x3u)) return live[0].first;
    return pbs(live, 0, table_of([](int v) { return v != 0; }));
}

// scalar_ne(cond, 0) then a CMUX per block: keep_if(4 t + c) + zero_if(4 f + c)
Char Graph::if_then_else(const Char& c, const Char& t, const Char& f) {
    const BlockId cb = cond_bit(c);
    if (is_trivial(cb)) return trivial_value(cb) ? t : f;
    Char r;
    for (int i = 0; i < 4; i++) {
        if (t[i] == f[i]) { r[i] = t[i]; continue; }
        const BlockId k = bivar(t[i], cb, [](int x, int cnd) { return (cnd & 1) ? x : 0; });
        const BlockId z = bivar(f[i], cb, [](int x, int cnd) { return (cnd & 1) ? 0 : x; });
        r[i] = lin({{k, 1}, {z, 1}}, 0, (nodes[t[i]].vset | nodes[f[i]].vset) & 0xF);
    }
    return r;
}

Char Graph::is_whitespace(const Char& a) {
    Char r = eq(a, trivial_char(0x20));
    for (uint8_t w : {0x09, 0x0A, 0x0B, 0x0C, 0x0D}) r = bitor_(r, eq(a, trivial_char(w)));
    return r;
}
Char Graph::is_uppercase(const Char& a) { return bitand_(ge(a, trivial_char(0x41)), le(a, trivial_char(0x5A))); }
Char Graph::is_lowercase(const Char& a) { return bitand_(ge(a, trivial_char(0x61)), le(a, trivial_char(0x7A))); }
Char Graph::flip(const Char& a) { return sub(trivial_char(1), a); }

// ---- wide reductions (plaintext-identical to chains of bitand / bitor / add on 0/1 chars)
static const int kChunk = 15;
// AND / OR of up to 16 flags in ONE PBS: the half-step table with e = 0 everywhere is the threshold [x >= 16]
// (its negacyclic half reads 1 - e), so  AND_k = [sum + (16 - k) >= 16]  and  OR = [sum + 15 >= 16].  A 16-entry
// table of whole steps cannot do either for 16 flags (17 different sums), which cost one more dependency level.
static const int kFan = 16;
static std::array<uint8_t, 16> threshold_table() {
    std::array<uint8_t, 16> t;
    t.fill(0x80);
    return t;
}
static std::vector<size_t> fan_chunks(size_t n) {
    const size_t c = (n + kFan - 1) / kFan;
    std::vector<size_t> sizes(c, n / c);
    for (size_t i = 0; i < n % c; i++) sizes[i]++;
    return sizes;
}

// AND and OR are idempotent: an operand that occurs twice (identical windows over trivial chars share one node through
// CSE) is kept once, so that a sum never carries a coefficient > 1 on one PBS output (its noise would count squared)
static void dedupe(std::vector<BlockId>& v) {
    std::vector<BlockId> out;
    for (auto b : v) if (std::find(out.begin(), out.end(), b) == out.end()) out.push_back(b);
    v.swap(out);
}

BlockId Graph::and_flags(std::vector<BlockId> cur) {
    dedupe(cur);
    if (cur.empty()) return trivial_block(1);
    while (cur.size() > 1) {
        std::vector<BlockId> nxt;
        size_t i = 0;
        for (size_t k : fan_chunks(cur.size())) {
            if (k == 1) { nxt.push_back(cur[i++]); continue; }
            std::vector<std::pair<BlockId, int>> ops;
            for (size_t j = 0; j < k; j++) ops.push_back({cur[i + j], 1});
            nxt.push_back(pbs(ops, kFan - (int)k, threshold_table()));
            i += k;
        }
        cur.swap(nxt);
    }
    return cur[0];
}

BlockId Graph::or_flags(std::vector<BlockId> cur) {
    dedupe(cur);
    if (cur.empty()) return trivial_block(0);
    while (cur.size() > 1) {
        std::vector<BlockId> nxt;
        size_t i = 0;
        for (size_t k : fan_chunks(cur.size())) {
            if (k == 1) { nxt.push_back(cur[i++]); continue; }
            std::vector<std::pair<BlockId, int>> ops;
            for (size_t j = 0; j < k; j++) ops.push_back({cur[i + j], 1});
            nxt.push_back(pbs(ops, kFan - 1, threshold_table()));
            i += k;
        }
        cur.swap(nxt);
    }
    return cur[0];
}

Char Graph::and_all(const std::vector<Char>& flags) {
    std::vector<BlockId> cur;
    for (auto& c : flags) cur.push_back(cond_bit(c));
    return flag_char(and_flags(cur));
}

Char Graph::or_all(const std::vector<Char>& flags) {
    std::vector<BlockId> cur;
    for (auto& c : flags) cur.push_back(cond_bit(c));
    return flag_char(or_flags(cur));
}

// OR over windows of (AND over the window's flags): one level per 16-fold of each (contains over 8-char windows: the
// 16 nibble flags of a window are ONE threshold PBS, 250 windows are 16 + 1 more)
Char Graph::or_of_ands(const std::vector<std::vector<BlockId>>& windows) {
    std::vector<BlockId> w;
    for (auto& flags : windows) {
        if (flags.empty()) return trivial_char(1);   // an empty AND is true
        w.push_back(and_flags(flags));
    }
    return flag_char(or_flags(w));
}

// column compression: each column holds blocks of weight 4^c; chunks whose maximum sum is <= 15 are
// replaced by (sum & 3) in the same column and (sum >> 2) in the next one.  Chunks are cut greedily from
// the left, so sums over prefixes of one flag list share their leading chunks through CSE.
std::vector<BlockId> Graph::sum_digits(const std::vector<BlockId>& flags, int ndigits) {
    std::vector<std::vector<BlockId>> col(ndigits + 1);
    col[0] = flags;
    auto vmax = [&](BlockId b) { int m = 0; for (int v = 0; v < 16; v++) if (nodes[b].vset >> v & 1) m = v; return m; };
    for (int guard = 0; guard < 64; guard++) {
        bool busy = false;
        std::vector<std::vector<BlockId>> nxt(ndigits + 1);
        for (int c = 0; c < ndigits; c++) {
            if (col[c].size() <= 1) { for (auto b : col[c]) nxt[c].push_back(b); continue; }
            busy = true;
            size_t i = 0;
            while (i < col[c].size()) {
                std::vector<std::pair<BlockId, int>> ops;
                int total = 0;
                // a block that occurs m times in a chunk counts m^2 in the noise: close the chunk before the bound
                std::map<BlockId, int> mult;
                auto noise_with = [&](BlockId b) {
                    float s2 = 0;
                    for (auto& kv : mult) { const int m = kv.second + (kv.first == b ? 1 : 0); s2 += (float)m * m * nodes[kv.first].noise2; }
                    if (!mult.count(b)) s2 += nodes[b].noise2;
                    return s2;
                };
                while (i < col[c].size() && total + vmax(col[c][i]) <= 15 && ops.size() < FHESTR_MAX_TERMS &&
                       (ops.empty() || noise_with(col[c][i]) <= kNoise2Limit)) {
                    total += vmax(col[c][i]);
                    ops.push_back({col[c][i], 1});
                    mult[col[c][i]]++;
                    i++;
                }
                if (ops.size() == 1) { nxt[c].push_back(ops[0].first); continue; }
                nxt[c].push_back(pbs(ops, 0, table_of([](int v) { return v & 3; })));
                if (c < ndigits - 1 && total >= 4) nxt[c + 1].push_back(pbs(ops, 0, table_of([](int v) { return v >> 2; })));
            }
        }
        col.swap(nxt);
        if (!busy) break;
    }
    std::vector<BlockId> r(ndigits);
    for (int c = 0; c < ndigits; c++) r[c] = col[c].empty() ? trivial_block(0) : col[c][0];
    return r;
}

Char Graph::sum_flags(const std::vector<Char>& flags) {
    std::vector<BlockId> f;
    for (auto& c : flags) f.push_back(cond_bit(c));
    const std::vector<BlockId> d = sum_digits(f, 4);
    return Char{d[0], d[1], d[2], d[3]};
}

BlockId Graph::not_flag(BlockId b) { return lin({{b, -1}}, 1, 0x3); }

BlockId Graph::mul_flag(BlockId flag, BlockId blk) {
    if (is_trivial(flag)) return trivial_value(flag) ? blk : trivial_block(0);
    if (is_trivial(blk)) {
        const int v = trivial_value(blk);
        if (v == 0) return trivial_block(0);
        return lin({{flag, v}}, 0, 1u | (1u << v));
    }
    return bivar(blk, flag, [](int x, int c) { return (c & 1) ? x : 0; });
}

Char Graph::mul_flag_char(BlockId flag, const Char& c) {
    return Char{mul_flag(flag, c[0]), mul_flag(flag, c[1]), mul_flag(flag, c[2]), mul_flag(flag, c[3])};
}

Char Graph::add_disjoint(const std::vector<Char>& parts) {
    Char r;
    for (int b = 0; b < 4; b++) {
        std::vector<std::pair<BlockId, int>> ops;
        uint32_t vs = 1;
        for (auto& p : parts) { ops.push_back({p[b], 1}); vs |= nodes[p[b]].vset & 0xF; }
        r[b] = lin(ops, 0, vs);
    }
    return r;
}

// Stable compaction.  z_i = number of NUL chars before position i; a non-NUL char has to move left by z_i.
// Routing by the bits of z_i, least significant first, never makes two chars meet (for p < q non-NUL,
// z_q - z_p <= q - p - 1, and the partial shifts differ by at most that).  Every position carries its char
// blocks plus z as base-4 control digits (two routing bits per block, forced to 0 on NUL chars so that they
// "stay" and contribute nothing); layer b rebuilds position i as
//     stay(x_i, ctrl_i) + move(x_{i+2^b}, ctrl_{i+2^b})
// with two bivariate LUTs per carried block.  z_i comes from chunked prefix counts: an exact base count per
// chunk of 13 (sum_digits over the flag prefix, shared through CSE) plus the local count (<= 12, leveled),
// normalised by one carry chain.
// The routing layers of compact_nonzero for long strings, at about half the PBS (throughput matters there, depth does
// not).  Per layer only the MOVING part of a block is a PBS,  mv_i = (bit of the control digit y_i) ? x_i : 0  on
// x_i + 4 y_i; the staying part is the plain difference, so  x'_i = x_i - mv_i + mv_(i+sh)  is leveled.  The price is
// noise: x' gains two units per layer, so it carries weight 1 in the PBS input, the control digit weight 4 and the digit
// has to be a fresh PBS output: a digit is refreshed when it becomes the active one, routed between its two layers by
// ONE PBS on 4 y_i + y_(i+sh), and while it waits it is routed like the data.  Blocks are refreshed when their noise
// would no longer fit next to the control digit.  PBS per position at 1025 chars: 161 -> about 90.
std::vector<Char> Graph::route_linear(const std::vector<Char>& s, std::vector<std::vector<BlockId>> ctrl, int B, int nd) {
    const size_t L = s.size();
    std::vector<Char> cur = s;
    const BlockId zero = trivial_block(0);
    auto is_zero = [&](BlockId b) { return is_trivial(b) && trivial_value(b) == 0; };
    auto fresh_enough = [&](BlockId b) { return is_trivial(b) || nodes[b].noise2 + 16.f <= kNoise2Limit; };
    for (int b = 0; b < B; b++) {
        const size_t sh = (size_t)1 << b;
        const int kb = b / 2, bit = b & 1;
        auto move_tab = table_of([bit](int v) { return (((v >> 2) >> bit) & 1) ? (v & 3) : 0; });
        // the active digit as a fresh PBS output (digit 0 starts fresh; an odd layer gets it from the one-PBS routing below)
        std::vector<BlockId> y(L);
        for (size_t i = 0; i < L; i++) {
            BlockId d = ctrl[i][kb];
            if (!is_trivial(d) && (nodes[d].kind == BKind::Linear || nodes[d].noise2 > 1.f)) d = refresh(d);
            y[i] = d;
        }
        auto moving = [&](BlockId x, size_t i) {   // the part of x that leaves position i in this layer
            if (is_zero(x) || is_zero(y[i])) return zero;
            if (is_trivial(y[i])) return ((trivial_value(y[i]) >> bit) & 1) ? x : zero;
            if (is_trivial(x)) { const int cx = trivial_value(x); return pbs({{y[i], 1}}, 0, table_of([bit, cx](int v) { return ((v >> bit) & 1) ? cx : 0; })); }
            return pbs({{x, 1}, {y[i], 4}}, 0, move_tab);
        };
        auto routed = [&](BlockId x_here, BlockId mv_here, BlockId x_in, BlockId mv_in) {
            uint32_t vs = 1u | (nodes[x_here].vset & 0xF);
            std::vector<std::pair<BlockId, int>> ops{{x_here, 1}, {mv_here, -1}};
            if (!is_zero(mv_in)) { ops.push_back({mv_in, 1}); vs |= nodes[x_in].vset & 0xF; }
            return lin(ops, 0, vs);
        };
        std::vector<Char> mvd(L);
        std::vector<std::vector<BlockId>> mvc(L, std::vector<BlockId>(nd, zero));
        for (size_t i = 0; i < L; i++) {
            for (int q = 0; q < 4; q++) {
                if (!fresh_enough(cur[i][q])) cur[i][q] = refresh(cur[i][q]);
                mvd[i][q] = moving(cur[i][q], i);
            }
            for (int d = kb + 1; d < nd; d++) {
                if (!fresh_enough(ctrl[i][d])) ctrl[i][d] = refresh(ctrl[i][d]);
                mvc[i][d] = moving(ctrl[i][d], i);
            }
        }
        std::vector<Char> nxt(L);
        std::vector<std::vector<BlockId>> nctrl(L, std::vector<BlockId>(nd, zero));
        for (size_t i = 0; i < L; i++) {
            const bool has_in = i + sh < L;
            for (int q = 0; q < 4; q++)
                nxt[i][q] = routed(cur[i][q], mvd[i][q], has_in ? cur[i + sh][q] : zero, has_in ? mvd[i + sh][q] : zero);
            for (int d = kb + 1; d < nd; d++)
                nctrl[i][d] = routed(ctrl[i][d], mvc[i][d], has_in ? ctrl[i + sh][d] : zero, has_in ? mvc[i + sh][d] : zero);
            if (bit == 0) {   // the active digit is needed once more: stays unless its own bit 0 is set, the incoming one moves if its bit 0 is set
                const BlockId a = y[i], c = has_in ? y[i + sh] : zero;
                auto stay_v = [](int v) { return (v & 1) ? 0 : v; };
                auto move_v = [](int v) { return (v & 1) ? v : 0; };
                if (is_trivial(a) && is_trivial(c)) nctrl[i][kb] = trivial_block(stay_v(trivial_value(a)) + move_v(trivial_value(c)));
                else if (is_trivial(c)) { const int cc = move_v(trivial_value(c)); nctrl[i][kb] = pbs({{a, 1}}, 0, table_of([=](int v) { return stay_v(v & 3) + cc; })); }
                else if (is_trivial(a)) { const int ca = stay_v(trivial_value(a)); nctrl[i][kb] = pbs({{c, 1}}, 0, table_of([=](int v) { return ca + move_v(v & 3); })); }
                // (both non-zero cannot happen in a compaction: two elements would collide; & 3 keeps the value set a digit)
                else nctrl[i][kb] = pbs({{a, 4}, {c, 1}}, 0, table_of([=](int v) { return (stay_v(v >> 2) + move_v(v & 3)) & 3; }));
            }
        }
        cur.swap(nxt);
        ctrl.swap(nctrl);
    }
    return cur;
}

std::vector<Char> Graph::compact_nonzero(const std::vector<Char>& s) {
    const size_t L = s.size();
    if (L <= 1) return s;
    int B = 0;
    while ((1u << B) < L) B++;  // shifts are < L <= 2^B
    const int nd = (B + 1) / 2;
    std::vector<BlockId> nzb(L), zf(L);
    for (size_t i = 0; i < L; i++) { nzb[i] = cond_bit(s[i]); zf[i] = not_flag(nzb[i]); }
    const size_t m = 13;
    std::vector<std::vector<BlockId>> ctrl(L, std::vector<BlockId>(nd));
    auto msg_tab = table_of([](int v) { return v & 3; });
    auto car_tab = table_of([](int v) { return v >> 2; });
    for (size_t c0 = 0; c0 < L; c0 += m) {
        std::vector<BlockId> base = sum_digits(std::vector<BlockId>(zf.begin(), zf.begin() + c0), nd);
        for (size_t i = c0; i < std::min(L, c0 + m); i++) {
            std::vector<std::pair<BlockId, int>> loc;
            for (size_t k = c0; k < i; k++) loc.push_back({zf[k], 1});
            BlockId carry = lin(loc, 0);
            for (int d = 0; d < nd; d++) {
                const BlockId cur = lin({{base[d], 1}, {carry, 1}}, 0);
                BlockId digit;
                if (!(nodes[cur].vset & ~0xFu)) { digit = cur; carry = trivial_block(0); }
                else {
                    digit = pbs({{cur, 1}}, 0, msg_tab);
                    carry = d + 1 < nd ? pbs({{cur, 1}}, 0, car_tab) : trivial_block(0);
                }
                ctrl[i][d] = mul_flag(nzb[i], digit);
            }
        }
    }
    if (L > kLinearRoutingMinLength) return route_linear(s, ctrl, B, nd);
    std::vector<Char> cur = s;
    for (int b = 0; b < B; b++) {
        const size_t sh = (size_t)1 << b;
        const int kb = b / 2, bit = b & 1;
        const int first_ctrl = bit ? kb + 1 : kb;  // control digits still needed after this layer
        auto stay_f = [bit](int x, int y) { return ((y >> bit) & 1) ? 0 : x; };
        auto move_f = [bit](int x, int y) { return ((y >> bit) & 1) ? x : 0; };
        auto route = [&](BlockId x_here, BlockId y_here, BlockId x_in, BlockId y_in, bool has_in) {
            BlockId st = (x_here == y_here) ? pbs({{x_here, 1}}, 0, table_of([&](int v) { return stay_f(v & 3, v & 3); }))
                                            : bivar(x_here, y_here, stay_f);
            if (!has_in) return st;
            BlockId mv = (x_in == y_in) ? pbs({{x_in, 1}}, 0, table_of([&](int v) { return move_f(v & 3, v & 3); }))
                                        : bivar(x_in, y_in, move_f);
                    return lin({{st, 1}, {mv, 1}}, 0, 1u | ((nodes[x_here].vset | nodes[x_in].vset) & 0xF));
        };
        std::vector<Char> nxt(L);
        std::vector<std::vector<BlockId>> nctrl(L, std::vector<BlockId>(nd));
        for (size_t i = 0; i < L; i++) {
            const bool has_in = i + sh < L;
            const BlockId y_here = ctrl[i][kb];
            const BlockId y_in = has_in ? ctrl[i + sh][kb] : trivial_block(0);
            for (int q = 0; q < 4; q++) nxt[i][q] = route(cur[i][q], y_here, has_in ? cur[i + sh][q] : trivial_block(0), y_in, has_in);
            for (int d = 0; d < nd; d++)
                nctrl[i][d] = d >= first_ctrl ? route(ctrl[i][d], y_here, has_in ? ctrl[i + sh][d] : trivial_block(0), y_in, has_in)
                                              : trivial_block(0);
        }
        cur.swap(nxt);
        ctrl.swap(nctrl);
    }
    return cur;
}

Char Graph::nonzero(const Char& a) { return flag_char(cond_bit(a)); }

// nibble-wide equality: two blocks of each char are packed lo + 4 hi and SUBTRACTED raw, exactly like the
// reference's comparison recipe (cmp above; noise2 = 34): the difference lies in [-15, 15] and is zero iff both
// blocks agree, so one `is zero` LUT (negative inputs carry the padding bit and return -f(.) = 0) replaces two
// block-equality PBS.  A char equality is 2 PBS + the AND tree instead of 4.
BlockId Graph::nibble_eq(BlockId a_lo, BlockId a_hi, BlockId b_lo, BlockId b_hi) {
    if (a_lo == b_lo && a_hi == b_hi) return trivial_block(1);
    SignedScope sc(signed_ok);
    return pbs({{a_lo, 1}, {a_hi, 4}, {b_lo, -1}, {b_hi, -4}}, 0, table_of([](int v) { return v == 0; }));
}

std::vector<BlockId> Graph::nibble_eq_flags(const std::vector<std::pair<Char, Char>>& pairs) {
    std::vector<BlockId> flags;
    for (auto& pr : pairs)
        for (int h = 0; h < 2; h++)
            flags.push_back(nibble_eq(pr.first[2 * h], pr.first[2 * h + 1], pr.second[2 * h], pr.second[2 * h + 1]));
    return flags;
}

Char Graph::block_and_eq(const std::vector<std::pair<Char, Char>>& pairs) {
    std::vector<Char> flags;
    for (auto& pr : pairs)
        for (int h = 0; h < 2; h++)
            flags.push_back(flag_char(nibble_eq(pr.first[2 * h], pr.first[2 * h + 1], pr.second[2 * h], pr.second[2 * h + 1])));
    return and_all(flags);
}

// first_i = m_i AND no earlier flag.  Inside a chunk of <= 15 flags: u = sum_{k<i} m_k - m_i + 1 is 0
// exactly for the first set flag; chunks are ranked the same way on their ANY flags, recursively.
std::vector<Char> Graph::first_one_hot(const std::vector<Char>& flags, Char* any) {
    const size_t n = flags.size();
    std::vector<BlockId> m;
    for (auto& c : flags) m.push_back(cond_bit(c));
    auto is_zero_tab = table_of([](int v) { return v == 0; });
    auto first_in = [&](const std::vector<BlockId>& v) {
        std::vector<BlockId> out;
        for (size_t i = 0; i < v.size(); i++) {
            std::vector<BlockId> earlier(v.begin(), v.begin() + i);
            dedupe(earlier);   // "some earlier flag is set" does not count repeats
            if (std::find(earlier.begin(), earlier.end(), v[i]) != earlier.end()) { out.push_back(trivial_block(0)); continue; }
            std::vector<std::pair<BlockId, int>> ops;
            for (auto b : earlier) ops.push_back({b, 1});
            ops.push_back({v[i], -1});
            out.push_back(pbs(ops, 1, is_zero_tab));
        }
        return out;
    };
    std::vector<Char> res;
    if (n <= (size_t)kChunk) {
        for (auto b : first_in(m)) res.push_back(flag_char(b));
        if (any) {
            std::vector<Char> tmp;
            for (auto b : m) tmp.push_back(flag_char(b));
            *any = or_all(tmp);
        }
        return res;
    }
    std::vector<Char> chunk_any;
    std::vector<std::vector<BlockId>> local;
    for (size_t i = 0; i < n; i += kChunk) {
        std::vector<BlockId> ch(m.begin() + i, m.begin() + std::min(n, i + kChunk));
        std::vector<Char> tmp;
        for (auto b : ch) tmp.push_back(flag_char(b));
        chunk_any.push_back(or_all(tmp));
        local.push_back(first_in(ch));
    }
    std::vector<Char> chunk_first = first_one_hot(chunk_any, any);
    for (size_t c = 0; c < local.size(); c++)
        for (auto b : local[c])
            res.push_back(flag_char(bivar(b, chunk_first[c][0], [](int x, int y) { return x & y & 1; })));
    return res;
}

Char Graph::select_by_one_hot(const std::vector<Char>& onehot, const std::vector<uint8_t>& values, const Char& none,
                              uint8_t none_value) {
    // per block: digit = 1*[any flag whose digit is 1] + 2*[... is 2] + 3*[... is 3]; the three class flags
    // are OR trees (the flags are one-hot, so at most one class fires) and the result is a clean digit
    Char r;
    for (int blk = 0; blk < 4; blk++) {
        std::vector<std::pair<BlockId, int>> ops;
        for (int d = 1; d <= 3; d++) {
            std::vector<Char> cls;
            for (size_t i = 0; i < onehot.size(); i++)
                if (((values[i] >> (2 * blk)) & 3) == d) cls.push_back(onehot[i]);
            if (((none_value >> (2 * blk)) & 3) == d) cls.push_back(none);
            if (cls.empty()) continue;
            ops.push_back({or_all(cls)[0], d});
        }
        r[blk] = lin(ops, 0, 0xF);
    }
    return r;
}

// ------------------------------------------------------------------------------------ compile
// Incremental: every call emits the jobs of the nodes that are reachable from the outputs marked since
// the last commit() and have not been computed yet.  After the program has run, commit() turns everything
// that now has an arena slot into level-0 atoms, so recording can simply continue (the eager path of the
// reference's one-op-at-a-time callers is "record one op, compile, run, commit").
bool Graph::compile(CompiledProgram& out, std::string& err, uint32_t slot_align) {
    if (!error.empty()) { err = error; return false; }
    if (slot_align == 0) slot_align = 1;
    out = CompiledProgram();
    std::vector<char> live(nodes.size(), 0);
    std::vector<BlockId> stack;
    for (auto b : outputs) {
        if (nodes[b].kind == BKind::Linear) nodes[b].materialized = true;
        if (!live[b]) { live[b] = 1; stack.push_back(b); }
    }
    while (!stack.empty()) {
        const BlockId b = stack.back();
        stack.pop_back();
        if (nodes[b].done) continue;
        for (auto& t : nodes[b].terms)
            if (!live[t.blk]) { live[t.blk] = 1; stack.push_back(t.blk); }
    }
    for (auto b : outputs)
        if (nodes[b].kind == BKind::Trivial && nodes[b].slot < 0) {
            nodes[b].slot = (int32_t)next_slot++;
            out.trivial_slots.push_back({(uint32_t)nodes[b].slot, (uint8_t)nodes[b].cst});
        }
    int depth = 0;
    for (size_t b = 0; b < nodes.size(); b++) if (live[b] && !nodes[b].done) depth = std::max(depth, (int)nodes[b].level);
    std::vector<std::vector<BlockId>> pbs_at(depth + 1), lin_at(depth + 1);
    for (size_t b = 0; b < nodes.size(); b++) {
        if (!live[b] || nodes[b].done) continue;
        if (nodes[b].kind == BKind::Pbs) pbs_at[nodes[b].level].push_back((BlockId)b);
        else if (nodes[b].kind == BKind::Linear && nodes[b].materialized) lin_at[nodes[b].level].push_back((BlockId)b);
    }
    for (int l = 0; l <= depth; l++) {
        // the PBS results of a level are contiguous and padded to a multiple of slot_align, so that with
        // `slot_align` ranks every rank's share is one equal slice of one all-gather buffer
        for (auto b : pbs_at[l]) nodes[b].slot = (int32_t)next_slot++;
        if (!pbs_at[l].empty()) next_slot += (slot_align - pbs_at[l].size() % slot_align) % slot_align;
        for (auto b : lin_at[l]) nodes[b].slot = (int32_t)next_slot++;
    }
    out.n_slots = next_slot;
    out.n_inputs = n_input_nodes;
    auto emit = [&](BlockId b) {
        const BlockNode& n = nodes[b];
        fhestr_job j;
        memset(&j, 0, sizeof j);
        j.dst = (uint32_t)n.slot;
        j.lut = n.kind == BKind::Pbs ? n.lut : -1;
        j.n_terms = (uint32_t)n.terms.size();
        for (size_t t = 0; t < n.terms.size(); t++) {
            const BlockNode& s = nodes[n.terms[t].blk];
            if (s.slot < 0) { err = "internal: job source has no slot"; return false; }
            j.src[t] = (uint32_t)s.slot;
            j.coeff[t] = n.terms[t].coeff;
        }
        j.constant = ((uint64_t)mod32(n.cst)) << delta_log;
        out.jobs.push_back(j);
        return true;
    };
    out.level_offsets.push_back(0);
    {
        size_t n_jobs = 0;
        for (int l = 0; l <= depth; l++) n_jobs += pbs_at[l].size() + lin_at[l].size();
        out.jobs.reserve(n_jobs);
    }
    for (int l = 0; l <= depth; l++) {
        if (pbs_at[l].empty() && lin_at[l].empty()) continue;
        out.level_first_dst.push_back(pbs_at[l].empty() ? 0u : (uint32_t)nodes[pbs_at[l][0]].slot);
        for (auto b : pbs_at[l]) if (!emit(b)) return false;
        for (auto b : lin_at[l]) if (!emit(b)) return false;
        out.level_pbs.push_back((uint32_t)pbs_at[l].size());
        out.n_pbs += pbs_at[l].size();
        out.level_offsets.push_back((uint32_t)out.jobs.size());
    }
    pending.clear();
    for (size_t b = 0; b < nodes.size(); b++)
        if (live[b] && !nodes[b].done && nodes[b].slot >= 0) pending.push_back((BlockId)b);
    return true;
}

void Graph::commit() {
    for (auto b : pending) { nodes[b].done = true; nodes[b].level = 0; }
    pending.clear();
    outputs.clear();
    // values that exist are level-0 atoms now; re-level whatever has been recorded but not computed
    for (size_t b = 0; b < nodes.size(); b++) {
        BlockNode& n = nodes[b];
        if (n.done || n.terms.empty()) continue;
        if (n.kind == BKind::Pbs) n.level = 1 + level_of(n.terms, false);
        else if (n.kind == BKind::Linear) n.level = level_of(n.terms, true);
    }
}

}  // namespace fhestr
