// strings.cpp -- see strings.h.  Each method names the reference function it restates; the faithful
// branch issues the reference's own primitive sequence, the fast branch a plaintext-identical rewrite.
#include "strings.h"

#include <algorithm>

namespace fhestr {

static const size_t kMaxFindLength = 255;   // /root/reference/src/main.rs:20
static const size_t kMaxRepetitions = 16;   // /root/reference/src/main.rs:17

static size_t adjust_end_of_pattern(size_t e) { return e == 0 ? 1 : e; }  // utils.rs:106

// ------------------------------------------------------------------------------------ helpers
Char StringOps::match_at(const Str& s, size_t i, const Str& pattern, bool reversed) {
    if (fast) {
        std::vector<std::pair<Char, Char>> pairs;
        for (size_t j = 0; j < pattern.size(); j++) pairs.push_back({pattern[j], s[i + j]});
        return g.block_and_eq(pairs);
    }
    Char flag = one();
    for (size_t jj = 0; jj < pattern.size(); jj++) {
        const size_t j = reversed ? pattern.size() - 1 - jj : jj;
        flag = g.bitand_(flag, g.eq(pattern[j], s[i + j]));
    }
    return flag;
}

Char StringOps::char_cmp(const Char& a, const Char& b, int op) {
    switch (op) {
        case 0: return g.lt(a, b);
        case 1: return g.le(a, b);
        case 2: return g.gt(a, b);
        default: return g.ge(a, b);
    }
}

std::vector<Char> StringOps::last_one_hot(const std::vector<Char>& flags, Char* any) {
    std::vector<Char> rev(flags.rbegin(), flags.rend());
    std::vector<Char> r = g.first_one_hot(rev, any);
    std::reverse(r.begin(), r.end());
    return r;
}

// 1 unless c is NUL or ASCII whitespace (0x09-0x0D, 0x20): two nibble classifiers and one combiner
Char StringOps::is_not_blank(const Char& c) {
    std::array<uint8_t, 16> th{}, tl{}, tc{};
    for (int v = 0; v < 16; v++) {
        th[v] = v == 0 ? 1 : (v == 2 ? 2 : 0);
        tl[v] = v == 0 ? 1 : ((v >= 9 && v <= 13) ? 2 : 0);
    }
    for (int v = 0; v < 16; v++) {
        const int h = v >> 2, l = v & 3;
        const bool blank = (h == 1 && l >= 1) || (h == 2 && l == 1);
        tc[v] = blank ? 0 : 1;
    }
    const BlockId h = g.pbs({{c[3], 4}, {c[2], 1}}, 0, th);
    const BlockId l = g.pbs({{c[1], 4}, {c[0], 1}}, 0, tl);
    return g.flag_char(g.pbs({{h, 4}, {l, 1}}, 0, tc));
}

// 1 iff c is ASCII whitespace (0x09-0x0D, 0x20; NUL is not): the same two nibble classifiers, another combiner
Char StringOps::is_blank_not_nul(const Char& c) {
    std::array<uint8_t, 16> th{}, tl{}, tc{};
    for (int v = 0; v < 16; v++) {
        th[v] = v == 0 ? 1 : (v == 2 ? 2 : 0);
        tl[v] = v == 0 ? 1 : ((v >= 9 && v <= 13) ? 2 : 0);
    }
    for (int v = 0; v < 16; v++) {
        const int h = v >> 2, l = v & 3;
        tc[v] = ((h == 1 && l == 2) || (h == 2 && l == 1)) ? 1 : 0;
    }
    const BlockId h = g.pbs({{c[3], 4}, {c[2], 1}}, 0, th);
    const BlockId l = g.pbs({{c[1], 4}, {c[0], 1}}, 0, tl);
    return g.flag_char(g.pbs({{h, 4}, {l, 1}}, 0, tc));
}

// flag: c in [first, first + 25] for first = 0x41 ('A'..'Z') or 0x61 ('a'..'z')
static Char letter_range_flag(Graph& g, const Char& c, int first) {
    const int row = first >> 4;
    std::array<uint8_t, 16> th{}, tl{}, tc{};
    for (int v = 0; v < 16; v++) {
        th[v] = v == row ? 1 : (v == row + 1 ? 2 : 0);
        tl[v] = v == 0 ? 0 : (v <= 10 ? 1 : 2);
    }
    for (int v = 0; v < 16; v++) {
        const int h = v >> 2, l = v & 3;
        tc[v] = ((h == 1 && l >= 1) || (h == 2 && l <= 1)) ? 1 : 0;
    }
    const BlockId h = g.pbs({{c[3], 4}, {c[2], 1}}, 0, th);
    const BlockId l = g.pbs({{c[1], 4}, {c[0], 1}}, 0, tl);
    return g.flag_char(g.pbs({{h, 4}, {l, 1}}, 0, tc));
}

// suffix OR of 0/1 blocks: out[i] = OR_{k >= i} f[k]; chunks of 15, chunk ORs scanned recursively
static std::vector<BlockId> suffix_or(Graph& g, const std::vector<BlockId>& f) {
    const size_t n = f.size(), m = 14;
    std::array<uint8_t, 16> nz{};
    for (int v = 1; v < 16; v++) nz[v] = 1;
    std::vector<BlockId> out(n);
    if (n == 0) return out;
    std::vector<BlockId> chunk_or;
    // OR is idempotent: a flag that occurs twice (shared node) is summed once, so no coefficient exceeds 1
    auto add_once = [](std::vector<std::pair<BlockId, int>>& ops, BlockId b) {
        for (auto& o : ops) if (o.first == b) return;
        ops.push_back({b, 1});
    };
    for (size_t c0 = 0; c0 < n; c0 += m) {
        std::vector<std::pair<BlockId, int>> ops;
        for (size_t k = c0; k < std::min(n, c0 + m); k++) add_once(ops, f[k]);
        chunk_or.push_back(ops.size() == 1 ? ops[0].first : g.pbs(ops, 0, nz));
    }
    std::vector<BlockId> later;  // later[c] = OR of chunks > c
    if (chunk_or.size() > 1) {
        std::vector<BlockId> tail(chunk_or.begin() + 1, chunk_or.end());
        later = suffix_or(g, tail);
    }
    for (size_t c0 = 0, c = 0; c0 < n; c0 += m, c++) {
        const size_t hi = std::min(n, c0 + m);
        for (size_t i = c0; i < hi; i++) {
            std::vector<std::pair<BlockId, int>> ops;
            for (size_t k = i; k < hi; k++) add_once(ops, f[k]);
            if (c < later.size()) add_once(ops, later[c]);
            out[i] = ops.size() == 1 ? ops[0].first : g.pbs(ops, 0, nz);
        }
    }
    return out;
}

// ------------------------------------------------------------------------------------ utils.rs
Str StringOps::bubble_zeroes_right(const Str& s) {
    if (fast) return g.compact_nonzero(s);
    Str r = s;
    const size_t n = r.size();
    for (size_t pass = 0; pass < n; pass++)
        for (size_t i = 0; i + 1 < n; i++) {
            const Char should_swap = g.eq(r[i], zero());
            const Char a = g.if_then_else(should_swap, r[i + 1], r[i]);
            const Char b = g.if_then_else(should_swap, zero(), r[i + 1]);
            r[i] = a;
            r[i + 1] = b;
        }
    return r;
}

// ------------------------------------------------------------------------------------ mod.rs
Str StringOps::to_upper(const Str& s) {
    Str out;
    const Char cst = g.trivial_char(32);  // fhestring.rs:24
    for (auto& b : s) {
        if (fast) {
            const BlockId f = letter_range_flag(g, b, 0x61)[0];
            Char r = b;
            r[2] = g.lin({{b[2], 1}, {f, -2}}, 0, 0xF);
            out.push_back(r);
        } else {
            const Char not_lower = g.flip(g.is_lowercase(b));
            out.push_back(g.sub(b, g.if_then_else(not_lower, zero(), cst)));
        }
    }
    return out;
}

Str StringOps::to_lower(const Str& s) {
    Str out;
    const Char cst = g.trivial_char(32);
    for (auto& b : s) {
        if (fast) {
            const BlockId f = letter_range_flag(g, b, 0x41)[0];
            Char r = b;
            r[2] = g.lin({{b[2], 1}, {f, 2}}, 0, 0xF);
            out.push_back(r);
        } else {
            const Char not_upper = g.flip(g.is_uppercase(b));
            out.push_back(g.add(b, g.if_then_else(not_upper, zero(), cst)));
        }
    }
    return out;
}

Char StringOps::contains(const Str& s, const Str& needle) {
    if (s.empty() && needle.empty()) return one();
    if (needle.size() > s.size()) return zero();
    const size_t end = s.size() - needle.size();
    if (fast) {
        std::vector<std::vector<BlockId>> windows;
        for (size_t i = 0; i <= end; i++) {
            std::vector<std::pair<Char, Char>> pairs;
            for (size_t j = 0; j < needle.size(); j++) pairs.push_back({needle[j], s[i + j]});
            windows.push_back(g.nibble_eq_flags(pairs));
        }
        return g.or_of_ands(windows);
    }
    Char result = zero();
    for (size_t i = 0; i <= end; i++) {
        Char cur = one();
        for (size_t j = 0; j < needle.size(); j++) cur = g.bitand_(cur, g.eq(s[i + j], needle[j]));
        result = g.bitor_(result, cur);
    }
    return result;
}

Char StringOps::ends_with(const Str& s, const Str& needle) {
    if (s.empty() && needle.empty()) return one();
    if (needle.size() > s.size()) return zero();
    const size_t end = s.size() - needle.size();
    if (fast) {
        // the last window made only of non-NUL chars decides
        std::vector<Char> all_nonzero, match;
        for (size_t i = 0; i <= end; i++) {
            std::vector<Char> nz;
            for (size_t j = 0; j < needle.size(); j++) nz.push_back(g.nonzero(s[i + j]));
            all_nonzero.push_back(g.and_all(nz));
            match.push_back(match_at(s, i, needle, false));
        }
        const std::vector<Char> last = last_one_hot(all_nonzero, nullptr);
        std::vector<Char> terms;
        for (size_t i = 0; i <= end; i++) terms.push_back(g.flag_char(g.mul_flag(last[i][0], match[i][0])));
        return g.or_all(terms);
    }
    Char result = zero();
    for (size_t i = 0; i <= end; i++) {
        Char cur = one(), nonzero = one();
        for (size_t j = 0; j < needle.size(); j++) {
            cur = g.bitand_(cur, g.eq(s[i + j], needle[j]));
            nonzero = g.bitand_(nonzero, g.ne(s[i + j], zero()));
        }
        result = g.if_then_else(nonzero, cur, result);
    }
    return result;
}

Char StringOps::starts_with(const Str& s, const Str& pattern) {
    if (pattern.size() > s.size()) return zero();
    if (s.empty() && pattern.empty()) return one();
    if (fast) return match_at(s, 0, pattern, false);
    Char result = one();
    for (size_t j = 0; j < pattern.size(); j++) result = g.bitand_(result, g.eq(s[j], pattern[j]));
    return result;
}

Char StringOps::is_empty(const Str& s) {
    if (s.empty()) return one();
    if (fast) {
        std::vector<Char> nz;
        for (auto& c : s) nz.push_back(g.nonzero(c));
        return g.flag_char(g.not_flag(g.or_all(nz)[0]));
    }
    Char result = one();
    for (auto& c : s) result = g.bitand_(result, g.eq(c, zero()));
    return result;
}

Char StringOps::len(const Str& s) {
    if (s.empty()) return zero();
    if (fast) {
        std::vector<Char> nz;
        for (auto& c : s) nz.push_back(g.nonzero(c));
        return g.sum_flags(nz);
    }
    Char result = zero();
    for (auto& c : s) result = g.add(result, g.ne(c, zero()));
    return result;
}

Str StringOps::repeat_clear(const Str& s, size_t repetitions) {
    if (repetitions == 0) return Str();
    Str r;
    for (size_t k = 0; k < repetitions; k++) r.insert(r.end(), s.begin(), s.end());
    return bubble_zeroes_right(r);
}

Str StringOps::repeat(const Str& s, const Char& repetitions) {
    const size_t n = s.size();
    Str r(kMaxRepetitions * n, zero());
    for (size_t i = 0; i < kMaxRepetitions; i++) {
        const Char copy_flag = g.lt(g.trivial_char((uint8_t)i), repetitions);
        for (size_t j = 0; j < n; j++)
            r[i * n + j] = fast ? g.mul_flag_char(copy_flag[0], s[j]) : g.if_then_else(copy_flag, s[j], zero());
    }
    return bubble_zeroes_right(r);
}

Str StringOps::handle_longer_from(const Str& bytes_in, const Str& from, Str to, const Char& n, bool use_counter) {
    Str bytes = bytes_in;
    bytes.push_back(zero());
    while (to.size() < from.size()) to.push_back(zero());
    Str result = bytes;
    if (from.size() > result.size()) return bubble_zeroes_right(result);
    const size_t end = adjust_end_of_pattern(result.size() - from.size());
    if (!fast) {
        Char counter = zero();
        for (size_t i = 0; i < end; i++) {
            Char flag = one();
            for (size_t j = 0; j < from.size(); j++) flag = g.bitand_(flag, g.eq(from[j], bytes[i + j]));
            if (use_counter) {
                counter = g.add(counter, flag);
                flag = g.bitand_(flag, g.ge(n, counter));
            }
            for (size_t k = 0; k < to.size(); k++) result[i + k] = g.if_then_else(flag, to[k], result[i + k]);
        }
        return bubble_zeroes_right(result);
    }
    // all match flags come from the ORIGINAL bytes (mod.rs:864), so they are independent
    std::vector<Char> m(end);
    for (size_t i = 0; i < end; i++) m[i] = match_at(bytes, i, from, false);
    if (use_counter) {
        // counter after window i = (number of RAW matches in windows 0..i) mod 256; prefix sums share
        // their leading chunks through CSE
        const std::vector<Char> raw = m;
        for (size_t i = 0; i < end; i++) {
            const Char counter = g.sum_flags(std::vector<Char>(raw.begin(), raw.begin() + i + 1));
            m[i] = g.flag_char(g.mul_flag(g.ge(n, counter)[0], raw[i][0]));
        }
    }
    const size_t T = to.size();
    for (size_t p = 0; p < result.size(); p++) {
        // windows i = p, p-1, ..., p-T+1 write position p; the largest i (written last) wins
        std::vector<Char> cand;
        std::vector<size_t> kk;
        for (size_t k = 0; k < T && k <= p; k++) {
            const size_t i = p - k;
            if (i >= end) continue;
            cand.push_back(m[i]);
            kk.push_back(k);
        }
        if (cand.empty()) continue;
        const std::vector<Char> first = g.first_one_hot(cand, nullptr);
        std::vector<Char> parts;
        std::vector<std::pair<BlockId, int>> none_ops;
        for (size_t c = 0; c < cand.size(); c++) {
            parts.push_back(g.mul_flag_char(first[c][0], to[kk[c]]));
            none_ops.push_back({first[c][0], -1});
        }
        const BlockId none = g.lin(none_ops, 1, 0x3);
        parts.push_back(g.mul_flag_char(none, bytes[p]));
        result[p] = g.add_disjoint(parts);
    }
    return bubble_zeroes_right(result);
}

Str StringOps::handle_shorter_from(const Str& bytes_in, const Str& from, const Str& to, const Char& n, bool use_counter) {
    // op-by-op in both modes (the copy-buffer scheme is inherently serial in i)
    Str bytes = bytes_in;
    bytes.push_back(zero());
    const size_t size_difference = to.size() - from.size();
    size_t max_len = bytes.empty() ? to.size() : to.size() * bytes.size() + bytes.size();
    if (from.empty()) max_len = (bytes.size() + (bytes.size() + 1) * to.size()) + 1;
    Str result = bytes;
    while (result.size() < max_len) result.push_back(zero());
    Str copy_buffer(max_len, zero());
    Str ignore(max_len, one());
    Char counter = zero();
    for (size_t i = 0; i + to.size() < result.size(); i++) {
        Char flag = one();
        for (size_t j = 0; j < from.size(); j++) {
            flag = g.bitand_(flag, g.eq(from[j], result[i + j]));
            flag = g.bitand_(flag, ignore[i + j]);
        }
        if (from.empty()) flag = (i % (to.size() + 1) == 0) ? one() : zero();
        if (use_counter) {
            counter = g.add(counter, flag);
            flag = g.bitand_(flag, g.ge(n, counter));
        }
        for (size_t k = 0; k < max_len; k++) copy_buffer[k] = g.if_then_else(flag, result[k], zero());
        for (size_t k = 0; k < to.size(); k++) {
            result[i + k] = g.if_then_else(flag, to[k], result[i + k]);
            ignore[i + k] = g.bitand_(ignore[i + k], g.if_then_else(flag, zero(), one()));
        }
        for (size_t k = i + to.size(); k < max_len; k++)
            result[k] = g.if_then_else(flag, copy_buffer[k - size_difference], result[k]);
    }
    return result;
}

Str StringOps::replace(const Str& s, const Str& from, const Str& to) {
    if (from.size() >= to.size()) return handle_longer_from(s, from, to, zero(), false);
    return handle_shorter_from(s, from, to, zero(), false);
}

Str StringOps::replacen(const Str& s, const Str& from, const Str& to, const Char& n) {
    if (from.size() >= to.size()) return handle_longer_from(s, from, to, n, true);
    return handle_shorter_from(s, from, to, n, true);
}

bool StringOps::rfind(const Str& s_in, const Str& pattern, Char& out) {
    Str s = s_in;
    s.push_back(zero());
    if (s.size() >= kMaxFindLength + pattern.size()) {
        error = "Maximum supported size for find reached";
        return false;
    }
    if (pattern.empty()) {
        if (fast) {
            std::vector<Char> nz;
            std::vector<uint8_t> values;
            for (size_t i = 0; i < s.size(); i++) { nz.push_back(g.nonzero(s[i])); values.push_back((uint8_t)(i + 1)); }
            out = g.select_by_one_hot(last_one_hot(nz, nullptr), values, zero(), 0);
            return true;
        }
        Char last = zero();
        for (size_t i = 0; i < s.size(); i++)
            last = g.if_then_else(g.ne(s[i], zero()), g.trivial_char((uint8_t)(i + 1)), last);
        out = last;
        return true;
    }
    if (pattern.size() > s.size()) { out = g.trivial_char(255); return true; }
    const size_t end = adjust_end_of_pattern(s.size() - pattern.size());
    if (fast && end > 15) {
        std::vector<BlockId> m;
        for (size_t w = 0; w < end; w++) m.push_back(g.cond_bit(match_at(s, w, pattern, false)));
        out = first_index_fast(m, true);
        return true;
    }
    if (fast) {
        std::vector<Char> m;
        std::vector<uint8_t> values;
        for (size_t i = 0; i < end; i++) { m.push_back(match_at(s, i, pattern, false)); values.push_back((uint8_t)i); }
        Char any;
        const std::vector<Char> last = last_one_hot(m, &any);
        out = g.select_by_one_hot(last, values, g.flag_char(g.not_flag(any[0])), (uint8_t)kMaxFindLength);
        return true;
    }
    Char pos = g.trivial_char((uint8_t)kMaxFindLength);
    for (size_t i = 0; i < end; i++)
        pos = g.if_then_else(match_at(s, i, pattern, false), g.trivial_char((uint8_t)i), pos);
    out = pos;
    return true;
}

// find / rfind over more than 15 windows (at most 255, so at most 17 chunks of 15): index of the first (last) set flag,
// 255 if none.  Inside a chunk "match and no earlier match" is one first-in PBS per window; the chunks are ranked the same way
// on their ANY flags (the 16th and 17th chunk through one OR of the first fifteen); a result digit is then
//     sum_d d * [ some chunk c is the first one AND its first window has digit d ]
// where the inner flag is one PBS on  h_c + (sum of the chunk's first-in flags whose index has that digit)  -- the sum
// is 0/1 because at most one of them is set -- and the OR over the chunks is again an exclusive sum, cleaned by one PBS.
// Eight levels where the one-hot over all windows followed by two levels of class ORs took ten.
Char StringOps::first_index_fast(const std::vector<BlockId>& flags, bool last) {
    std::array<uint8_t, 16> is_zero_tab{}, nz_tab{}, is2_tab{};
    is_zero_tab[0] = 1;
    is2_tab[2] = 1;
    for (int v = 1; v < 16; v++) nz_tab[v] = 1;
    const size_t windows = flags.size();
    // position k of the ranking order is window index_of[k]: ascending for the first match, descending for the last
    std::vector<size_t> index_of(windows);
    for (size_t k = 0; k < windows; k++) index_of[k] = last ? windows - 1 - k : k;
    std::vector<BlockId> m(windows);
    for (size_t k = 0; k < windows; k++) m[k] = flags[index_of[k]];
    const size_t C = (windows + 14) / 15;
    std::vector<BlockId> fm(windows), any(C), h(C);
    for (size_t c = 0; c < C; c++) {
        const size_t lo = 15 * c, hi = std::min(windows, lo + 15);
        std::vector<std::pair<BlockId, int>> all;
        for (size_t w = lo; w < hi; w++) {
            std::vector<std::pair<BlockId, int>> ops;
            for (size_t k = lo; k < w; k++) ops.push_back({m[k], 1});
            ops.push_back({m[w], -1});
            fm[w] = w == lo ? m[w] : g.pbs(ops, 1, is_zero_tab);
            all.push_back({m[w], 1});
        }
        any[c] = all.size() == 1 ? all[0].first : g.pbs(all, 0, nz_tab);
    }
    // rank the chunks
    BlockId overall;
    {
        const size_t direct = std::min<size_t>(C, 15);
        std::vector<std::pair<BlockId, int>> head;
        for (size_t c = 0; c < direct; c++) {
            std::vector<std::pair<BlockId, int>> ops;
            for (size_t k = 0; k < c; k++) ops.push_back({any[k], 1});
            ops.push_back({any[c], -1});
            h[c] = c == 0 ? any[0] : g.pbs(ops, 1, is_zero_tab);
            head.push_back({any[c], 1});
        }
        BlockId pre = head.size() == 1 ? head[0].first : g.pbs(head, 0, nz_tab);   // OR of the first fifteen
        std::vector<std::pair<BlockId, int>> seen{{pre, 1}};
        for (size_t c = direct; c < C; c++) {
            std::vector<std::pair<BlockId, int>> ops = seen;
            ops.push_back({any[c], -1});
            h[c] = g.pbs(ops, 1, is_zero_tab);
            seen.push_back({any[c], 1});
        }
        overall = seen.size() == 1 ? pre : g.pbs(seen, 0, nz_tab);
    }
    const BlockId none = g.not_flag(overall);
    Char r;
    for (int q = 0; q < 4; q++) {
        std::vector<std::pair<BlockId, int>> digit_ops;
        for (int d = 1; d <= 3; d++) {
            std::vector<std::pair<BlockId, int>> lo_half, hi_half;   // two partial sums: a leveled job takes 16 terms
            for (size_t c = 0; c < C; c++) {
                std::vector<std::pair<BlockId, int>> cls;
                for (size_t w = 15 * c; w < std::min(windows, 15 * c + 15); w++)
                    if ((int)((index_of[w] >> (2 * q)) & 3) == d) cls.push_back({fm[w], 1});
                if (cls.empty()) continue;
                const BlockId in_chunk = g.lin(cls, 0, 0x3);
                if (g.is_trivial(in_chunk) && g.trivial_value(in_chunk) == 0) continue;
                const BlockId hit = g.pbs({{h[c], 1}, {in_chunk, 1}}, 0, is2_tab);
                (c < 9 ? lo_half : hi_half).push_back({hit, 1});
            }
            if (((kMaxFindLength >> (2 * q)) & 3) == (size_t)d) hi_half.push_back({none, 1});
            std::vector<std::pair<BlockId, int>> both;
            if (!lo_half.empty()) both.push_back({g.lin(lo_half, 0, 0x3), 1});
            if (!hi_half.empty()) both.push_back({g.lin(hi_half, 0, 0x3), 1});
            if (both.empty()) continue;
            const BlockId sum = g.lin(both, 0, 0x3);
            const size_t n_terms = lo_half.size() + hi_half.size();
            digit_ops.push_back({n_terms == 1 ? sum : g.pbs({{sum, 1}}, 0, nz_tab), d});
        }
        r[q] = g.lin(digit_ops, 0, 0xF);
    }
    return r;
}

bool StringOps::find(const Str& s, const Str& pattern, Char& out) {
    if (s.empty() && pattern.empty()) { out = zero(); return true; }
    if (s.size() >= kMaxFindLength + pattern.size()) {
        error = "Maximum supported size for find reached";
        return false;
    }
    if (pattern.size() > s.size()) { out = g.trivial_char(255); return true; }
    const size_t end = s.size() - pattern.size();
    if (fast && end + 1 > 15) {
        std::vector<BlockId> m;
        for (size_t w = 0; w <= end; w++) m.push_back(g.cond_bit(match_at(s, w, pattern, true)));
        out = first_index_fast(m, false);
        return true;
    }
    if (fast) {
        std::vector<Char> m;
        std::vector<uint8_t> values;
        for (size_t i = 0; i <= end; i++) { m.push_back(match_at(s, i, pattern, true)); values.push_back((uint8_t)i); }
        Char any;
        const std::vector<Char> first = g.first_one_hot(m, &any);
        out = g.select_by_one_hot(first, values, g.flag_char(g.not_flag(any[0])), (uint8_t)kMaxFindLength);
        return true;
    }
    Char pos = g.trivial_char((uint8_t)kMaxFindLength);
    for (size_t ii = 0; ii <= end; ii++) {
        const size_t i = end - ii;
        pos = g.if_then_else(match_at(s, i, pattern, true), g.trivial_char((uint8_t)i), pos);
    }
    out = pos;
    return true;
}

Char StringOps::eq(const Str& s, const Str& o) {
    const size_t n = std::min(s.size(), o.size());
    if (fast && s.size() <= 255 && o.size() <= 255) {
        // (s_i == 0 && o_i == 0) || s_i == o_i  is just s_i == o_i.  With equal prefixes the two lengths (counts of
        // non-NUL chars, no u8 wrap below 256 chars) differ exactly when the longer buffer's tail holds a non-NUL char,
        // so "lengths equal" is "the tail is all NUL" and joins the same AND: no length is computed at all
        std::vector<std::pair<Char, Char>> pairs;
        for (size_t i = 0; i < n; i++) pairs.push_back({s[i], o[i]});
        std::vector<Char> flags;
        for (auto b : g.nibble_eq_flags(pairs)) flags.push_back(g.flag_char(b));
        const Str& longer = s.size() > o.size() ? s : o;
        for (size_t i = n; i < longer.size(); i++) flags.push_back(g.flag_char(g.not_flag(g.cond_bit(g.nonzero(longer[i])))));
        return g.and_all(flags);
    }
    const Char len1 = len(s), len2 = len(o);
    if (fast) {
        std::vector<std::pair<Char, Char>> pairs;
        for (size_t i = 0; i < n; i++) pairs.push_back({s[i], o[i]});
        pairs.push_back({len1, len2});
        return g.block_and_eq(pairs);
    }
    Char is_eq = one();
    const Char lengths_ne = g.ne(len1, len2);
    for (size_t i = 0; i < n; i++) {
        const Char are_equal = g.eq(s[i], o[i]);
        Char res = g.bitand_(g.eq(s[i], zero()), g.eq(o[i], zero()));
        res = g.bitor_(res, are_equal);
        is_eq = g.bitand_(is_eq, res);
    }
    return g.if_then_else(lengths_ne, zero(), is_eq);
}

Char StringOps::ne(const Str& s, const Str& o) {
    const Char e = eq(s, o);
    return fast ? g.flag_char(g.not_flag(e[0])) : g.flip(e);
}

Char StringOps::eq_ignore_case(const Str& s, const Str& o) { return eq(to_lower(s), to_lower(o)); }

StripResult StringOps::strip_prefix(const Str& s, const Str& pattern) {
    Str result = s;
    if (pattern.size() > result.size()) return StripResult{result, zero()};
    const size_t end = std::min(pattern.size(), result.size());
    Char flag = one();
    if (end == 0 && !pattern.empty() && s.empty()) flag = zero();
    if (fast) {
        if (end > 0) flag = match_at(result, 0, pattern, false);
        const BlockId keep = g.not_flag(flag[0]);
        for (size_t j = 0; j < end; j++) result[j] = g.mul_flag_char(keep, result[j]);
    } else {
        for (size_t j = 0; j < end; j++) flag = g.bitand_(flag, g.eq(pattern[j], result[j]));
        for (size_t j = 0; j < end; j++) result[j] = g.if_then_else(flag, zero(), result[j]);
    }
    return StripResult{bubble_zeroes_right(result), flag};
}

StripResult StringOps::strip_suffix(const Str& s_in, const Str& needle) {
    Str s = s_in;
    if (needle.size() > s.size()) return StripResult{s, zero()};
    const size_t end = s.size() - needle.size();
    if (fast && end < 255) {   // window indices stay below the reference's "none" sentinel 255 (pos is a u8)
        std::vector<Char> all_nonzero, match;
        for (size_t i = 0; i <= end; i++) {
            std::vector<Char> nz;
            for (size_t j = 0; j < needle.size(); j++) nz.push_back(g.nonzero(s[i + j]));
            all_nonzero.push_back(g.and_all(nz));
            match.push_back(match_at(s, i, needle, false));
        }
        const std::vector<Char> last = last_one_hot(all_nonzero, nullptr);
        std::vector<BlockId> sel(end + 1);
        std::vector<Char> sel_c;
        for (size_t i = 0; i <= end; i++) { sel[i] = g.mul_flag(last[i][0], match[i][0]); sel_c.push_back(g.flag_char(sel[i])); }
        const Char should = g.or_all(sel_c);
        for (size_t p = 0; p < s.size(); p++) {
            std::vector<std::pair<BlockId, int>> ops;
            for (size_t j = 0; j < needle.size() && j <= p; j++)
                if (p - j <= end) ops.push_back({sel[p - j], -1});
            if (ops.empty()) continue;
            s[p] = g.mul_flag_char(g.lin(ops, 1, 0x3), s[p]);
        }
        return StripResult{s, should};
    }
    const Char two_five_five = g.trivial_char(255);
    Char pos = two_five_five;
    for (size_t i = 0; i <= end; i++) {
        Char found = one(), nonzero = one();
        for (size_t j = 0; j < needle.size(); j++) {
            found = g.bitand_(found, g.eq(s[i + j], needle[j]));
            nonzero = g.bitand_(nonzero, g.ne(s[i + j], zero()));
        }
        const Char cur = g.if_then_else(found, g.trivial_char((uint8_t)i), two_five_five);
        pos = g.if_then_else(nonzero, cur, pos);
    }
    const Char should = g.ne(pos, two_five_five);
    for (size_t i = 0; i <= end; i++) {
        const Char mask = g.eq(g.trivial_char((uint8_t)i), pos);
        for (size_t j = 0; j < needle.size(); j++) s[i + j] = g.if_then_else(mask, zero(), s[i + j]);
    }
    return StripResult{s, should};
}

Char StringOps::comparison(const Str& s_in, const Str& o_in, int op) {
    Str s = s_in, o = o_in;
    size_t n = std::min(s.size(), o.size());
    if (n == 0) { s.push_back(zero()); o.push_back(zero()); n = 1; }
    if (fast && s.size() <= 255 && o.size() <= 255 && n <= 15 * 15) {
        // ret = comparison at the first differing index; without one, the length comparison decides.
        //  * per char ONE packed-pair comparison gives d_i (differs) and t_i (differs in the op's direction; at a
        //    differing index ge is gt and le is lt);
        //  * "t_i and no earlier difference" is one first-in PBS inside a chunk of 15 chars, and at most one of them
        //    is set per chunk, so their plain sum is the chunk's verdict; the same form ranks the chunks;
        //  * with equal prefixes the lengths (no u8 wrap below 256 chars) compare like "the longer buffer's tail
        //    holds a non-NUL char", a constant when both buffers have the same size.
        const bool greater = op >= 2;
        auto is_zero_tab = std::array<uint8_t, 16>{};
        is_zero_tab[0] = 1;
        std::array<uint8_t, 16> nz_tab{};
        for (int v = 1; v < 16; v++) nz_tab[v] = 1;
        std::vector<BlockId> d(n), t(n);
        for (size_t i = 0; i < n; i++) std::tie(d[i], t[i]) = g.differs_and_strict(s[i], o[i], greater);
        std::vector<BlockId> chunk_any, chunk_verdict;
        for (size_t c0 = 0; c0 < n; c0 += 15) {
            const size_t hi = std::min(n, c0 + 15);
            std::vector<std::pair<BlockId, int>> firsts, diffs;
            for (size_t i = c0; i < hi; i++) {
                std::vector<std::pair<BlockId, int>> ops;
                for (size_t k = c0; k < i; k++) ops.push_back({d[k], 1});
                ops.push_back({t[i], -1});
                firsts.push_back({i == c0 ? t[i] : g.pbs(ops, 1, is_zero_tab), 1});
                diffs.push_back({d[i], 1});
            }
            chunk_verdict.push_back(g.lin(firsts, 0, 0x3));
            chunk_any.push_back(diffs.size() == 1 ? diffs[0].first : g.pbs(diffs, 0, nz_tab));
        }
        std::vector<std::pair<BlockId, int>> verdicts, anys;
        for (size_t c = 0; c < chunk_any.size(); c++) {
            std::vector<std::pair<BlockId, int>> ops;
            for (size_t k = 0; k < c; k++) ops.push_back({chunk_any[k], 1});
            ops.push_back({chunk_verdict[c], -1});
            // chunk 0 has no earlier chunk; with several chunks its 15-term sum is refreshed to one flag (same level as
            // the others' first-in PBS), so the noise of the final sum stays at one unit per chunk
            if (c == 0) verdicts.push_back({chunk_any.size() > 1 ? g.pbs({{chunk_verdict[0], 1}}, 0, nz_tab) : chunk_verdict[0], 1});
            else verdicts.push_back({g.pbs(ops, 1, is_zero_tab), 1});
            anys.push_back({chunk_any[c], 1});
        }
        const BlockId at_first = g.lin(verdicts, 0, 0x3);
        const BlockId any = anys.size() == 1 ? anys[0].first : g.pbs(anys, 0, nz_tab);
        // length comparison with equal prefixes: len1 - len2 = +/- (non-NUL chars in the longer buffer's tail)
        BlockId by_len;
        if (s.size() == o.size()) by_len = g.trivial_block((op == 1 || op == 3) ? 1 : 0);
        else {
            const bool s_longer = s.size() > o.size();
            const Str& longer = s_longer ? s : o;
            std::vector<Char> tail;
            for (size_t i = n; i < longer.size(); i++) tail.push_back(g.nonzero(longer[i]));
            const BlockId T = g.or_all(tail)[0];
            // s longer: ge 1, gt T, le !T, lt 0;   o longer: le 1, lt T, ge !T, gt 0
            const int as_s_longer[4] = {0 /*lt*/, 2 /*le: !T*/, 1 /*gt: T*/, 3 /*ge: 1*/};
            const int as_o_longer[4] = {1 /*lt: T*/, 3 /*le: 1*/, 0 /*gt*/, 2 /*ge: !T*/};
            const int kind = s_longer ? as_s_longer[op] : as_o_longer[op];
            by_len = kind == 0 ? g.trivial_block(0) : kind == 3 ? g.trivial_block(1) : kind == 1 ? T : g.not_flag(T);
        }
        // at_first implies any, so a = at_first + any is 0 (no difference: the lengths decide), 1 (first difference
        // against the op) or 2 (for it); x = a + 3 by_len keeps the weights, hence the noise, small
        std::array<uint8_t, 16> fin{};
        for (int v = 0; v < 6; v++) { const int a = v % 3, b = v / 3; fin[v] = (uint8_t)(a == 0 ? b : (a == 2)); }
        return g.flag_char(g.pbs({{at_first, 1}, {any, 1}, {by_len, 3}}, 0, fin));
    }
    const Char len1 = len(s), len2 = len(o);
    if (fast) {
        // ret = comparison at the first differing index; without one, the length comparison decides
        std::vector<Char> differs, cmp;
        for (size_t i = 0; i < n; i++) { differs.push_back(g.ne(s[i], o[i])); cmp.push_back(char_cmp(s[i], o[i], op)); }
        Char any;
        const std::vector<Char> first = g.first_one_hot(differs, &any);
        std::vector<Char> terms;
        for (size_t i = 0; i < n; i++) terms.push_back(g.flag_char(g.mul_flag(first[i][0], cmp[i][0])));
        const Char at_first = g.or_all(terms);
        const Char by_len = char_cmp(len1, len2, op);   // ge == (eq | gt), le == (eq | lt) on u8
        std::array<uint8_t, 16> t{};
        for (int v = 0; v < 16; v++) t[v] = (v & 2) ? (v & 1) : ((v >> 2) & 1);
        return g.flag_char(g.pbs({{at_first[0], 1}, {any[0], 2}, {by_len[0], 4}}, 0, t));
    }
    Char encountered = zero(), became_one = zero(), ret = g.trivial_char(255);
    for (size_t i = 0; i < n; i++) {
        const Char cmp = char_cmp(s[i], o[i], op);
        const Char is_ne = g.ne(s[i], o[i]);
        encountered = g.bitor_(encountered, is_ne);
        const Char flag = g.bitand_(encountered, g.flip(became_one));
        became_one = g.bitor_(became_one, flag);
        ret = g.if_then_else(flag, cmp, ret);
    }
    const Char substrings_equal = g.eq(ret, g.trivial_char(255));
    const Char l_eq = g.eq(len1, len2), l_gt = g.gt(len1, len2), l_lt = g.lt(len1, len2);
    Char length_based;
    switch (op) {
        case 3: length_based = g.bitor_(l_eq, l_gt); break;
        case 1: length_based = g.bitor_(l_eq, l_lt); break;
        case 2: length_based = l_gt; break;
        default: length_based = l_lt; break;
    }
    return g.if_then_else(substrings_equal, length_based, ret);
}

Str StringOps::concatenate(const Str& s, const Str& o) {
    Str r = s;
    r.insert(r.end(), o.begin(), o.end());
    return bubble_zeroes_right(r);
}

// ------------------------------------------------------------------------------------ trim.rs
Str StringOps::trim_end(const Str& s) {
    Str result(s.size(), zero());
    if (fast) {
        std::vector<BlockId> nb;
        for (auto& c : s) nb.push_back(is_not_blank(c)[0]);
        const std::vector<BlockId> stop = suffix_or(g, nb);
        for (size_t i = 0; i < s.size(); i++) result[i] = g.mul_flag_char(stop[i], s[i]);
        return result;
    }
    Char stop = zero();
    for (size_t ii = 0; ii < s.size(); ii++) {
        const size_t i = s.size() - 1 - ii;
        const Char is_not_zero = g.ne(s[i], zero());
        const Char is_not_ws = g.flip(g.is_whitespace(s[i]));
        stop = g.bitor_(stop, g.bitand_(is_not_ws, is_not_zero));
        result[i] = g.if_then_else(stop, s[i], zero());
    }
    return result;
}

Str StringOps::trim_start(const Str& s) {
    Str result(s.size(), zero());
    if (fast) {
        std::vector<BlockId> nb;
        for (auto& c : s) nb.push_back(is_not_blank(c)[0]);
        std::reverse(nb.begin(), nb.end());
        std::vector<BlockId> stop = suffix_or(g, nb);
        std::reverse(stop.begin(), stop.end());
        for (size_t i = 0; i < s.size(); i++) result[i] = g.mul_flag_char(stop[i], s[i]);
        return bubble_zeroes_right(result);
    }
    Char stop = zero();
    for (size_t i = 0; i < s.size(); i++) {
        const Char is_not_zero = g.ne(s[i], zero());
        const Char is_not_ws = g.flip(g.is_whitespace(s[i]));
        stop = g.bitor_(stop, g.bitand_(is_not_ws, is_not_zero));
        result[i] = g.if_then_else(stop, s[i], zero());
    }
    return bubble_zeroes_right(result);
}

Str StringOps::trim(const Str& s) { return trim_start(trim_end(s)); }

// ------------------------------------------------------------------------------------ split.rs
// Faithful mode records the reference's own op order (a serial scan over the string with L x L copy buffers; the
// graph folds what cannot change the value: trivial buffer indices, untouched buffers).  Fast mode records the
// plaintext-identical parallel form of the scan (split_scan_fast below, the fast branch of split_ascii_whitespace) and
// of the chains inside clear_pattern_from_result, on top of the fast replace / bubble_zeroes_right / starts_with.
Char StringOps::rsplit_pattern_matching(size_t i, const Str& s, const Str& pattern, Str& ignore) {
    Char found = one();
    if (pattern.empty()) {
        const Char is_pad = g.eq(s[i], zero());
        if (i >= 1) {
            const Char prev_non_pad = g.ne(s[i - 1], zero());
            const Char end_of_string = g.bitand_(prev_non_pad, is_pad);
            found = g.if_then_else(end_of_string, one(), zero());
            found = g.bitor_(found, g.if_then_else(is_pad, zero(), one()));
        } else {
            found = g.if_then_else(is_pad, zero(), one());
        }
    } else if (pattern.size() > s.size() || i + pattern.size() >= s.size()) {
        found = zero();
    } else {
        for (size_t j = 0; j < pattern.size(); j++) {
            found = g.bitand_(found, g.eq(s[i + j], pattern[j]));
            found = g.bitand_(found, ignore[i + j]);
        }
    }
    for (size_t j = 0; j < pattern.size(); j++)
        if (i + j < s.size()) ignore[i + j] = g.bitand_(ignore[i + j], g.if_then_else(found, zero(), one()));
    return found;
}

Char StringOps::split_pattern_matching(size_t i, const Str& s, const Str& pattern, Str& ignore) {
    Char found = one();
    if (pattern.size() > s.size() || (long long)i < (long long)pattern.size() - 1) {
        found = zero();
    } else {
        for (size_t j = 0; j < pattern.size(); j++) {
            const size_t k = i - pattern.size() + 1 + j;
            found = g.bitand_(found, g.eq(s[k], pattern[j]));
            found = g.bitand_(found, ignore[k]);
        }
    }
    for (size_t j = 0; j < pattern.size(); j++)
        if (i + j < s.size()) ignore[i + j] = g.bitand_(ignore[i + j], g.if_then_else(found, zero(), one()));
    return found;
}

void StringOps::copy_logic(size_t i, const Char* n, const Str& s, std::vector<Str>& result, const Char& allow, const Char& ccb) {
    for (size_t j = 0; j < s.size(); j++) {
        Char copy_flag = g.eq(g.trivial_char((uint8_t)j), ccb);
        if (n) copy_flag = g.bitand_(copy_flag, allow);
        result[j][i] = g.if_then_else(copy_flag, s[i], result[j][i]);
    }
}

void StringOps::handle_n_case(const Char& found, const Char* n, Char& ccb, Char& stop) {
    if (!n) {
        ccb = g.if_then_else(found, g.add(ccb, one()), ccb);
        return;
    }
    stop = g.bitor_(stop, g.eq(ccb, g.sub(*n, one())));
    ccb = g.if_then_else(g.bitand_(found, g.flip(stop)), g.add(ccb, one()), ccb);
}

void StringOps::clear_pattern_from_result(const Char* n, std::vector<Str>& result, const Str& pattern, bool inclusive, bool terminator) {
    const size_t size = result.size();
    const Str to(pattern.size(), zero());
    if (n && fast) {
        // stop_replacing after buffer i = OR_{k <= i} (n == k + 1): a prefix OR instead of a chain of size bitors
        std::vector<BlockId> hit;
        for (size_t i = 0; i < size; i++)
            hit.push_back(g.block_and_eq({{*n, g.trivial_char((uint8_t)(((i & 255) + 1) & 255))}})[0]);
        std::reverse(hit.begin(), hit.end());
        std::vector<BlockId> stop = suffix_or(g, hit);
        std::reverse(stop.begin(), stop.end());
        for (size_t i = 0; i < size; i++) {
            const Str current = bubble_zeroes_right(result[i]);
            const Str replacement = replace(current, pattern, to);
            const BlockId go = g.not_flag(stop[i]);
            for (size_t j = 0; j < size; j++)
                result[i][j] = g.add_disjoint({g.mul_flag_char(stop[i], current[j]), g.mul_flag_char(go, replacement[j])});
        }
        return;
    }
    if (n) {
        Char stop_replacing = zero();
        for (size_t i = 0; i < size; i++) {
            stop_replacing = g.bitor_(stop_replacing, g.eq(*n, g.add(g.trivial_char((uint8_t)i), one())));
            const Str current = bubble_zeroes_right(result[i]);
            const Str replacement = replace(current, pattern, to);
            for (size_t j = 0; j < size; j++) result[i][j] = g.if_then_else(stop_replacing, current[j], replacement[j]);
        }
        return;
    }
    if (!inclusive) {
        for (size_t i = 0; i < size; i++) result[i] = replace(result[i], pattern, to);
    } else {
        for (size_t i = 0; i < size; i++) result[i] = bubble_zeroes_right(result[i]);
    }
    if (terminator && fast) {
        // the last run of all-zero buffers that start with the pattern is deleted: non_zero_found before buffer i is the
        // OR over the LATER buffers, a suffix OR instead of a chain
        std::vector<BlockId> nz(size), sw(size);
        for (size_t i = 0; i < size; i++) {
            std::vector<Char> any;
            for (size_t j = 0; j < size; j++) any.push_back(g.nonzero(result[i][j]));
            nz[i] = g.or_all(any)[0];
            sw[i] = g.cond_bit(starts_with(result[i], pattern));
        }
        const std::vector<BlockId> later = suffix_or(g, nz);   // later[i] = OR_{k >= i} nz[k]
        for (size_t i = 0; i < size; i++) {
            std::vector<Char> all{g.flag_char(sw[i]), g.flag_char(g.not_flag(nz[i]))};
            if (i + 1 < size) all.push_back(g.flag_char(g.not_flag(later[i + 1])));
            const BlockId keep = g.not_flag(g.and_all(all)[0]);
            for (size_t j = 0; j < size; j++) result[i][j] = g.mul_flag_char(keep, result[i][j]);
        }
        return;
    }
    if (terminator) {
        Char non_zero_found = zero();
        for (size_t ii = 0; ii < size; ii++) {
            const size_t i = size - 1 - ii;
            Char is_buff_zero = one();
            for (size_t j = 0; j < size; j++) is_buff_zero = g.bitand_(is_buff_zero, g.eq(result[i][j], zero()));
            const Char sw = starts_with(result[i], pattern);
            const Char should_delete = g.bitand_(g.bitand_(sw, is_buff_zero), g.flip(non_zero_found));
            for (size_t j = 0; j < size; j++) result[i][j] = g.if_then_else(should_delete, zero(), result[i][j]);
            non_zero_found = g.bitor_(non_zero_found, g.flip(is_buff_zero));
        }
    }
}

// Column t of the split buffers: buffer j takes `src` iff the number of set flags in `seen` -- capped at *cap in the
// n forms -- equals j.  Up to 15 distinct flags and no (or a clear) cap: "count == j" is ONE PBS on the flag sum per
// buffer; otherwise the count is formed as a u8 (sum_flags), capped, and compared with j.
void StringOps::copy_to_counted_buffer(const std::vector<Char>& seen, const Char* cap, const Char& src, size_t t,
                                       std::vector<Str>& result) {
    const size_t size = result.size();
    std::vector<std::pair<BlockId, int>> live;
    for (auto& c : seen) if (!(g.is_trivial(c[0]) && g.trivial_value(c[0]) == 0)) live.push_back({c[0], 1});
    bool repeats = false;   // a shared node would be summed with a coefficient > 1
    for (size_t a = 0; a < live.size() && !repeats; a++)
        for (size_t b = a + 1; b < live.size(); b++) if (live[a].first == live[b].first) { repeats = true; break; }
    // a clear cap (rsplit_once: n = 2; the *_clear forms) keeps the one-PBS form: buffer j < cap takes count == j,
    // buffer cap takes count >= cap, later buffers nothing
    int clear_cap = -1;
    if (cap) {
        bool triv = true;
        for (int q = 0; q < 4; q++) triv = triv && g.is_trivial((*cap)[q]);
        if (triv) { clear_cap = 0; for (int q = 0; q < 4; q++) clear_cap |= (g.trivial_value((*cap)[q]) & 3) << (2 * q); }
    }
    if ((!cap || clear_cap >= 0) && live.size() <= 15 && !repeats) {
        for (size_t j = 0; j < size && j <= live.size(); j++) {
            if (clear_cap >= 0 && (int)j > clear_cap) break;
            std::array<uint8_t, 16> tab{};
            if (clear_cap >= 0 && (int)j == clear_cap) { for (int v = (int)j; v < 16; v++) tab[v] = 1; }
            else tab[j & 15] = 1;
            const BlockId hit = live.empty() ? g.trivial_block(j == 0 ? 1 : 0) : g.pbs(live, 0, tab);
            result[j][t] = g.mul_flag_char(hit, src);
        }
        return;
    }
    Char ccb = g.sum_flags(seen);
    if (cap) {
        const BlockId capped = g.cond_bit(g.ge(ccb, *cap));
        ccb = g.add_disjoint({g.mul_flag_char(capped, *cap), g.mul_flag_char(g.not_flag(capped), ccb)});
    }
    for (size_t j = 0; j < size; j++) {
        const BlockId hit = g.cond_bit(g.block_and_eq({{g.trivial_char((uint8_t)(j & 255)), ccb}}));
        result[j][t] = g.mul_flag_char(hit, src);
    }
}

// Depth-minimised scan of _split / _rsplit (split.rs:883-988, :307-393), plaintext-identical to the serial one:
//  * the raw window matches come from the ORIGINAL string and are independent;
//  * the `ignore` bookkeeping of the reference only ever blocks a match because of an EARLIER match whose marked range
//    [i', i' + P - 1] meets the window: for split (windows END at i) that is a match at i' in [i - 2P + 2, i - 1], for
//    rsplit (windows START at i, scanned downwards) one at i' in [i + 1, i + P - 1].  At most one match fits such a
//    range, so found_i = raw_i AND no match in the range is ONE PBS on 2 raw_i + (sum of the range): one level per
//    position instead of four;
//  * the copy buffer index before position t is the number of matches seen so far -- capped at n - 1 in the n forms,
//    where the reference stops incrementing once it reaches n - 1 -- and cell (j, t) takes s[t] iff it equals j.
SplitResult StringOps::split_scan_fast(const Str& s, const Str& pattern, const Char* n, bool reverse) {
    const size_t size = s.size(), P = pattern.size();
    const BlockId f0 = g.trivial_block(0), f1 = g.trivial_block(1);
    std::vector<BlockId> found(size, f0);
    auto is2 = [](int v) { return (int)(v == 2); };
    std::array<uint8_t, 16> is2_tab{}, nz_tab{};
    for (int v = 0; v < 16; v++) { is2_tab[v] = (uint8_t)is2(v); nz_tab[v] = (uint8_t)(v != 0); }
    // found_i from raw_i and the earlier matches at `range` (most recent first)
    auto gate = [&](BlockId raw, std::vector<BlockId> range) {
        if (g.is_trivial(raw) && g.trivial_value(raw) == 0) return f0;
        std::vector<std::pair<BlockId, int>> ops{{raw, 2}};
        const size_t direct = 12;   // 2 raw + 12 + 1 <= 15: the value-set analysis does not know that at most one flag of the range is set
        if (range.size() > direct + 1) {   // the far end of the range was known long ago: fold it into one flag
            std::vector<std::pair<BlockId, int>> far;
            std::vector<BlockId> far_flags;
            for (size_t k = direct; k < range.size(); k++) {
                far.push_back({range[k], 1});
                if (far.size() == 15) { far_flags.push_back(g.pbs(far, 0, nz_tab)); far.clear(); }
            }
            if (!far.empty()) far_flags.push_back(far.size() == 1 ? far[0].first : g.pbs(far, 0, nz_tab));
            range.resize(direct);
            std::vector<Char> ff;
            for (auto b : far_flags) ff.push_back(g.flag_char(b));
            range.push_back(g.or_all(ff)[0]);
        }
        for (auto b : range) ops.push_back({b, 1});
        return g.pbs(ops, 0, is2_tab);
    };
    if (!reverse) {
        for (size_t i = 0; i < size; i++) {
            if (P > size || i + 1 < P) continue;
            const BlockId raw = P == 0 ? f1 : g.cond_bit(match_at(s, i + 1 - P, pattern, false));
            std::vector<BlockId> range;
            for (size_t d = 1; d + 2 <= 2 * P && d <= i; d++) range.push_back(found[i - d]);   // i' in [i - 2P + 2, i - 1]
            found[i] = gate(raw, range);
        }
    } else {
        std::vector<BlockId> nzs;
        if (P == 0) for (auto& c : s) nzs.push_back(g.cond_bit(g.nonzero(c)));
        for (size_t ii = 0; ii < size; ii++) {
            const size_t i = size - 1 - ii;
            if (P == 0) {   // a pad right after the last char, or any non-pad char: nz_i OR nz_(i-1)
                found[i] = i >= 1 ? g.pbs({{nzs[i], 1}, {nzs[i - 1], 1}}, 0, nz_tab) : nzs[i];
                continue;
            }
            if (P > size || i + P >= size) continue;
            const BlockId raw = g.cond_bit(match_at(s, i, pattern, false));
            std::vector<BlockId> range;
            for (size_t d = 1; d + 1 <= P && i + d < size; d++) range.push_back(found[i + d]);       // i' in [i + 1, i + P - 1]
            found[i] = gate(raw, range);
        }
    }
    // copy buffer index before each position, in scan order
    std::vector<Char> seen;   // 0/1 chars counted so far
    if (!reverse && P == 0 && n) {
        const Char skip_first = g.bitand_(g.gt(*n, one()), g.le(*n, len(s)));   // split.rs:917-927
        seen.push_back(skip_first);
    }
    const BlockId allow = n ? g.cond_bit(g.nonzero(*n)) : f1;
    const Char cap = n ? g.sub(*n, one()) : zero();
    std::vector<Str> result(size, Str(size, zero()));
    for (size_t tt = 0; tt < size; tt++) {
        const size_t t = reverse ? size - 1 - tt : tt;
        const Char src = n ? g.mul_flag_char(allow, s[t]) : s[t];
        copy_to_counted_buffer(seen, n ? &cap : nullptr, src, t, result);
        seen.push_back(g.flag_char(found[t]));
    }
    std::vector<Char> any;
    for (auto b : found) any.push_back(g.flag_char(b));
    return SplitResult{result, g.or_all(any)};
}

SplitResult StringOps::rsplit_impl(const Str& s_in, const Str& pattern, bool inclusive, bool terminator, const Char* n) {
    Str s = s_in;
    s.push_back(zero());
    if (fast) {
        SplitResult r = split_scan_fast(s, pattern, n, true);
        clear_pattern_from_result(n, r.buffers, pattern, inclusive, terminator);
        return r;
    }
    const size_t size = s.size();
    Char ccb = zero(), stop = zero(), found_any = zero();
    std::vector<Str> result(size, Str(size, zero()));
    const Char allow = n ? g.ne(*n, zero()) : zero();
    Str ignore(size, one());
    for (size_t ii = 0; ii < size; ii++) {
        const size_t i = size - 1 - ii;
        copy_logic(i, n, s, result, allow, ccb);
        const Char found = rsplit_pattern_matching(i, s, pattern, ignore);
        found_any = g.bitor_(found_any, found);
        handle_n_case(found, n, ccb, stop);
    }
    clear_pattern_from_result(n, result, pattern, inclusive, terminator);
    return SplitResult{result, found_any};
}

SplitResult StringOps::split_impl(const Str& s_in, const Str& pattern, bool inclusive, bool terminator, const Char* n) {
    Str s = s_in;
    s.push_back(zero());
    if (fast) {
        SplitResult r = split_scan_fast(s, pattern, n, false);
        clear_pattern_from_result(n, r.buffers, pattern, inclusive, terminator);
        return r;
    }
    const size_t size = s.size();
    Char ccb = zero(), stop = zero(), found_any = zero();
    std::vector<Str> result(size, Str(size, zero()));
    const Char allow = n ? g.ne(*n, zero()) : zero();
    Str ignore(size, one());
    if (pattern.empty() && n) {
        const Char skip_first = g.bitand_(g.gt(*n, one()), g.le(*n, len(s)));
        ccb = g.if_then_else(skip_first, one(), ccb);
    }
    for (size_t i = 0; i < size; i++) {
        copy_logic(i, n, s, result, allow, ccb);
        const Char found = split_pattern_matching(i, s, pattern, ignore);
        found_any = g.bitor_(found_any, found);
        handle_n_case(found, n, ccb, stop);
    }
    clear_pattern_from_result(n, result, pattern, inclusive, terminator);
    return SplitResult{result, found_any};
}

SplitResult StringOps::split_ascii_whitespace(const Str& s) {
    const size_t size = s.size();
    if (fast) {
        // the buffer index at position i is the number of (whitespace after non-whitespace) transitions up to and
        // including i; only non-whitespace chars are ever copied, so the reference's second pass (whitespace -> NUL
        // inside the buffers, split.rs:1425-1434) cannot change a value and is not recorded
        std::vector<BlockId> ws(size), inc(size);
        std::array<uint8_t, 16> is1{};
        is1[1] = 1;
        for (size_t i = 0; i < size; i++) ws[i] = is_blank_not_nul(s[i])[0];
        for (size_t i = 0; i < size; i++)
            inc[i] = i == 0 ? g.trivial_block(0) : g.pbs({{ws[i], 1}, {ws[i - 1], 2}}, 0, is1);
        std::vector<Str> result(size, Str(size, zero()));
        std::vector<Char> seen;
        for (size_t i = 0; i < size; i++) {
            seen.push_back(g.flag_char(inc[i]));
            const Char src = g.mul_flag_char(g.not_flag(ws[i]), s[i]);
            copy_to_counted_buffer(seen, nullptr, src, i, result);
        }
        for (size_t j = 0; j < size; j++) result[j] = bubble_zeroes_right(result[j]);
        std::vector<Char> any;
        for (auto b : ws) any.push_back(g.flag_char(b));
        return SplitResult{result, g.or_all(any)};
    }
    Char ccb = zero(), prev_ws = one(), found_any = zero();
    std::vector<Str> result(size, Str(size, zero()));
    for (size_t i = 0; i < size; i++) {
        const Char found = g.is_whitespace(s[i]);
        found_any = g.bitor_(found_any, found);
        ccb = g.if_then_else(g.bitand_(found, g.flip(prev_ws)), g.add(ccb, one()), ccb);
        const Char not_ws = g.flip(g.is_whitespace(s[i]));
        for (size_t j = 0; j < size; j++) {
            const Char copy_flag = g.bitand_(g.eq(g.trivial_char((uint8_t)j), ccb), not_ws);
            result[j][i] = g.if_then_else(copy_flag, s[i], result[j][i]);
        }
        prev_ws = found;
    }
    for (size_t j = 0; j < size; j++)
        for (size_t k = 0; k < size; k++)
            result[j][k] = g.if_then_else(g.is_whitespace(result[j][k]), zero(), result[j][k]);
    for (size_t j = 0; j < size; j++) result[j] = bubble_zeroes_right(result[j]);
    return SplitResult{result, found_any};
}

}  // namespace fhestr
