// graph.h -- host-side op graph: the 11 per-character primitives of the reference
// (/root/reference/src/ciphertext/fheasciichar.rs:17-168) recorded as radix-block PBS jobs, then
// flattened into dependency levels of independent jobs for the batched engine.
//
// This replaces what tfhe-rs' integer::ServerKey does behind those primitives: the
// per-op recipes over 4 blocks of 2 message + 2 carry bits, expressed with 16-entry LUTs over
// `sum coeff*block + const`.  It is pure host code (no CUDA): tests interpret the compiled job list on
// plaintext block values, the engine executes the same list on ciphertexts.
//
// Every block carries, on the host only:
//   vset    the set of plaintext values (mod 32, bit 4 = padding bit) the block can hold.  A PBS whose
//           output set is a single value is folded to a trivial constant (this subsumes tfhe-rs'
//           trivial-ciphertext short cut and constant propagation through flag logic);
//   noise2  squared 2-norm of the block as a combination of fresh PBS outputs.  The largest value the
//           reference's own recipes reach is 34 (comparison: raw subtraction of two packed pairs
//           (lo + 4 hi) - (lo' + 4 hi')); a refresh PBS is inserted before that would be exceeded;
//   level   PBS depth at which the value exists.
#pragma once
#include <stdint.h>

#include <array>
#include <functional>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/fhestr_engine.h"

namespace fhestr {

typedef uint32_t BlockId;
constexpr BlockId kNoBlock = 0xffffffffu;
typedef std::array<BlockId, 4> Char;  // little-endian base-4 digits of a u8

struct Term {
    BlockId blk;
    int32_t coeff;
};

enum class BKind : uint8_t { Trivial, Input, Pbs, Linear };

struct BlockNode {
    BKind kind = BKind::Trivial;
    bool materialized = false;   // Linear only: has its own arena slot (acts as an atom)
    uint32_t vset = 1;           // bit v set <=> value v (mod 32) possible
    float noise2 = 0.f;
    int32_t level = 0;
    int32_t lut = -1;            // Pbs: graph-local LUT id
    std::vector<Term> terms;     // Pbs input / Linear definition, over atoms only
    int32_t cst = 0;             // constant in units of delta (mod 32)
    int32_t slot = -1;           // arena slot, assigned by compile() (inputs: at creation)
    bool done = false;           // value exists in the arena (input, or computed and committed)
};

struct CompiledProgram {
    std::vector<fhestr_job> jobs;          // level by level; PBS jobs before leveled jobs inside a level
    std::vector<uint32_t> level_offsets;   // n_levels + 1
    std::vector<uint32_t> level_pbs;       // PBS jobs per level
    std::vector<uint32_t> level_first_dst; // first arena slot written by the PBS jobs of a level (contiguous)
    uint32_t n_slots = 0;
    uint32_t n_inputs = 0;                 // input blocks occupy slots [0, n_inputs)
    uint64_t n_pbs = 0;
    std::vector<std::pair<uint32_t, uint8_t>> trivial_slots;  // (slot, value) to initialise
};

class Graph {
public:
    static constexpr float kNoise2Limit = 34.0f;
    int delta_log = 59;

    // ---- block level
    BlockId trivial_block(int value);            // value mod 32
    BlockId input_block();
    BlockId lin(const std::vector<std::pair<BlockId, int>>& ops, int cst, uint32_t declared_vset = 0);
    BlockId pbs(const std::vector<std::pair<BlockId, int>>& ops, int cst, const std::array<uint8_t, 16>& table);
    BlockId bivar(BlockId x, BlockId y, const std::function<int(int, int)>& f);  // f over 2-bit x 2-bit... see .cpp
    BlockId refresh(BlockId b);
    bool is_trivial(BlockId b) const { return nodes[b].kind == BKind::Trivial; }
    int trivial_value(BlockId b) const;
    const BlockNode& node(BlockId b) const { return nodes[b]; }

    // ---- the reference's primitives (fheasciichar.rs)
    Char input_char();
    Char trivial_char(uint8_t v);                     // :17-25
    Char eq(const Char& a, const Char& b);            // :35
    Char ne(const Char& a, const Char& b);            // :40
    Char le(const Char& a, const Char& b);            // :45
    Char lt(const Char& a, const Char& b);            // :50
    Char ge(const Char& a, const Char& b);            // :55
    Char gt(const Char& a, const Char& b);            // :60
    Char bitand_(const Char& a, const Char& b);       // :65
    Char bitor_(const Char& a, const Char& b);        // :74
    Char sub(const Char& a, const Char& b);           // :83
    Char add(const Char& a, const Char& b);           // :87
    Char if_then_else(const Char& c, const Char& t, const Char& f);  // :93
    Char is_whitespace(const Char& a);                // :106
    Char is_uppercase(const Char& a);                 // :132
    Char is_lowercase(const Char& a);                 // :146
    Char flip(const Char& a);                         // :161
    // plaintext-equivalent wide reductions used by the depth-minimised string algorithms
    Char and_all(const std::vector<Char>& flags);     // AND of 0/1 chars, tree of sum + threshold LUTs (fan-in 16)
    Char or_all(const std::vector<Char>& flags);      // OR  of 0/1 chars, tree of sum + threshold LUTs (fan-in 16)
    BlockId and_flags(std::vector<BlockId> flags);    // the same on 0/1 blocks
    BlockId or_flags(std::vector<BlockId> flags);
    Char sum_flags(const std::vector<Char>& flags);   // u8 sum (mod 256) of 0/1 chars
    Char nonzero(const Char& a);                      // a != 0 as one PBS over the block sum
    Char block_and_eq(const std::vector<std::pair<Char, Char>>& pairs);  // AND_i (a_i == b_i), nibble level
    BlockId nibble_eq(BlockId a_lo, BlockId a_hi, BlockId b_lo, BlockId b_hi);  // (a_lo + 4 a_hi) == (b_lo + 4 b_hi)
    std::vector<BlockId> nibble_eq_flags(const std::vector<std::pair<Char, Char>>& pairs);  // two flags per char pair
    Char or_of_ands(const std::vector<std::vector<BlockId>>& windows);  // OR_w AND_i flags[w][i], depth-minimised
    // generalisation of sum_flags: exact sum of 0/1 blocks as `ndigits` clean base-4 digits (little endian)
    std::vector<BlockId> sum_digits(const std::vector<BlockId>& flags, int ndigits);
    // (a != b, a strictly greater (greater = true) or smaller than b) as two 0/1 blocks off the same packed-pair
    // comparison the reference's recipe uses (cmp): 2 + 2 PBS per char pair
    std::pair<BlockId, BlockId> differs_and_strict(const Char& a, const Char& b, bool greater);
    BlockId cond_bit(const Char& c);                  // c != 0 as a 0/1 block (scalar_ne(c, 0))
    Char flag_char(BlockId b);                        // BooleanBlock::into_radix: [b, 0, 0, 0]
    BlockId not_flag(BlockId b);                      // 1 - b, leveled
    BlockId mul_flag(BlockId flag, BlockId blk);      // flag ? blk : 0 for a clean 2-bit blk
    Char mul_flag_char(BlockId flag, const Char& c);
    Char add_disjoint(const std::vector<Char>& parts); // block-wise sum of chars of which at most one is non-zero
    // stable compaction: non-NUL chars first, order kept (== bubble_zeroes_right in value), as a
    // log-depth routing network driven by the per-char count of preceding NULs
    std::vector<Char> compact_nonzero(const std::vector<Char>& s);
    // one-hot of the first set flag (all zero if none) -- replaces priority-select chains
    std::vector<Char> first_one_hot(const std::vector<Char>& flags, Char* any);
    // u8 value sum_i onehot_i * value_i + (none ? none_value : 0), refreshed to clean blocks
    Char select_by_one_hot(const std::vector<Char>& onehot, const std::vector<uint8_t>& values, const Char& none, uint8_t none_value);

    // ---- compile
    void mark_output(const Char& c) { for (auto b : c) outputs.push_back(b); }
    // returns false and fills error on failure
    bool compile(CompiledProgram& out, std::string& error, uint32_t slot_align = 1);
    void commit();               // the compiled program has run: its results are level-0 atoms from now on
    uint32_t slots_used() const { return next_slot; }
    void reserve_slots(uint32_t first_free) { if (first_free > next_slot) next_slot = first_free; }
    int slot_of(BlockId b) const { return nodes[b].slot; }
    const std::vector<std::array<uint8_t, 16>>& luts() const { return lut_tables; }
    uint64_t pbs_recorded() const { return n_pbs_nodes; }

    std::string error;  // sticky: first recording error (bad value range, ...)

private:
    std::vector<BlockNode> nodes;
    std::vector<BlockId> outputs;
    std::vector<std::array<uint8_t, 16>> lut_tables;
    std::map<std::array<uint8_t, 16>, int> lut_ids;
    // common-subexpression table of the PBS nodes: open addressing over node ids, keyed by a hash of (lut, constant,
    // terms) and compared against the node itself -- no key strings (they were 40 % of the recording time of a
    // 138 k-PBS graph)
    std::vector<uint64_t> cse_tab;   // (lower hash half << 32) | node id; all ones = free
    size_t cse_count = 0;
    static uint64_t cse_hash(int lut, int cst, const std::vector<Term>& terms);
    BlockId cse_find(uint64_t h, int lut, int cst, const std::vector<Term>& terms) const;
    void cse_insert(uint64_t h, BlockId id);
    BlockId triv_cache[32];
    bool triv_cache_init = false;
    uint64_t n_pbs_nodes = 0;
    uint32_t next_slot = 0;
    uint32_t n_input_nodes = 0;
    std::vector<BlockId> pending;  // nodes given a slot by the last compile()
    std::vector<std::pair<BlockId, int>> flat_scratch;   // flatten(): merged (block, coefficient) pairs

    int lut_id(const std::array<uint8_t, 16>& t);
    void flatten(const std::vector<std::pair<BlockId, int>>& ops, int cst, std::vector<Term>& terms, int& out_cst);
    uint32_t vset_of(const std::vector<std::pair<BlockId, int>>& ops, int cst) const;
    float noise2_of(const std::vector<Term>& terms) const;
    int level_of(const std::vector<Term>& terms, bool for_linear) const;
    bool signed_ok = false;
    void fail(const std::string& m) { if (error.empty()) error = m; }
    Char cmp(const Char& a, const Char& b, int op);
    static constexpr size_t kLinearRoutingMinLength = 64;   // longer strings: half the PBS per routing layer, more levels
    std::vector<Char> route_linear(const std::vector<Char>& s, std::vector<std::vector<BlockId>> ctrl, int B, int nd);
    std::array<BlockId, 4> propagate(std::array<BlockId, 4> s);
};

}  // namespace fhestr
