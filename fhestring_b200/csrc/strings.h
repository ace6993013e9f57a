// strings.h -- the reference's MyServerKey string methods (/root/reference/src/server_key/mod.rs,
// trim.rs, /root/reference/src/utils.rs:28-46) recorded into a Graph.
//
// Two recordings of every method give the same plaintext for all inputs:
//   faithful  the reference's own sequence of per-char primitives, op by op (the graph then levelises it
//             and folds what cannot change the value);
//   fast      the same function re-associated for depth: AND/OR chains become sum + LUT trees, priority
//             select chains become one-hot encodes, u8 accumulations become column compression,
//             bubble_zeroes_right becomes a log-depth routing network.
// tests/test_graph_strings.py checks both against oracle/fhestring_plain.py.
#pragma once
#include <string>
#include <vector>

#include "graph.h"

namespace fhestr {

typedef std::vector<Char> Str;

struct StripResult {
    Str string;
    Char found;
};

struct SplitResult {           // FheSplit (/root/reference/src/ciphertext/fhesplit.rs:5-8)
    std::vector<Str> buffers;  // max_no_buffers padded strings of equal length
    Char found;
};

class StringOps {
public:
    StringOps(Graph& graph, bool fast_mode) : g(graph), fast(fast_mode) {}

    Str bubble_zeroes_right(const Str& s);                                  // utils.rs:28
    Str to_upper(const Str& s);                                             // mod.rs:65
    Str to_lower(const Str& s);                                             // mod.rs:110
    Char contains(const Str& s, const Str& needle);                         // mod.rs:151
    Char ends_with(const Str& s, const Str& needle);                        // mod.rs:241
    Char starts_with(const Str& s, const Str& pattern);                     // mod.rs:344
    Char is_empty(const Str& s);                                            // mod.rs:431
    Char len(const Str& s);                                                 // mod.rs:478
    Str repeat_clear(const Str& s, size_t repetitions);                     // mod.rs:517
    Str repeat(const Str& s, const Char& repetitions);                      // mod.rs:567
    Str replace(const Str& s, const Str& from, const Str& to);              // mod.rs:624
    Str replacen(const Str& s, const Str& from, const Str& to, const Char& n);  // mod.rs:1729
    bool rfind(const Str& s, const Str& pattern, Char& out);                // mod.rs:727 (false: "panic")
    bool find(const Str& s, const Str& pattern, Char& out);                 // mod.rs:1010
    Char eq(const Str& s, const Str& o);                                    // mod.rs:1122
    Char ne(const Str& s, const Str& o);                                    // mod.rs:1178
    Char eq_ignore_case(const Str& s, const Str& o);                        // mod.rs:1221
    StripResult strip_prefix(const Str& s, const Str& pattern);             // mod.rs:1261
    StripResult strip_suffix(const Str& s, const Str& needle);              // mod.rs:1335
    Char comparison(const Str& s, const Str& o, int op);                    // mod.rs:1470 (0 lt 1 le 2 gt 3 ge)
    Str concatenate(const Str& s, const Str& o);                            // mod.rs:1864
    Str trim_end(const Str& s);                                             // trim.rs:36
    Str trim_start(const Str& s);                                           // trim.rs:86
    Str trim(const Str& s);                                                 // trim.rs:146
    // split family (/root/reference/src/server_key/split.rs); n == nullptr is the reference's Option::None
    SplitResult rsplit_impl(const Str& s, const Str& pattern, bool inclusive, bool terminator, const Char* n);  // :307
    SplitResult split_impl(const Str& s, const Str& pattern, bool inclusive, bool terminator, const Char* n);   // :883
    SplitResult split_ascii_whitespace(const Str& s);                                                          // :1377

    std::string error;  // set when a method hits one of the reference's panics

private:
    Graph& g;
    bool fast;
    Char zero() { return g.trivial_char(0); }
    Char one() { return g.trivial_char(1); }
    Char match_at(const Str& s, size_t i, const Str& pattern, bool reversed);
    Char char_cmp(const Char& a, const Char& b, int op);
    Str handle_longer_from(const Str& bytes, const Str& from, Str to, const Char& n, bool use_counter);
    Str handle_shorter_from(const Str& bytes, const Str& from, const Str& to, const Char& n, bool use_counter);
    std::vector<Char> last_one_hot(const std::vector<Char>& flags, Char* any);
    Char first_index_fast(const std::vector<BlockId>& flags, bool last);   // find / rfind over 16..255 windows, 8 levels
    Char is_not_blank(const Char& c);
    Char is_blank_not_nul(const Char& c);
    Char rsplit_pattern_matching(size_t i, const Str& s, const Str& pattern, Str& ignore);   // split.rs:10
    Char split_pattern_matching(size_t i, const Str& s, const Str& pattern, Str& ignore);    // split.rs:70
    void copy_logic(size_t i, const Char* n, const Str& s, std::vector<Str>& result, const Char& allow, const Char& ccb);  // :108
    void handle_n_case(const Char& found, const Char* n, Char& ccb, Char& stop);              // split.rs:137
    void copy_to_counted_buffer(const std::vector<Char>& seen, const Char* cap, const Char& src, size_t t, std::vector<Str>& result);
    SplitResult split_scan_fast(const Str& s, const Str& pattern, const Char* n, bool reverse);   // parallel form of the scans of :307 / :883
    void clear_pattern_from_result(const Char* n, std::vector<Str>& result, const Str& pattern, bool inclusive, bool terminator);  // :180
};

}  // namespace fhestr
