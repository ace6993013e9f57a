// graph_capi.cpp -- C ABI of the op graph (include/fhestr_engine.h, "op graph" section): char ids in,
// char ids out, compile to dependency levels, run on the engine.  Host code only; the device work is
// what fhestr_program_run launches.
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "graph.h"
#include "strings.h"

using namespace fhestr;

struct fhestr_graph {
    Graph g;
    std::vector<Char> chars;
    CompiledProgram prog;
    bool compiled = false;
    std::string err;
    // graph-local LUT id -> engine LUT id, for the engine the graph was last bound to
    fhestr_engine* bound = nullptr;
    std::vector<int32_t> engine_lut;
};

static int gfail(fhestr_graph* g, int code, const std::string& m) {
    g->err = m;
    return code;
}

static bool valid_ids(const fhestr_graph* g, const uint32_t* ids, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) if (ids[i] >= g->chars.size()) return false;
    return true;
}

static uint32_t push_char(fhestr_graph* g, const Char& c) {
    g->chars.push_back(c);
    return (uint32_t)g->chars.size() - 1;
}

#pragma GCC visibility push(default)
extern "C" {
int fhestr_graph_bind(fhestr_graph* g, fhestr_engine* e, fhestr_program** out);

int fhestr_graph_create(int32_t delta_log, fhestr_graph** out) {
    if (!out || delta_log != 59) return FHESTR_E_INVALID;  // the recipes are written for 2 message + 2 carry bits
    fhestr_graph* g = new (std::nothrow) fhestr_graph();
    if (!g) return FHESTR_E_STATE;
    g->g.delta_log = delta_log;
    *out = g;
    return FHESTR_OK;
}

void fhestr_graph_destroy(fhestr_graph* g) { delete g; }

const char* fhestr_graph_last_error(const fhestr_graph* g) { return g ? g->err.c_str() : "null graph"; }

int fhestr_graph_input_chars(fhestr_graph* g, uint32_t count, uint32_t* ids, uint32_t* slots) {
    if (!g || !ids) return FHESTR_E_INVALID;
    for (uint32_t i = 0; i < count; i++) {
        const Char c = g->g.input_char();
        ids[i] = push_char(g, c);
        if (slots) for (int b = 0; b < 4; b++) slots[4 * i + b] = (uint32_t)g->g.slot_of(c[b]);
    }
    return FHESTR_OK;
}

int fhestr_graph_trivial_chars(fhestr_graph* g, const uint8_t* values, uint32_t count, uint32_t* ids) {
    if (!g || !values || !ids) return FHESTR_E_INVALID;
    for (uint32_t i = 0; i < count; i++) ids[i] = push_char(g, g->g.trivial_char(values[i]));
    return FHESTR_OK;
}

static int graph_char_op_impl(fhestr_graph* g, int op, uint32_t a, uint32_t b, uint32_t c, uint32_t* out) {
    if (!g || !out) return FHESTR_E_INVALID;
    const bool unary = op >= FHESTR_OP_IS_WHITESPACE;
    const bool ternary = op == FHESTR_OP_IF_THEN_ELSE;
    if (a >= g->chars.size() || (!unary && b >= g->chars.size()) || (ternary && c >= g->chars.size()))
        return gfail(g, FHESTR_E_INVALID, "char id out of range");
    const Char A = g->chars[a];
    const Char B = unary ? A : g->chars[b];
    Char r;
    switch (op) {
        case FHESTR_OP_EQ: r = g->g.eq(A, B); break;
        case FHESTR_OP_NE: r = g->g.ne(A, B); break;
        case FHESTR_OP_LE: r = g->g.le(A, B); break;
        case FHESTR_OP_LT: r = g->g.lt(A, B); break;
        case FHESTR_OP_GE: r = g->g.ge(A, B); break;
        case FHESTR_OP_GT: r = g->g.gt(A, B); break;
        case FHESTR_OP_BITAND: r = g->g.bitand_(A, B); break;
        case FHESTR_OP_BITOR: r = g->g.bitor_(A, B); break;
        case FHESTR_OP_SUB: r = g->g.sub(A, B); break;
        case FHESTR_OP_ADD: r = g->g.add(A, B); break;
        case FHESTR_OP_IF_THEN_ELSE: r = g->g.if_then_else(A, B, g->chars[c]); break;
        case FHESTR_OP_IS_WHITESPACE: r = g->g.is_whitespace(A); break;
        case FHESTR_OP_IS_UPPERCASE: r = g->g.is_uppercase(A); break;
        case FHESTR_OP_IS_LOWERCASE: r = g->g.is_lowercase(A); break;
        case FHESTR_OP_FLIP: r = g->g.flip(A); break;
        default: return gfail(g, FHESTR_E_INVALID, "unknown char op");
    }
    if (!g->g.error.empty()) return gfail(g, FHESTR_E_INVALID, g->g.error);
    *out = push_char(g, r);
    return FHESTR_OK;
}

static int graph_string_op_impl(fhestr_graph* g, int method, int fast, const fhestr_str_arg* args, uint32_t n_args,
                           uint64_t clear_n, uint32_t* out_chars, uint32_t out_cap, uint32_t* out_len,
                           uint32_t* out_char) {
    if (!g || (!args && n_args)) return FHESTR_E_INVALID;
    std::vector<Str> A(n_args);
    for (uint32_t i = 0; i < n_args; i++) {
        if (args[i].len && !args[i].chars) return gfail(g, FHESTR_E_INVALID, "null string argument");
        if (!valid_ids(g, args[i].chars, args[i].len)) return gfail(g, FHESTR_E_INVALID, "char id out of range");
        for (uint32_t j = 0; j < args[i].len; j++) A[i].push_back(g->chars[args[i].chars[j]]);
    }
    auto need = [&](uint32_t n) { return n_args == n; };
    auto need_char = [&](uint32_t i) { return i < n_args && A[i].size() == 1; };
    StringOps ops(g->g, fast != 0);
    bool has_str = false, has_char = false;
    Str rs;
    Char rc{};
    bool ok = true;
    switch (method) {
        case FHESTR_M_CONTAINS: ok = need(2); if (ok) { rc = ops.contains(A[0], A[1]); has_char = true; } break;
        case FHESTR_M_ENDS_WITH: ok = need(2); if (ok) { rc = ops.ends_with(A[0], A[1]); has_char = true; } break;
        case FHESTR_M_STARTS_WITH: ok = need(2); if (ok) { rc = ops.starts_with(A[0], A[1]); has_char = true; } break;
        case FHESTR_M_IS_EMPTY: ok = need(1); if (ok) { rc = ops.is_empty(A[0]); has_char = true; } break;
        case FHESTR_M_LEN: ok = need(1); if (ok) { rc = ops.len(A[0]); has_char = true; } break;
        case FHESTR_M_REPEAT_CLEAR: ok = need(1); if (ok) { rs = ops.repeat_clear(A[0], (size_t)clear_n); has_str = true; } break;
        case FHESTR_M_REPEAT: ok = need(2) && need_char(1); if (ok) { rs = ops.repeat(A[0], A[1][0]); has_str = true; } break;
        case FHESTR_M_REPLACE: ok = need(3); if (ok) { rs = ops.replace(A[0], A[1], A[2]); has_str = true; } break;
        case FHESTR_M_REPLACEN: ok = need(4) && need_char(3); if (ok) { rs = ops.replacen(A[0], A[1], A[2], A[3][0]); has_str = true; } break;
        case FHESTR_M_RFIND: ok = need(2); if (ok) { if (!ops.rfind(A[0], A[1], rc)) return gfail(g, FHESTR_E_INVALID, ops.error); has_char = true; } break;
        case FHESTR_M_FIND: ok = need(2); if (ok) { if (!ops.find(A[0], A[1], rc)) return gfail(g, FHESTR_E_INVALID, ops.error); has_char = true; } break;
        case FHESTR_M_EQ: ok = need(2); if (ok) { rc = ops.eq(A[0], A[1]); has_char = true; } break;
        case FHESTR_M_NE: ok = need(2); if (ok) { rc = ops.ne(A[0], A[1]); has_char = true; } break;
        case FHESTR_M_EQ_IGNORE_CASE: ok = need(2); if (ok) { rc = ops.eq_ignore_case(A[0], A[1]); has_char = true; } break;
        case FHESTR_M_STRIP_PREFIX: ok = need(2); if (ok) { StripResult r = ops.strip_prefix(A[0], A[1]); rs = r.string; rc = r.found; has_str = has_char = true; } break;
        case FHESTR_M_STRIP_SUFFIX: ok = need(2); if (ok) { StripResult r = ops.strip_suffix(A[0], A[1]); rs = r.string; rc = r.found; has_str = has_char = true; } break;
        case FHESTR_M_LT: ok = need(2); if (ok) { rc = ops.comparison(A[0], A[1], 0); has_char = true; } break;
        case FHESTR_M_LE: ok = need(2); if (ok) { rc = ops.comparison(A[0], A[1], 1); has_char = true; } break;
        case FHESTR_M_GT: ok = need(2); if (ok) { rc = ops.comparison(A[0], A[1], 2); has_char = true; } break;
        case FHESTR_M_GE: ok = need(2); if (ok) { rc = ops.comparison(A[0], A[1], 3); has_char = true; } break;
        case FHESTR_M_CONCATENATE: ok = need(2); if (ok) { rs = ops.concatenate(A[0], A[1]); has_str = true; } break;
        case FHESTR_M_TO_UPPER: ok = need(1); if (ok) { rs = ops.to_upper(A[0]); has_str = true; } break;
        case FHESTR_M_TO_LOWER: ok = need(1); if (ok) { rs = ops.to_lower(A[0]); has_str = true; } break;
        case FHESTR_M_TRIM_END: ok = need(1); if (ok) { rs = ops.trim_end(A[0]); has_str = true; } break;
        case FHESTR_M_TRIM_START: ok = need(1); if (ok) { rs = ops.trim_start(A[0]); has_str = true; } break;
        case FHESTR_M_TRIM: ok = need(1); if (ok) { rs = ops.trim(A[0]); has_str = true; } break;
        case FHESTR_M_BUBBLE_ZEROES_RIGHT: ok = need(1); if (ok) { rs = ops.bubble_zeroes_right(A[0]); has_str = true; } break;
        default: return gfail(g, FHESTR_E_INVALID, "unknown string method");
    }
    if (!ok) return gfail(g, FHESTR_E_INVALID, "wrong arguments for this string method");
    if (!g->g.error.empty()) return gfail(g, FHESTR_E_INVALID, g->g.error);
    if (has_str) {
        if (!out_chars || !out_len) return gfail(g, FHESTR_E_INVALID, "method returns a string: out_chars/out_len needed");
        *out_len = (uint32_t)rs.size();
        if (rs.size() > out_cap) return gfail(g, FHESTR_E_INVALID, "result string does not fit out_cap");
        for (size_t i = 0; i < rs.size(); i++) out_chars[i] = push_char(g, rs[i]);
    } else if (out_len) {
        *out_len = 0;
    }
    if (has_char) {
        if (!out_char) return gfail(g, FHESTR_E_INVALID, "method returns a char: out_char needed");
        *out_char = push_char(g, rc);
    }
    return FHESTR_OK;
}

static int graph_split_op_impl(fhestr_graph* g, int method, int fast, const fhestr_str_arg* args, uint32_t n_args,
                          uint32_t* out_chars, uint32_t out_cap, uint32_t* n_buffers, uint32_t* buffer_len,
                          uint32_t* out_found) {
    if (!g || (!args && n_args) || !out_chars || !n_buffers || !buffer_len || !out_found) return FHESTR_E_INVALID;
    std::vector<Str> A(n_args);
    for (uint32_t i = 0; i < n_args; i++) {
        if (args[i].len && !args[i].chars) return gfail(g, FHESTR_E_INVALID, "null string argument");
        if (!valid_ids(g, args[i].chars, args[i].len)) return gfail(g, FHESTR_E_INVALID, "char id out of range");
        for (uint32_t j = 0; j < args[i].len; j++) A[i].push_back(g->chars[args[i].chars[j]]);
    }
    const bool wants_n = method == FHESTR_S_SPLITN || method == FHESTR_S_RSPLITN;
    const uint32_t need = method == FHESTR_S_SPLIT_ASCII_WHITESPACE ? 1u : (wants_n ? 3u : 2u);
    if (n_args != need || (wants_n && A[2].size() != 1)) return gfail(g, FHESTR_E_INVALID, "wrong arguments for this split method");
    StringOps ops(g->g, fast != 0);
    SplitResult r;
    const Char two = g->g.trivial_char(2);
    switch (method) {
        case FHESTR_S_SPLIT: r = ops.split_impl(A[0], A[1], false, false, nullptr); break;
        case FHESTR_S_RSPLIT: r = ops.rsplit_impl(A[0], A[1], false, false, nullptr); break;
        case FHESTR_S_SPLIT_INCLUSIVE: r = ops.split_impl(A[0], A[1], true, false, nullptr); break;
        case FHESTR_S_SPLIT_TERMINATOR: r = ops.split_impl(A[0], A[1], false, true, nullptr); break;
        case FHESTR_S_RSPLIT_TERMINATOR: r = ops.rsplit_impl(A[0], A[1], false, true, nullptr); break;
        case FHESTR_S_RSPLIT_ONCE: r = ops.rsplit_impl(A[0], A[1], false, false, &two); break;
        case FHESTR_S_SPLITN: r = ops.split_impl(A[0], A[1], false, false, &A[2][0]); break;
        case FHESTR_S_RSPLITN: r = ops.rsplit_impl(A[0], A[1], false, false, &A[2][0]); break;
        case FHESTR_S_SPLIT_ASCII_WHITESPACE: r = ops.split_ascii_whitespace(A[0]); break;
        default: return gfail(g, FHESTR_E_INVALID, "unknown split method");
    }
    if (!g->g.error.empty()) return gfail(g, FHESTR_E_INVALID, g->g.error);
    *n_buffers = (uint32_t)r.buffers.size();
    *buffer_len = r.buffers.empty() ? 0u : (uint32_t)r.buffers[0].size();
    if ((uint64_t)*n_buffers * *buffer_len > out_cap) return gfail(g, FHESTR_E_INVALID, "split result does not fit out_cap");
    for (size_t b = 0; b < r.buffers.size(); b++)
        for (size_t i = 0; i < r.buffers[b].size(); i++) out_chars[b * *buffer_len + i] = push_char(g, r.buffers[b][i]);
    *out_found = push_char(g, r.found);
    return FHESTR_OK;
}

int fhestr_graph_mark_output(fhestr_graph* g, const uint32_t* ids, uint32_t count) {
    if (!g || (!ids && count)) return FHESTR_E_INVALID;
    if (!valid_ids(g, ids, count)) return gfail(g, FHESTR_E_INVALID, "char id out of range");
    for (uint32_t i = 0; i < count; i++) g->g.mark_output(g->chars[ids[i]]);
    return FHESTR_OK;
}

static int graph_compile_impl(fhestr_graph* g, uint32_t slot_align, fhestr_graph_info* info) {
    if (!g) return FHESTR_E_INVALID;
    std::string err;
    if (!g->g.compile(g->prog, err, slot_align)) return gfail(g, FHESTR_E_INVALID, err);
    g->compiled = true;
    if (info) {
        info->n_levels = (uint32_t)g->prog.level_pbs.size();
        info->n_jobs = (uint32_t)g->prog.jobs.size();
        info->n_luts = (uint32_t)g->g.luts().size();
        info->n_trivial = (uint32_t)g->prog.trivial_slots.size();
        info->slots_used = g->prog.n_slots;
        info->n_pbs = g->prog.n_pbs;
        info->n_pbs_recorded = g->g.pbs_recorded();
    }
    return FHESTR_OK;
}

int fhestr_graph_get_program(const fhestr_graph* g, fhestr_job* jobs, uint32_t* level_offsets, uint32_t* level_pbs,
                             uint32_t* level_first_dst) {
    if (!g || !g->compiled) return FHESTR_E_STATE;
    const CompiledProgram& p = g->prog;
    if (jobs && !p.jobs.empty()) memcpy(jobs, p.jobs.data(), p.jobs.size() * sizeof(fhestr_job));
    if (level_offsets) memcpy(level_offsets, p.level_offsets.data(), p.level_offsets.size() * sizeof(uint32_t));
    if (level_pbs && !p.level_pbs.empty()) memcpy(level_pbs, p.level_pbs.data(), p.level_pbs.size() * sizeof(uint32_t));
    if (level_first_dst && !p.level_first_dst.empty())
        memcpy(level_first_dst, p.level_first_dst.data(), p.level_first_dst.size() * sizeof(uint32_t));
    return FHESTR_OK;
}

int fhestr_graph_get_luts(const fhestr_graph* g, uint8_t* tables) {
    if (!g || !tables) return FHESTR_E_INVALID;
    const auto& L = g->g.luts();
    for (size_t i = 0; i < L.size(); i++) memcpy(tables + 16 * i, L[i].data(), 16);
    return FHESTR_OK;
}

int fhestr_graph_get_trivials(const fhestr_graph* g, uint32_t* slots, uint8_t* values) {
    if (!g || !g->compiled) return FHESTR_E_STATE;
    for (size_t i = 0; i < g->prog.trivial_slots.size(); i++) {
        if (slots) slots[i] = g->prog.trivial_slots[i].first;
        if (values) values[i] = g->prog.trivial_slots[i].second;
    }
    return FHESTR_OK;
}

int fhestr_graph_char_slots(const fhestr_graph* g, const uint32_t* ids, uint32_t count, uint32_t* slots) {
    if (!g || !ids || !slots) return FHESTR_E_INVALID;
    if (!valid_ids(g, ids, count)) return FHESTR_E_INVALID;
    for (uint32_t i = 0; i < count; i++)
        for (int b = 0; b < 4; b++) {
            const int s = g->g.slot_of(g->chars[ids[i]][b]);
            if (s < 0) return FHESTR_E_STATE;  // not an output of a compile
            slots[4 * i + b] = (uint32_t)s;
        }
    return FHESTR_OK;
}

int fhestr_graph_reserve_slots(fhestr_graph* g, uint32_t first_free) {
    if (!g) return FHESTR_E_INVALID;
    g->g.reserve_slots(first_free);
    return FHESTR_OK;
}

static int graph_bind_impl(fhestr_graph* g, fhestr_engine* e, fhestr_program** out) {
    if (!g || !e || !out) return FHESTR_E_INVALID;
    if (!g->compiled) return gfail(g, FHESTR_E_STATE, "graph not compiled");
    if (g->bound != e) { g->bound = e; g->engine_lut.clear(); }
    const auto& L = g->g.luts();
    while (g->engine_lut.size() < L.size()) {
        int32_t id = -1;
        const int rc = fhestr_lut_register(e, L[g->engine_lut.size()].data(), &id);
        if (rc) return gfail(g, rc, std::string("lut_register: ") + fhestr_last_error(e));
        g->engine_lut.push_back(id);
    }
    for (auto& tv : g->prog.trivial_slots) {
        const int rc = fhestr_ct_trivial(e, tv.first, 1, &tv.second);
        if (rc) return gfail(g, rc, std::string("ct_trivial: ") + fhestr_last_error(e));
    }
    std::vector<fhestr_job> jobs = g->prog.jobs;
    for (auto& j : jobs) if (j.lut >= 0) j.lut = g->engine_lut[j.lut];
    const int rc = fhestr_program_create(e, jobs.data(), g->prog.level_offsets.data(),
                                         (uint32_t)g->prog.level_pbs.size(), out);
    if (rc) return gfail(g, rc, std::string("program_create: ") + fhestr_last_error(e));
    return FHESTR_OK;
}

int fhestr_graph_commit(fhestr_graph* g) {
    if (!g) return FHESTR_E_INVALID;
    g->g.commit();
    g->compiled = false;
    return FHESTR_OK;
}

int fhestr_graph_execute(fhestr_graph* g, fhestr_engine* e, uint32_t rank, uint32_t world) {
    fhestr_program* p = nullptr;
    int rc = fhestr_graph_bind(g, e, &p);
    if (rc) return rc;
    rc = fhestr_program_run(e, p, 0, (uint32_t)g->prog.level_pbs.size(), rank, world);
    if (!rc) rc = fhestr_sync(e);
    if (rc) g->err = std::string("program_run: ") + fhestr_last_error(e);
    fhestr_program_destroy(p);
    if (rc) return rc;
    return fhestr_graph_commit(g);
}

int fhestr_graph_char_op(fhestr_graph* g, int op, uint32_t a, uint32_t b, uint32_t c, uint32_t* out) {
    try {
        return graph_char_op_impl(g, op, a, b, c, out);
    } catch (const std::exception& ex) {   // nothing may unwind across the C ABI
        if (g) g->err = std::string("fhestr_graph_char_op: ") + ex.what();
        return FHESTR_E_STATE;
    } catch (...) {
        if (g) g->err = "fhestr_graph_char_op: unknown exception";
        return FHESTR_E_STATE;
    }
}

int fhestr_graph_string_op(fhestr_graph* g, int method, int fast, const fhestr_str_arg* args, uint32_t n_args,
                           uint64_t clear_n, uint32_t* out_chars, uint32_t out_cap, uint32_t* out_len,
                           uint32_t* out_char) {
    try {
        return graph_string_op_impl(g, method, fast, args, n_args, clear_n, out_chars, out_cap, out_len, out_char);
    } catch (const std::exception& ex) {   // nothing may unwind across the C ABI
        if (g) g->err = std::string("fhestr_graph_string_op: ") + ex.what();
        return FHESTR_E_STATE;
    } catch (...) {
        if (g) g->err = "fhestr_graph_string_op: unknown exception";
        return FHESTR_E_STATE;
    }
}

int fhestr_graph_split_op(fhestr_graph* g, int method, int fast, const fhestr_str_arg* args, uint32_t n_args,
                          uint32_t* out_chars, uint32_t out_cap, uint32_t* n_buffers, uint32_t* buffer_len,
                          uint32_t* out_found) {
    try {
        return graph_split_op_impl(g, method, fast, args, n_args, out_chars, out_cap, n_buffers, buffer_len, out_found);
    } catch (const std::exception& ex) {   // nothing may unwind across the C ABI
        if (g) g->err = std::string("fhestr_graph_split_op: ") + ex.what();
        return FHESTR_E_STATE;
    } catch (...) {
        if (g) g->err = "fhestr_graph_split_op: unknown exception";
        return FHESTR_E_STATE;
    }
}

int fhestr_graph_compile(fhestr_graph* g, uint32_t slot_align, fhestr_graph_info* info) {
    try {
        return graph_compile_impl(g, slot_align, info);
    } catch (const std::exception& ex) {   // nothing may unwind across the C ABI
        if (g) g->err = std::string("fhestr_graph_compile: ") + ex.what();
        return FHESTR_E_STATE;
    } catch (...) {
        if (g) g->err = "fhestr_graph_compile: unknown exception";
        return FHESTR_E_STATE;
    }
}

int fhestr_graph_bind(fhestr_graph* g, fhestr_engine* e, fhestr_program** out) {
    try {
        return graph_bind_impl(g, e, out);
    } catch (const std::exception& ex) {   // nothing may unwind across the C ABI
        if (g) g->err = std::string("fhestr_graph_bind: ") + ex.what();
        return FHESTR_E_STATE;
    } catch (...) {
        if (g) g->err = "fhestr_graph_bind: unknown exception";
        return FHESTR_E_STATE;
    }
}

}  // extern "C"
#pragma GCC visibility pop
