//! Raw bindings to `include/fhestr_engine.h` -- the entry points the patched
//! `src/ciphertext/fheasciichar.rs`, `src/client_key.rs` and `src/server_key/mod.rs` of fhestring call.
//! Authored, NOT compiled in this repository's build image (no Rust toolchain); the same ABI is exercised
//! end to end through ctypes by `tests/`.
#![allow(non_camel_case_types, non_snake_case)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct fhestr_engine { _p: [u8; 0] }
#[repr(C)] pub struct fhestr_graph { _p: [u8; 0] }
#[repr(C)] pub struct fhestr_program { _p: [u8; 0] }

pub const FHESTR_OK: c_int = 0;
pub const FHESTR_E_INVALID: c_int = -1;
pub const FHESTR_E_CUDA: c_int = -2;
pub const FHESTR_E_STATE: c_int = -3;
pub const FHESTR_E_NOGPU: c_int = -4;
pub const FHESTR_MAX_TERMS: usize = 16;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct fhestr_params {
    pub n: i32, pub N: i32, pub k: i32, pub pbs_base_log: i32, pub pbs_level: i32,
    pub ks_base_log: i32, pub ks_level: i32, pub delta_log: i32,
}

/// PARAM_MESSAGE_2_CARRY_2_KS_PBS (src/main.rs:3,43)
pub const PARAM_MESSAGE_2_CARRY_2_KS_PBS: fhestr_params = fhestr_params {
    n: 742, N: 2048, k: 1, pbs_base_log: 23, pbs_level: 1, ks_base_log: 3, ks_level: 5, delta_log: 59,
};

#[repr(C)]
#[derive(Clone, Copy)]
pub struct fhestr_job {
    pub dst: u32, pub lut: i32, pub n_terms: u32,
    pub src: [u32; FHESTR_MAX_TERMS], pub coeff: [i32; FHESTR_MAX_TERMS], pub constant: u64,
}

#[repr(C)]
pub struct fhestr_str_arg { pub chars: *const u32, pub len: u32 }

#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct fhestr_graph_info {
    pub n_levels: u32, pub n_jobs: u32, pub n_luts: u32, pub n_trivial: u32, pub slots_used: u32,
    pub n_pbs: u64, pub n_pbs_recorded: u64,
}

// FheAsciiChar methods (fheasciichar.rs line)
pub const FHESTR_OP_EQ: c_int = 0; pub const FHESTR_OP_NE: c_int = 1; pub const FHESTR_OP_LE: c_int = 2;
pub const FHESTR_OP_LT: c_int = 3; pub const FHESTR_OP_GE: c_int = 4; pub const FHESTR_OP_GT: c_int = 5;
pub const FHESTR_OP_BITAND: c_int = 6; pub const FHESTR_OP_BITOR: c_int = 7; pub const FHESTR_OP_SUB: c_int = 8;
pub const FHESTR_OP_ADD: c_int = 9; pub const FHESTR_OP_IF_THEN_ELSE: c_int = 10;
pub const FHESTR_OP_IS_WHITESPACE: c_int = 11; pub const FHESTR_OP_IS_UPPERCASE: c_int = 12;
pub const FHESTR_OP_IS_LOWERCASE: c_int = 13; pub const FHESTR_OP_FLIP: c_int = 14;
// MyServerKey methods (fhestr_graph_string_op)
pub const FHESTR_M_CONTAINS: c_int = 0; pub const FHESTR_M_ENDS_WITH: c_int = 1; pub const FHESTR_M_STARTS_WITH: c_int = 2;
pub const FHESTR_M_IS_EMPTY: c_int = 3; pub const FHESTR_M_LEN: c_int = 4; pub const FHESTR_M_REPEAT_CLEAR: c_int = 5;
pub const FHESTR_M_REPEAT: c_int = 6; pub const FHESTR_M_REPLACE: c_int = 7; pub const FHESTR_M_RFIND: c_int = 8;
pub const FHESTR_M_FIND: c_int = 9; pub const FHESTR_M_EQ: c_int = 10; pub const FHESTR_M_NE: c_int = 11;
pub const FHESTR_M_EQ_IGNORE_CASE: c_int = 12; pub const FHESTR_M_STRIP_PREFIX: c_int = 13;
pub const FHESTR_M_STRIP_SUFFIX: c_int = 14; pub const FHESTR_M_LT: c_int = 15; pub const FHESTR_M_LE: c_int = 16;
pub const FHESTR_M_GT: c_int = 17; pub const FHESTR_M_GE: c_int = 18; pub const FHESTR_M_REPLACEN: c_int = 19;
pub const FHESTR_M_CONCATENATE: c_int = 20; pub const FHESTR_M_TO_UPPER: c_int = 21; pub const FHESTR_M_TO_LOWER: c_int = 22;
pub const FHESTR_M_TRIM_END: c_int = 23; pub const FHESTR_M_TRIM_START: c_int = 24; pub const FHESTR_M_TRIM: c_int = 25;
pub const FHESTR_M_BUBBLE_ZEROES_RIGHT: c_int = 26;
// split family (fhestr_graph_split_op)
pub const FHESTR_S_SPLIT: c_int = 0; pub const FHESTR_S_RSPLIT: c_int = 1; pub const FHESTR_S_SPLIT_INCLUSIVE: c_int = 2;
pub const FHESTR_S_SPLIT_TERMINATOR: c_int = 3; pub const FHESTR_S_RSPLIT_TERMINATOR: c_int = 4;
pub const FHESTR_S_RSPLIT_ONCE: c_int = 5; pub const FHESTR_S_SPLITN: c_int = 6; pub const FHESTR_S_RSPLITN: c_int = 7;
pub const FHESTR_S_SPLIT_ASCII_WHITESPACE: c_int = 8;

extern "C" {
    // lifetime, keys, arena
    pub fn fhestr_engine_create(p: *const fhestr_params, device: c_int, arena_blocks: u64, external_arena: *mut c_void,
                                out: *mut *mut fhestr_engine) -> c_int;
    pub fn fhestr_engine_destroy(e: *mut fhestr_engine);
    pub fn fhestr_last_error(e: *const fhestr_engine) -> *const c_char;
    pub fn fhestr_sync(e: *mut fhestr_engine) -> c_int;
    pub fn fhestr_load_keys(e: *mut fhestr_engine, bsk_std: *const u64, ksk: *const u64) -> c_int;
    pub fn fhestr_lut_register(e: *mut fhestr_engine, table: *const u8, lut_id: *mut i32) -> c_int;
    pub fn fhestr_ct_upload(e: *mut fhestr_engine, first_block: u32, count: u32, host: *const u64) -> c_int;
    pub fn fhestr_ct_download(e: *mut fhestr_engine, first_block: u32, count: u32, host: *mut u64) -> c_int;
    pub fn fhestr_ct_trivial(e: *mut fhestr_engine, first_block: u32, count: u32, values: *const u8) -> c_int;
    // raw hot path
    pub fn fhestr_pbs_batch(e: *mut fhestr_engine, jobs: *const fhestr_job, n_jobs: u32) -> c_int;
    pub fn fhestr_program_create(e: *mut fhestr_engine, jobs: *const fhestr_job, level_offsets: *const u32, n_levels: u32,
                                 out: *mut *mut fhestr_program) -> c_int;
    pub fn fhestr_program_run(e: *mut fhestr_engine, p: *mut fhestr_program, first_level: u32, last_level: u32,
                              rank: u32, world: u32) -> c_int;
    pub fn fhestr_program_destroy(p: *mut fhestr_program);
    // op graph
    pub fn fhestr_graph_create(delta_log: i32, out: *mut *mut fhestr_graph) -> c_int;
    pub fn fhestr_graph_destroy(g: *mut fhestr_graph);
    pub fn fhestr_graph_last_error(g: *const fhestr_graph) -> *const c_char;
    pub fn fhestr_graph_input_chars(g: *mut fhestr_graph, count: u32, ids: *mut u32, slots: *mut u32) -> c_int;
    pub fn fhestr_graph_trivial_chars(g: *mut fhestr_graph, values: *const u8, count: u32, ids: *mut u32) -> c_int;
    pub fn fhestr_graph_char_op(g: *mut fhestr_graph, op: c_int, a: u32, b: u32, c: u32, out: *mut u32) -> c_int;
    pub fn fhestr_graph_string_op(g: *mut fhestr_graph, method: c_int, fast: c_int, args: *const fhestr_str_arg, n_args: u32,
                                  clear_n: u64, out_chars: *mut u32, out_cap: u32, out_len: *mut u32, out_char: *mut u32) -> c_int;
    pub fn fhestr_graph_split_op(g: *mut fhestr_graph, method: c_int, fast: c_int, args: *const fhestr_str_arg, n_args: u32,
                                 out_chars: *mut u32, out_cap: u32, n_buffers: *mut u32, buffer_len: *mut u32,
                                 out_found: *mut u32) -> c_int;
    pub fn fhestr_graph_mark_output(g: *mut fhestr_graph, ids: *const u32, count: u32) -> c_int;
    pub fn fhestr_graph_compile(g: *mut fhestr_graph, slot_align: u32, info: *mut fhestr_graph_info) -> c_int;
    pub fn fhestr_graph_char_slots(g: *const fhestr_graph, ids: *const u32, count: u32, slots: *mut u32) -> c_int;
    pub fn fhestr_graph_reserve_slots(g: *mut fhestr_graph, first_free: u32) -> c_int;
    pub fn fhestr_graph_execute(g: *mut fhestr_graph, e: *mut fhestr_engine, rank: u32, world: u32) -> c_int;
    // multi-GPU
    pub fn fhestr_peer_export(e: *mut fhestr_engine, arena_handle_64: *mut c_void, flags_handle_64: *mut c_void) -> c_int;
    pub fn fhestr_peer_attach(e: *mut fhestr_engine, rank: u32, world: u32, arena_handles: *const c_void,
                              flags_handles: *const c_void) -> c_int;
    pub fn fhestr_peer_detach(e: *mut fhestr_engine) -> c_int;
    pub fn fhestr_comm_unique_id(unique_id_128: *mut c_void) -> c_int;
    pub fn fhestr_comm_init(e: *mut fhestr_engine, rank: u32, world: u32, unique_id_128: *const c_void) -> c_int;
}
