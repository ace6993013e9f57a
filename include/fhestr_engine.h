/*
 * fhestr_engine.h -- C ABI of libfhestr_engine.so, the B200-native batched-PBS engine that sits under
 * fhestring's per-character primitives.
 *
 * The reference (MakisChristou/fhestring) has no FFI: the seam this header defines is the set of
 * tfhe-rs calls made by /root/reference/src/ciphertext/fheasciichar.rs (lines 23, 28, 32, 36-37, 41-42,
 * 46-47, 51-52, 56-57, 61-62, 70, 79, 84, 89, 99, 102), the key hand-over of
 * /root/reference/src/client_key.rs:31-39 and the server-key holder /root/reference/src/server_key/mod.rs:13-16.
 * INTEGRATION.md shows the Rust `-sys` binding a maintainer would put behind those call sites.
 *
 * Conventions
 *   - plain C types only; no C++ or torch types cross this boundary; nothing returned points into
 *     engine memory except the error string.
 *   - every function returns 0 on success, a negative FHESTR_E_* code on failure;
 *     fhestr_last_error() gives the message.  The engine never falls back to a CPU path.
 *   - ciphertext blocks live in a device-resident arena of big-key LWE ciphertexts
 *     ([arena_blocks][N+1] u64, mask then body); callers name them by block index.
 *   - one engine per process and GPU, one submitting host thread (calls are not re-entrant).
 */
#ifndef FHESTR_ENGINE_H
#define FHESTR_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FHESTR_OK 0
#define FHESTR_E_INVALID (-1)   /* bad argument / unsupported parameter set */
#define FHESTR_E_CUDA (-2)      /* CUDA runtime error (message has the call site) */
#define FHESTR_E_STATE (-3)     /* keys not loaded, arena overflow, ... */
#define FHESTR_E_NOGPU (-4)     /* no usable sm_100 device: there is no CPU fallback */

#define FHESTR_MAX_TERMS 16

typedef struct fhestr_engine fhestr_engine;

/* TFHE parameter set, passed as data and validated.  Supported by the kernels:
 * N = 2048, k = 1, pbs_level = 1, pbs_base_log = 23, n <= 767, (ks_base_log+1)*ks_level <= 32. */
typedef struct {
    int32_t n;            /* small LWE dimension (742) */
    int32_t N;            /* polynomial size (2048) */
    int32_t k;            /* GLWE dimension (1) */
    int32_t pbs_base_log; /* 23 */
    int32_t pbs_level;    /* 1 */
    int32_t ks_base_log;  /* 3 */
    int32_t ks_level;     /* 5 */
    int32_t delta_log;    /* 59: plaintext = value << delta_log, bit 63 is the padding bit */
} fhestr_params;

/* One PBS job (== shortint apply_lookup_table on a leveled combination of blocks):
 *   arena[dst] = PBS_lut( sum_t coeff[t] * arena[src[t]]  +  constant * e_body )
 * lut < 0 means "leveled only": the linear combination is written to dst without bootstrapping. */
typedef struct {
    uint32_t dst;
    int32_t lut;
    uint32_t n_terms;
    uint32_t src[FHESTR_MAX_TERMS];
    int32_t coeff[FHESTR_MAX_TERMS];
    uint64_t constant;    /* raw torus value added to the body (e.g. value << delta_log) */
} fhestr_job;

/* ---- lifetime ------------------------------------------------------------------------------- */
/* external_arena: optional device pointer to arena_blocks*(N+1) u64 owned by the caller (e.g. a torch
 * tensor, so torch.distributed can all-gather it); NULL lets the engine allocate. */
int fhestr_engine_create(const fhestr_params* params, int device, uint64_t arena_blocks,
                         void* external_arena, fhestr_engine** out);
void fhestr_engine_destroy(fhestr_engine* e);
const char* fhestr_last_error(const fhestr_engine* e);   /* e may be NULL: last create() error */
/* run all engine work on this cudaStream_t (e.g. torch's current stream); NULL = engine's own */
int fhestr_set_stream(fhestr_engine* e, void* cuda_stream);
int fhestr_sync(fhestr_engine* e);
void* fhestr_arena_ptr(fhestr_engine* e);                /* device pointer, for NCCL plumbing */

/* ---- key store (replaces holding tfhe::integer::ServerKey, server_key/mod.rs:13-16) ----------- */
/* bsk_std: host, standard-domain GGSW [n][pbs_level][k+1 rows][k+1 cols][N] u64
 * ksk:     host, [N*k][ks_level][n+1] u64, level 1 first.
 * Converted once to the engine's Fourier / tiled layouts (kernel K6). */
int fhestr_load_keys(fhestr_engine* e, const uint64_t* bsk_std, const uint64_t* ksk);

/* ---- LUT registry (replaces generate_lookup_table; kernel K5) ---------------------------------- */
/* table has 2^(63-delta_log) entries (16): f(x) over the 4-bit block value.  A block whose value v carries the padding
 * bit (v in [16, 32)) reads -f(v - 16) (negacyclic).  HALF-STEP tables: every entry given as 0x80 | e stands for
 * e - 1/2, and the engine adds the 1/2 back to the result, so f(v) = e[v] for v < 16 and 1 - e[v - 16] for
 * v in [16, 32): with e = 0 everywhere that is the threshold [v >= 16], which turns an AND or an OR over 16 flags
 * (sum + constant) into ONE PBS. */
int fhestr_lut_register(fhestr_engine* e, const uint8_t* table, int32_t* lut_id);
int fhestr_lut_download(fhestr_engine* e, int32_t lut_id, uint64_t* out_poly /* [N] */);

/* ---- ciphertext arena (replaces owning BaseRadixCiphertext values, fheasciichar.rs:7-10) ------- */
/* upload is asynchronous on the engine stream: a pageable `host` buffer is staged before the call returns, a PINNED
 * one must stay valid until the next synchronising call (fhestr_sync, fhestr_ct_download) */
int fhestr_ct_upload(fhestr_engine* e, uint32_t first_block, uint32_t count, const uint64_t* host);
int fhestr_ct_download(fhestr_engine* e, uint32_t first_block, uint32_t count, uint64_t* host);
/* blocks named one by one (the chars of a result string sit wherever their last level left them): gathered on the
 * device, ONE copy to the host */
int fhestr_ct_download_slots(fhestr_engine* e, const uint32_t* slots, uint32_t count, uint64_t* host);
/* the same without the synchronisation: `pinned_host` must be page-locked and stay valid until the next synchronising
 * call on the stream the copy was issued on (lets a caller overlap the copy-out of batch i with the PBS of batch i+1
 * by switching streams with fhestr_set_stream and ordering them with its own events) */
int fhestr_ct_download_async(fhestr_engine* e, uint32_t first_block, uint32_t count, uint64_t* pinned_host);
/* create_trivial_radix (fheasciichar.rs:23): block b gets mask 0, body values[b] << delta_log */
int fhestr_ct_trivial(fhestr_engine* e, uint32_t first_block, uint32_t count, const uint8_t* values);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* One dependency level of independent jobs: K0+K1 (linear combination + keyswitch), then
 * K2+K3+K4 (mod-switch, blind rotation, sample extract).  Asynchronous on the engine stream. */
int fhestr_pbs_batch(fhestr_engine* e, const fhestr_job* jobs, uint32_t n_jobs);

/* A program = several dependency levels uploaded once and replayed without host work in between
 * (how the levelised string algorithms run).  level_offsets has n_levels+1 entries into jobs. */
typedef struct fhestr_program fhestr_program;
int fhestr_program_create(fhestr_engine* e, const fhestr_job* jobs, const uint32_t* level_offsets,
                          uint32_t n_levels, fhestr_program** out);
/* run levels [first_level, last_level).  world > 1 (after fhestr_comm_init): each level's PBS jobs are
 * sharded over the ranks (fhestr_shard_range), the result blocks are all-gathered in place over NCCL, and
 * the leveled jobs run replicated -- every rank ends each level with the same arena. */
int fhestr_program_run(fhestr_engine* e, fhestr_program* p, uint32_t first_level, uint32_t last_level,
                       uint32_t rank, uint32_t world);
int fhestr_program_level_jobs(const fhestr_program* p, uint32_t level, uint32_t* n_jobs);
void fhestr_program_destroy(fhestr_program* p);

/* ---- op graph: the reference's per-char primitives and string methods, recorded then batched ---------- */
/* How the unchanged sequential callers of /root/reference/src/server_key/*.rs become dependency-level
 * batches: every FheAsciiChar the Rust side holds is a 32-bit char id into a graph; the primitives of
 * /root/reference/src/ciphertext/fheasciichar.rs:17-168 RECORD instead of computing; decrypt (or any
 * explicit flush) compiles what is reachable into levels of independent PBS jobs and runs them.  Eager use
 * is "record one op, execute" -- same calls, no batching. */
typedef struct fhestr_graph fhestr_graph;

enum {  /* fhestr_graph_char_op: FheAsciiChar methods (fheasciichar.rs line) */
    FHESTR_OP_EQ = 0 /* :35 */, FHESTR_OP_NE = 1 /* :40 */, FHESTR_OP_LE = 2 /* :45 */, FHESTR_OP_LT = 3 /* :50 */,
    FHESTR_OP_GE = 4 /* :55 */, FHESTR_OP_GT = 5 /* :60 */, FHESTR_OP_BITAND = 6 /* :65 */, FHESTR_OP_BITOR = 7 /* :74 */,
    FHESTR_OP_SUB = 8 /* :83 */, FHESTR_OP_ADD = 9 /* :87 */, FHESTR_OP_IF_THEN_ELSE = 10 /* :93 */,
    FHESTR_OP_IS_WHITESPACE = 11 /* :106 */, FHESTR_OP_IS_UPPERCASE = 12 /* :132 */,
    FHESTR_OP_IS_LOWERCASE = 13 /* :146 */, FHESTR_OP_FLIP = 14 /* :161 */
};
enum {  /* fhestr_graph_string_op: MyServerKey methods (server_key/mod.rs line unless noted) */
    FHESTR_M_CONTAINS = 0 /* :151 */, FHESTR_M_ENDS_WITH = 1 /* :241 */, FHESTR_M_STARTS_WITH = 2 /* :344 */,
    FHESTR_M_IS_EMPTY = 3 /* :431 */, FHESTR_M_LEN = 4 /* :478 */, FHESTR_M_REPEAT_CLEAR = 5 /* :517 */,
    FHESTR_M_REPEAT = 6 /* :567 */, FHESTR_M_REPLACE = 7 /* :624 */, FHESTR_M_RFIND = 8 /* :727 */,
    FHESTR_M_FIND = 9 /* :1010 */, FHESTR_M_EQ = 10 /* :1122 */, FHESTR_M_NE = 11 /* :1178 */,
    FHESTR_M_EQ_IGNORE_CASE = 12 /* :1221 */, FHESTR_M_STRIP_PREFIX = 13 /* :1261 */,
    FHESTR_M_STRIP_SUFFIX = 14 /* :1335 */, FHESTR_M_LT = 15 /* :1577 */, FHESTR_M_LE = 16 /* :1613 */,
    FHESTR_M_GT = 17 /* :1649 */, FHESTR_M_GE = 18 /* :1685 */, FHESTR_M_REPLACEN = 19 /* :1729 */,
    FHESTR_M_CONCATENATE = 20 /* :1864 */, FHESTR_M_TO_UPPER = 21 /* :65 */, FHESTR_M_TO_LOWER = 22 /* :110 */,
    FHESTR_M_TRIM_END = 23 /* trim.rs:36 */, FHESTR_M_TRIM_START = 24 /* trim.rs:86 */, FHESTR_M_TRIM = 25 /* trim.rs:146 */,
    FHESTR_M_BUBBLE_ZEROES_RIGHT = 26 /* utils.rs:28 */
};

typedef struct {            /* one string argument: char ids in order (FheString / Vec<FheAsciiChar>) */
    const uint32_t* chars;
    uint32_t len;
} fhestr_str_arg;

typedef struct {
    uint32_t n_levels, n_jobs, n_luts, n_trivial;
    uint32_t slots_used;    /* arena blocks the graph needs so far */
    uint64_t n_pbs;         /* PBS jobs in this compile */
    uint64_t n_pbs_recorded;/* PBS nodes recorded since the graph was created (before dead-code elimination) */
} fhestr_graph_info;

int fhestr_graph_create(int32_t delta_log, fhestr_graph** out);
void fhestr_graph_destroy(fhestr_graph* g);
const char* fhestr_graph_last_error(const fhestr_graph* g);
/* FheAsciiChar::encrypt (fheasciichar.rs:27): `count` encrypted chars; slots[4*i + b] is the arena block the
 * caller must upload block b of char i to (fhestr_ct_upload) before executing */
int fhestr_graph_input_chars(fhestr_graph* g, uint32_t count, uint32_t* ids, uint32_t* slots);
/* FheAsciiChar::encrypt_trivial (fheasciichar.rs:17-25) */
int fhestr_graph_trivial_chars(fhestr_graph* g, const uint8_t* values, uint32_t count, uint32_t* ids);
int fhestr_graph_char_op(fhestr_graph* g, int op, uint32_t a, uint32_t b, uint32_t c, uint32_t* out);
/* fast != 0: depth-minimised recording (same plaintext for every input); 0: the reference's op order.
 * Results: a string (out_chars/out_len, capacity out_cap) and/or a single char (*out_char), by method.
 * The reference's panics ("Maximum supported size for find reached") come back as FHESTR_E_INVALID. */
int fhestr_graph_string_op(fhestr_graph* g, int method, int fast, const fhestr_str_arg* args, uint32_t n_args,
                           uint64_t clear_n, uint32_t* out_chars, uint32_t out_cap, uint32_t* out_len,
                           uint32_t* out_char);
enum {  /* fhestr_graph_split_op: the split family of /root/reference/src/server_key/split.rs (line) */
    FHESTR_S_SPLIT = 0 /* :1038 */, FHESTR_S_RSPLIT = 1 /* :439 */, FHESTR_S_SPLIT_INCLUSIVE = 2 /* :1155 */,
    FHESTR_S_SPLIT_TERMINATOR = 3 /* :1267 */, FHESTR_S_RSPLIT_TERMINATOR = 4 /* :806 */,
    FHESTR_S_RSPLIT_ONCE = 5 /* :681 */, FHESTR_S_SPLITN = 6 /* :1497 */, FHESTR_S_RSPLITN = 7 /* :553 */,
    FHESTR_S_SPLIT_ASCII_WHITESPACE = 8 /* :1377 */
};
/* args: string, pattern (not for SPLIT_ASCII_WHITESPACE), and for SPLITN / RSPLITN the encrypted n as a 1-char
 * argument.  Result (FheSplit, fhesplit.rs:5-8): *n_buffers buffers of *buffer_len chars each, written to
 * out_chars[b * buffer_len + i] (capacity out_cap chars), and the pattern_found flag in *out_found. */
int fhestr_graph_split_op(fhestr_graph* g, int method, int fast, const fhestr_str_arg* args, uint32_t n_args,
                          uint32_t* out_chars, uint32_t out_cap, uint32_t* n_buffers, uint32_t* buffer_len,
                          uint32_t* out_found);
int fhestr_graph_mark_output(fhestr_graph* g, const uint32_t* ids, uint32_t count);
/* levelise everything the marked outputs need; slot_align = number of ranks the levels will be sharded over */
int fhestr_graph_compile(fhestr_graph* g, uint32_t slot_align, fhestr_graph_info* info);
/* the compiled job list (for inspection, tests, or a caller that drives fhestr_program_* itself) */
int fhestr_graph_get_program(const fhestr_graph* g, fhestr_job* jobs, uint32_t* level_offsets /* n_levels+1 */,
                             uint32_t* level_pbs, uint32_t* level_first_dst);
int fhestr_graph_get_luts(const fhestr_graph* g, uint8_t* tables /* [n_luts][16], graph-local ids */);
int fhestr_graph_get_trivials(const fhestr_graph* g, uint32_t* slots, uint8_t* values);
int fhestr_graph_char_slots(const fhestr_graph* g, const uint32_t* ids, uint32_t count, uint32_t* slots /* [count][4] */);
/* the arena blocks [0, first_free) are in use by something the graph did not record (a bound program of an earlier,
 * identical query run again on new inputs: a plan cache above the ABI): later slots are handed out from first_free on */
int fhestr_graph_reserve_slots(fhestr_graph* g, uint32_t first_free);
/* run the compiled levels on the engine (registers LUTs, writes trivial outputs, shards each level over
 * `world` ranks and, when the engine has a communicator, all-gathers each level's results), then commit:
 * computed chars become inputs of whatever is recorded next */
int fhestr_graph_execute(fhestr_graph* g, fhestr_engine* e, uint32_t rank, uint32_t world);
/* the same split in two, for callers that time or interleave the run themselves */
int fhestr_graph_bind(fhestr_graph* g, fhestr_engine* e, fhestr_program** out);
int fhestr_graph_commit(fhestr_graph* g);

/* ---- multi-GPU: one process per GPU, keys and arena replicated; per level either P2P stores + flag barrier (peer_*)
 * or one in-place NCCL all-gather (comm_*) ------------------------------------------------------------------ */
/* unique_id: 128 bytes from fhestr_comm_unique_id on rank 0, handed to the other ranks by the host */
int fhestr_comm_unique_id(void* unique_id_128);
int fhestr_comm_init(fhestr_engine* e, uint32_t rank, uint32_t world, const void* unique_id_128);
int fhestr_comm_destroy(fhestr_engine* e);
/* Preferred multi-GPU data path: no collective at all.  Every rank exports its arena (and a small flag array)
 * as cudaIpc handles (64 bytes each), the host exchanges them, fhestr_peer_attach maps the peers' memory; from
 * then on the sample-extract epilogue of the blind rotation stores each result block into ALL arenas over NVLink
 * and a flag barrier over peer memory closes the level.  arena_handles / flags_handles: [world][64] bytes by rank.
 * A barrier that times out (a peer died) sets a status word: every synchronising call (fhestr_sync,
 * fhestr_ct_download, fhestr_graph_execute) then fails with FHESTR_E_STATE instead of handing out incomplete
 * results; fhestr_peer_status reads the word.  A run starts with one such barrier, so no rank stores into a peer's
 * arena before that peer has finished with the previous run's results. */
int fhestr_peer_export(fhestr_engine* e, void* arena_handle_64, void* flags_handle_64);
int fhestr_peer_attach(fhestr_engine* e, uint32_t rank, uint32_t world, const void* arena_handles, const void* flags_handles);
int fhestr_peer_detach(fhestr_engine* e);
int fhestr_peer_status(fhestr_engine* e, uint32_t* timed_out);
/* the slice of a level's n_jobs PBS jobs that `rank` of `world` computes: [lo, hi), per = ceil(n_jobs/world)
 * (the all-gather moves `per` blocks per rank, so a level's result slots are padded to per*world) */
void fhestr_shard_range(uint32_t n_jobs, uint32_t rank, uint32_t world, uint32_t* lo, uint32_t* hi, uint32_t* per);

/* ---- test / measurement hooks ------------------------------------------------------------------ */
/* keyswitch only: small-key LWEs [n_jobs][n+1] to host (bit-exact check against the oracle) */
int fhestr_debug_keyswitch(fhestr_engine* e, const fhestr_job* jobs, uint32_t n_jobs, uint64_t* host_out);
/* blind rotation only on given small-key LWEs [count][n+1]; optional start accumulators
 * [count][2][N] instead of the rotated LUT; returns raw accumulators [count][2][N] */
int fhestr_debug_blind_rotate(fhestr_engine* e, const uint64_t* ks_host, const int32_t* lut_ids,
                              const uint64_t* init_acc_host, uint32_t count, uint64_t* acc_out_host);
/* DFMA-saturating microbenchmark: measured FP64 FMA throughput of this GPU in TFLOP/s */
int fhestr_measure_fp64_peak(fhestr_engine* e, double* tflops, double* sm_clock_mhz_hint);
/* number of kernels this engine has launched so far (bench.py's gpu_launches) */
uint64_t fhestr_kernel_launches(const fhestr_engine* e);
/* Which blind-rotation kernel runs a level.  mode 0 (default): by level size -- levels of at most wide_max_jobs PBS
 * jobs (0 = three times the SM count) run on the latency kernel, larger ones on the throughput kernel (four PBS per
 * SM, one pair of warps each) in waves of four jobs per SM, with a small last-wave remainder handed to the latency
 * kernel behind the full waves (at most one job per SM behind one to three waves, up to three per SM behind a single
 * one).  The latency kernel has two forms, also picked by size: ONE PBS per SM over 128 threads
 * with a two-tile key ring fed by bulk TMA (up to one job per SM), and TWO PBS per SM over 256 threads sharing one
 * key tile (above).  mode 1 forces the throughput kernel, 2 the latency kernel (form by size), 3 / 4 its single /
 * pair form (tests, measurements).  All kernels compute the same function; their outputs are different valid
 * ciphertexts of the same plaintext (different f64 FFT orders), see DESIGN.md. */
int fhestr_set_br_mode(fhestr_engine* e, int mode, int wide_max_jobs);
/* keyswitch implementation: 0 = tensor cores (u8 limb-split GEMM on tcgen05.mma kind::i8, default), 1 = CUDA cores (u64 IMAD);
 * both are exact and produce identical words */
int fhestr_set_keyswitch_path(fhestr_engine* e, int path);

/* per-kernel device timing (CUDA events on the engine stream around every keyswitch / blind-rotation
 * launch); get_timing synchronises the stream and returns the totals since the last reset */
int fhestr_set_timing(fhestr_engine* e, int enable);
int fhestr_get_timing(fhestr_engine* e, double* keyswitch_ms, double* blind_rotate_ms,
                      uint64_t* blind_rotate_launches, uint64_t* blind_rotate_pbs);

/* ---- client side (replaces MyClientKey, /root/reference/src/client_key.rs:9-106): host-only ------- */
/* Key generation, block/string encryption and decryption.  Runs on the CPU like the reference's
 * client; it is not on the PBS path.  All randomness (secret keys, masks, noise) is ChaCha20 output under a 256-bit
 * key: seed == 0 takes that key from the operating system (getrandom) -- the production setting, as tfhe-rs seeds
 * gen_keys_radix; seed != 0 derives it from the seed and is for TESTS AND BENCHMARKS ONLY (reproducible keys). */
typedef struct fhestr_client fhestr_client;
int fhestr_client_create(const fhestr_params* params, double lwe_std, double glwe_std, uint64_t seed,
                         fhestr_client** out);
void fhestr_client_destroy(fhestr_client* c);
/* gen_keys_radix (client_key.rs:31): fills the server key material in the layouts fhestr_load_keys takes */
int fhestr_client_server_keys(fhestr_client* c, uint64_t* bsk_std, uint64_t* ksk);
/* secret key bits, one byte per bit (tests use them to decrypt with the oracle) */
int fhestr_client_secret_keys(const fhestr_client* c, uint8_t* s_lwe /* [n] */, uint8_t* s_glwe /* [N] */);
/* block-level: values are raw block values (< 2^(64-delta_log)); ciphertexts are [count][N+1] u64 */
int fhestr_client_encrypt_blocks(fhestr_client* c, const uint8_t* values, uint32_t count, uint64_t* cts);
int fhestr_client_decrypt_blocks(const fhestr_client* c, const uint64_t* cts, uint32_t count,
                                 uint8_t* values /* (phase + delta/2) >> delta_log, padding bit dropped */,
                                 int64_t* phase_err /* optional: phase - value*delta, signed */);
/* radix level (client_key.rs:45-106): one u8 = 4 blocks of 2 message bits, little endian */
int fhestr_client_encrypt_u8(fhestr_client* c, const uint8_t* bytes, uint32_t count, uint64_t* cts /* [count][4][N+1] */);
int fhestr_client_decrypt_u8(const fhestr_client* c, const uint64_t* cts, uint32_t count, uint8_t* bytes);

#ifdef __cplusplus
}
#endif
#endif /* FHESTR_ENGINE_H */
