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
/* table has 2^(63-delta_log) entries (16): f(x) over the 4-bit block value */
int fhestr_lut_register(fhestr_engine* e, const uint8_t* table, int32_t* lut_id);
int fhestr_lut_download(fhestr_engine* e, int32_t lut_id, uint64_t* out_poly /* [N] */);

/* ---- ciphertext arena (replaces owning BaseRadixCiphertext values, fheasciichar.rs:7-10) ------- */
int fhestr_ct_upload(fhestr_engine* e, uint32_t first_block, uint32_t count, const uint64_t* host);
int fhestr_ct_download(fhestr_engine* e, uint32_t first_block, uint32_t count, uint64_t* host);
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
/* run levels [first_level, last_level) ; rank/world shard each level's jobs (multi-GPU: the caller
 * all-gathers the arena slices between levels; see fhestr_program_level_range) */
int fhestr_program_run(fhestr_engine* e, fhestr_program* p, uint32_t first_level, uint32_t last_level,
                       uint32_t rank, uint32_t world);
int fhestr_program_level_jobs(const fhestr_program* p, uint32_t level, uint32_t* n_jobs);
void fhestr_program_destroy(fhestr_program* p);

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
/* blind-rotation launch shape override for experiments: PBS per CTA (1, 2 or 4; 0 = automatic) */
int fhestr_set_pbs_per_cta(fhestr_engine* e, int pbs_per_cta);

/* per-kernel device timing (CUDA events on the engine stream around every keyswitch / blind-rotation
 * launch); get_timing synchronises the stream and returns the totals since the last reset */
int fhestr_set_timing(fhestr_engine* e, int enable);
int fhestr_get_timing(fhestr_engine* e, double* keyswitch_ms, double* blind_rotate_ms,
                      uint64_t* blind_rotate_launches, uint64_t* blind_rotate_pbs);

/* ---- client side (replaces MyClientKey, /root/reference/src/client_key.rs:9-106): host-only ------- */
/* Key generation, block/string encryption and decryption.  Runs on the CPU like the reference's
 * client; it is not on the PBS path.  Deterministic for a given seed. */
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
