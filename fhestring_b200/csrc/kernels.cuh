// kernels.cuh -- launcher declarations shared by engine.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fhestr_engine.h"
#include "br_core.cuh"
#include "br_wide.cuh"

namespace fhestr {

struct BrBatchArgs {
    const u64* ks;            // [B][n+1] keyswitched LWEs
    const u64* luts;          // [n_luts][N] body polynomials
    const u64* lut_post;      // [n_luts] constant added to the body of the extracted LWE (half-step tables), or nullptr
    const int32_t* lut_ids;   // [B] (one per job; taken from jobs when jobs != nullptr)
    const fhestr_job* jobs;   // [B] device job list (dst + lut) or nullptr
    u64* arena;               // [blocks][N+1]
    const cplx* bsk;          // Fourier BSK, engine layout
    const cplx* tf;           // inter-pass twiddles of the throughput kernel, [32 k1][32 n2]
    const u64* init_acc;      // optional [B][2][N]
    u64* out_acc;             // optional [B][2][N]
    u64* peer_arena[7];       // arenas of the other ranks (cudaIpc-mapped), n_peers of them
    int n_peers;
    const cplx* bsk_w;        // Fourier BSK in the spectrum order of the latency kernel (br_wide.cuh), [n][kWKeyTile]
    const WideConsts* wide_tab;   // [kWT] per-thread transform constants of the latency kernel
    int n;
    int B;
};

struct KsBatchArgs {
    const fhestr_job* jobs;   // [B] device
    const u64* arena;         // [blocks][N+1]
    const u64* ksk;           // [N][L][n+1]
    const u64* ksk_corr;      // [n+1]  2^(base_log-1) * sum_{i,lvl} ksk[i][lvl][c]
    u64* ks_out;              // [B][n+1]
    int n, B, base_log, level;
    // tensor-core path (keyswitch_tc.cu)
    const unsigned char* ksk8;   // [cols_padded*8][N*level] limb matrix, K-major
    unsigned char* ks_digits;    // [rows padded to 128][N*level] unsigned digits
    u64* ks_body;                // [B]
};

// K0+K1: linear combination + LWE keyswitch.  Returns the number of kernels launched.
int launch_keyswitch(const KsBatchArgs& a, cudaStream_t s);
// K2+K3+K4: mod-switch + blind rotation + sample extract, throughput form (one PBS per pair of warps, br_core.cuh)
int launch_blind_rotate(const BrBatchArgs& a, cudaStream_t s);
cudaError_t blind_rotate_configure();
// the same in its latency form: one PBS per 128-thread CTA, one CTA per SM, key tiles by bulk TMA (br_wide.cuh)
int launch_blind_rotate_wide(const BrBatchArgs& a, cudaStream_t s);
// pair form: two PBS per 256-thread CTA sharing one TMA-fed key tile per step
int launch_blind_rotate_wide2(const BrBatchArgs& a, cudaStream_t s);
cudaError_t blind_rotate_wide_configure();
int launch_bsk_convert_wide(const u64* bsk_std, int n, const WideConsts* tab, cplx* out, cudaStream_t s);
cudaError_t keyswitch_configure();  // opt in to the large dynamic shared memory carve-out
// the same on the tensor cores (tcgen05.mma kind::i8 u8 x u8 -> s32 limb-split GEMM, TMA-fed; keyswitch_tc.cu); bit-identical results
int launch_keyswitch_tc(const KsBatchArgs& a, cudaStream_t s);
cudaError_t keyswitch_tc_configure();
int launch_ksk_limbs(const u64* ksk, int K, int n, unsigned char* out, cudaStream_t s);
int ks_cols_padded(int n);
size_t ks_digit_rows(size_t B);
// cross-GPU level barrier over peer memory: every rank stores `epoch` into slot `rank` of each peer's flag array
// (after a system-scope fence), then waits until its own array holds `epoch` in every slot.  Returns launches.
int launch_peer_barrier(uint32_t* const* peer_flags, uint32_t* my_flags, uint32_t* status, int rank, int world, uint32_t epoch, cudaStream_t s);
// leveled jobs (lut < 0): dst = sum coeff*src + constant*e_body
int launch_linear(const fhestr_job* jobs, int B, u64* arena, cudaStream_t s);
// K5: 16-entry table -> body polynomial
int launch_lut_poly(const uint8_t* table_dev, int entries, int delta_log, u64* out, cudaStream_t s);
// K6: standard-domain BSK [n][2][2][N] -> Fourier layout [n][kBskStepElems]
int launch_bsk_convert(const u64* bsk_std, int n, const cplx* tf, cplx* out, cudaStream_t s);
int launch_ksk_correction(const u64* ksk, int rows /* N*L */, int n, int base_log, u64* corr, cudaStream_t s);
int launch_trivial(u64* arena, uint32_t first, uint32_t count, const uint8_t* values_dev, int delta_log, cudaStream_t s);
// arena[slots[b]] -> out[b], b < count
int launch_gather_blocks(const u64* arena, const uint32_t* slots_dev, uint32_t count, u64* out, cudaStream_t s);
// DFMA microbenchmark; returns total FMA count issued through *fmas
int launch_dfma_peak(double* sink, int iters, unsigned long long* fmas, cudaStream_t s);

}  // namespace fhestr
