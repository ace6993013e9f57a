// keyswitch.cu -- K0+K1 for sm_100a: leveled linear combination of arena blocks fused in front of the
// LWE keyswitch (big key, dimension N=2048 -> small key, dimension n=742).
//
//   x      = sum_t coeff[t] * arena[src[t]] + constant * e_body           (K0, never materialised)
//   out    = (0,...,0,x_body) - sum_i sum_lvl digit_{i,lvl} * KSK[i][lvl]   (K1)
//
// Replaces tfhe-rs keyswitch_lwe_ciphertext behind /root/reference/src/ciphertext/fheasciichar.rs:36-102.
//
// A CTA owns BT = 8 ciphertexts.  Phase 1 decomposes the 8 x 2048 mask words (A.4, balanced base-2^b
// digits) and parks them in shared memory as UNSIGNED digits d' = d + B/2 packed (b+1) bits each; the
// constant -B/2 * sum KSK is a per-key vector folded in at the end, so the inner loop is a pure
// unsigned multiply-accumulate: IMAD.WIDE.U32 + IMAD per (ciphertext, column, level).  Phase 2 streams
// the KSK (58 MiB, L2 resident) once per CTA: each thread owns 3 of the 743 output columns, every KSK
// word it loads is reused for the 8 ciphertexts from registers, the digits come as shared-memory
// broadcasts.
#include "kernels.cuh"

namespace fhestr {

constexpr int kKsBT = 8;         // ciphertexts per CTA
constexpr int kKsThreads = 256;
constexpr int kKsCols = 3;       // columns per thread (3 * 256 >= n + 1 for n <= 767)

__device__ __forceinline__ uint32_t pack_digits(u64 x, int base_log, int level) {
    // A.4: closest representable on base_log*level bits, then balanced digits, least significant
    // level first; stored unsigned (digit + 2^(base_log-1)) in fields of base_log+1 bits, level 1 in
    // the lowest field.
    const int rep = base_log * level;
    u64 state = ((x >> (64 - rep - 1)) + 1) >> 1;
    state &= (1ull << rep) - 1;
    const u64 mask = (1ull << base_log) - 1;
    const int fw = base_log + 1;
    uint32_t packed = 0;
    for (int lvl = level; lvl >= 1; lvl--) {
        const u64 d = state & mask;
        state >>= base_log;
        const u64 carry = (((d - 1) | state) & d) >> (base_log - 1);
        state += carry;
        const int digit = (int)d - (int)(carry << base_log);
        packed |= (uint32_t)(digit + (1 << (base_log - 1))) << (fw * (lvl - 1));
    }
    return packed;
}

__device__ __forceinline__ u64 job_lincomb(const fhestr_job& j, const u64* arena, int idx) {
    u64 x = 0;
    for (uint32_t t = 0; t < j.n_terms; t++)
        x += (u64)(i64)j.coeff[t] * arena[(size_t)j.src[t] * (kN + 1) + idx];
    return x;
}

// SPLIT = false: one CTA walks all N mask rows of its 8 ciphertexts and stores the result.
// SPLIT = true : blockIdx.y selects one of gridDim.y row ranges; partial sums are combined with 64-bit atomic
//                adds into a zeroed ks_out (wrapping adds commute: bit-identical to the unsplit kernel).  Used for
//                small dependency levels, where a single CTA per 8 ciphertexts would stream the whole 58 MiB key
//                through one SM at L2 latency (measured 4.4-5.1 ms for any batch below 2368).
template <int L, bool SPLIT>
__global__ void __launch_bounds__(kKsThreads) keyswitch_kernel(KsBatchArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t* dig = reinterpret_cast<uint32_t*>(smem);                  // [i][BT] packed digits
    u64* body = reinterpret_cast<u64*>(smem + (size_t)(SPLIT ? kN / (int)gridDim.y : kN) * kKsBT * 4);  // [BT]
    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * kKsBT;
    const int nb = min(kKsBT, A.B - b0);
    const int level = (L > 0) ? L : A.level;
    const int fw = A.base_log + 1;
    const uint32_t fmask = (1u << fw) - 1;

    const int rows = SPLIT ? kN / (int)gridDim.y : kN;       // mask rows handled by this CTA
    const int row0 = SPLIT ? (int)blockIdx.y * rows : 0;
    // phase 1: digits of the (never materialised) linear combination
    for (int idx = tid; idx < rows * kKsBT; idx += kKsThreads) {
        const int b = idx / rows, i = row0 + idx - b * rows;
        uint32_t packed = 0;
        if (b < nb) packed = pack_digits(job_lincomb(A.jobs[b0 + b], A.arena, i), A.base_log, level);
        else {  // neutral digits (value 0) for the padding rows
            for (int l = 0; l < level; l++) packed |= (1u << (A.base_log - 1)) << (fw * l);
        }
        dig[(i - row0) * kKsBT + b] = packed;
    }
    if (tid < kKsBT) {
        u64 v = 0;
        if (tid < nb) {
            const fhestr_job& j = A.jobs[b0 + tid];
            v = job_lincomb(j, A.arena, kN) + j.constant;
        }
        body[tid] = v;
    }
    __syncthreads();

    // phase 2: unsigned digit x KSK contraction
    const int ncol = A.n + 1;
    int col[kKsCols];
    bool live[kKsCols];
#pragma unroll
    for (int c = 0; c < kKsCols; c++) {
        col[c] = tid + c * kKsThreads;
        live[c] = col[c] < ncol;
        if (!live[c]) col[c] = 0;
    }
    u64 acc[kKsBT][kKsCols];
#pragma unroll
    for (int b = 0; b < kKsBT; b++)
#pragma unroll
        for (int c = 0; c < kKsCols; c++) acc[b][c] = 0;

    const u64* krow = A.ksk + (size_t)row0 * level * ncol;
#pragma unroll 1
    for (int i = 0; i < rows; i++) {
        uint32_t d[kKsBT];
        const uint4 d0 = *reinterpret_cast<const uint4*>(dig + i * kKsBT);
        const uint4 d1 = *reinterpret_cast<const uint4*>(dig + i * kKsBT + 4);
        d[0] = d0.x; d[1] = d0.y; d[2] = d0.z; d[3] = d0.w;
        d[4] = d1.x; d[5] = d1.y; d[6] = d1.z; d[7] = d1.w;
#pragma unroll
        for (int l = 0; l < ((L > 0) ? L : 8); l++) {
            if (L == 0 && l >= level) break;
            u64 kv[kKsCols];
#pragma unroll
            for (int c = 0; c < kKsCols; c++) kv[c] = __ldg(krow + col[c]);
            krow += ncol;
#pragma unroll
            for (int b = 0; b < kKsBT; b++) {
                const u64 dd = (d[b] >> (fw * l)) & fmask;
#pragma unroll
                for (int c = 0; c < kKsCols; c++) acc[b][c] += dd * kv[c];
            }
        }
    }
#pragma unroll
    for (int c = 0; c < kKsCols; c++) {
        if (!live[c]) continue;
        const u64 corr = A.ksk_corr[col[c]];
#pragma unroll
        for (int b = 0; b < kKsBT; b++) {
            if (b >= nb) continue;
            const u64 init = (col[c] == A.n) ? body[b] : 0ull;
            if (!SPLIT) A.ks_out[(size_t)(b0 + b) * ncol + col[c]] = init - acc[b][c] + corr;
            else {
                const u64 part = (blockIdx.y == 0 ? init + corr : 0ull) - acc[b][c];
                atomicAdd(reinterpret_cast<unsigned long long*>(A.ks_out + (size_t)(b0 + b) * ncol + col[c]), part);
            }
        }
    }
}

constexpr size_t kKsSmemBytes = (size_t)kN * kKsBT * 4 + kKsBT * 8;

cudaError_t keyswitch_configure() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(keyswitch_kernel<5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kKsSmemBytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(keyswitch_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kKsSmemBytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(keyswitch_kernel<5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kKsSmemBytes)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(keyswitch_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kKsSmemBytes);
}

int launch_keyswitch(const KsBatchArgs& a, cudaStream_t s) {
    if (a.B <= 0) return 0;
    const int grid = (a.B + kKsBT - 1) / kKsBT;
    // row splits: keep about 4 CTAs per SM in flight when the level is small (powers of two up to 64)
    int split = 1;
    while (split < 64 && grid * split * 2 <= 148 * 4) split *= 2;
    if (split == 1) {
        if (a.level == 5) keyswitch_kernel<5, false><<<grid, kKsThreads, kKsSmemBytes, s>>>(a);
        else keyswitch_kernel<0, false><<<grid, kKsThreads, kKsSmemBytes, s>>>(a);
        return 1;
    }
    cudaMemsetAsync(a.ks_out, 0, (size_t)a.B * (a.n + 1) * sizeof(u64), s);
    const dim3 g(grid, split);
    const size_t smem = (size_t)(kN / split) * kKsBT * 4 + kKsBT * 8;
    if (a.level == 5) keyswitch_kernel<5, true><<<g, kKsThreads, smem, s>>>(a);
    else keyswitch_kernel<0, true><<<g, kKsThreads, smem, s>>>(a);
    return 2;
}

// corr[c] = 2^(base_log-1) * sum_rows ksk[row][c]
__global__ void ksk_correction_kernel(const u64* ksk, int rows, int ncol, int base_log, u64* corr) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    u64 s = 0;
    for (int r = 0; r < rows; r++) s += ksk[(size_t)r * ncol + c];
    corr[c] = s << (base_log - 1);
}
int launch_ksk_correction(const u64* ksk, int rows, int n, int base_log, u64* corr, cudaStream_t s) {
    ksk_correction_kernel<<<(n + 1 + 127) / 128, 128, 0, s>>>(ksk, rows, n + 1, base_log, corr);
    return 1;
}

// leveled-only jobs: arena[dst] = sum coeff*src + constant*e_body  (dst may alias a source)
__global__ void linear_kernel(const fhestr_job* jobs, int B, u64* arena) {
    const int b = blockIdx.x;
    if (b >= B) return;
    const fhestr_job j = jobs[b];
    for (int idx = threadIdx.x; idx <= kN; idx += blockDim.x) {
        u64 x = job_lincomb(j, arena, idx);
        if (idx == kN) x += j.constant;
        arena[(size_t)j.dst * (kN + 1) + idx] = x;
    }
}
int launch_linear(const fhestr_job* jobs, int B, u64* arena, cudaStream_t s) {
    if (B <= 0) return 0;
    linear_kernel<<<B, 256, 0, s>>>(jobs, B, arena);
    return 1;
}

}  // namespace fhestr
