// keyswitch_mma.cu -- K0+K1 on the tensor cores: the LWE keyswitch as an exact u8 x u8 -> s32 GEMM.
//
//   out[b][c] = body_b[c == n] - sum_{i, lvl} digit_{b,i,lvl} * KSK[i][lvl][c]          
//
// The decomposed-digit x KSK contraction is GEMM-shaped (M = ciphertexts, K = N * levels = 10 240, N' = 743
// columns), so it belongs on the tensor cores; the CUDA-core u64 kernel (keyswitch.cu) re-streams the 58 MiB key
// once per 8 ciphertexts and is L2-bound at 2.9 TB/s (10.4 ms per 4096 ciphertexts, 14 % of a PBS step).  Limb
// split, all integer and exact:
//   A[b][k]      = unsigned digit d' = d + B/2 in [0, B]            (u8; the -B/2 sum KSK correction vector is
//                                                                   the one keyswitch.cu already uses)
//   B[c*8+l][k]  = byte l of KSK[i][lvl][c], k = i * levels + lvl   (u8, K-major; built once at key load)
//   C[b][c*8+l]  = sum_k A B  <= 10 240 * 8 * 255 < 2^25             (s32, no overflow)
//   out[b][c]    = body - sum_l C[b][c*8+l] << 8l + corr[c]          (mod 2^64: bit-exact with the u64 kernel)
// Kernel 1 writes the digits (linear combination of arena blocks fused in, K0); kernel 2 is a 128 x 128 x 64
// tiled GEMM on legacy IMMA.16832 (mma.sync m16n8k32; measured 573 TMAC/s on B200, this GEMM needs 0.25 TMAC per
// 4096 ciphertexts) with a 4-stage cp.async pipeline, XOR-swizzled shared memory and ldmatrix fragment loads;
// the epilogue recombines the 8 limbs of a column inside a quad with two shuffles.
#include "kernels.cuh"

namespace fhestr {

constexpr int kGemmBM = 128, kGemmBN = 128, kGemmBK = 64, kGemmStages = 4, kGemmThreads = 256;
constexpr int kGemmTileBytes = kGemmBM * kGemmBK;  // 8 KiB (A and B tiles have the same shape)
constexpr int kGemmSmemBytes = kGemmStages * 2 * kGemmTileBytes;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// rows of 64 bytes = 4 chunks of 16; chunk index XOR (row >> 1) & 3 makes the 8 row addresses of an ldmatrix hit
// 8 distinct 16-byte bank groups
__device__ __forceinline__ int swz(int row, int chunk) { return row * kGemmBK + ((chunk ^ ((row >> 1) & 3)) << 4); }

// ---- kernel 1: digits.  One CTA per ciphertext.
__global__ void __launch_bounds__(256) ks_digits_kernel(KsBatchArgs A) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int b = blockIdx.x;
    const fhestr_job& j = A.jobs[b];
    const int level = A.level, base_log = A.base_log;
    const int K = kN * level;
    for (int i = threadIdx.x; i < kN; i += blockDim.x) {
        u64 x = 0;
        for (uint32_t t = 0; t < j.n_terms; t++) x += (u64)(i64)j.coeff[t] * A.arena[(size_t)j.src[t] * (kN + 1) + i];
        // A.4: closest representable on base_log*level bits, balanced digits, least significant level first
        const int rep = base_log * level;
        u64 state = ((x >> (64 - rep - 1)) + 1) >> 1;
        state &= (1ull << rep) - 1;
        const u64 mask = (1ull << base_log) - 1;
        for (int lvl = level; lvl >= 1; lvl--) {
            const u64 d = state & mask;
            state >>= base_log;
            const u64 carry = (((d - 1) | state) & d) >> (base_log - 1);
            state += carry;
            const int digit = (int)d - (int)(carry << base_log);
            sm[i * level + (lvl - 1)] = (unsigned char)(digit + (1 << (base_log - 1)));
        }
    }
    if (threadIdx.x == 0) {
        u64 x = j.constant;
        for (uint32_t t = 0; t < j.n_terms; t++) x += (u64)(i64)j.coeff[t] * A.arena[(size_t)j.src[t] * (kN + 1) + kN];
        A.ks_body[b] = x;
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(A.ks_digits + (size_t)b * K);
    const uint4* src = reinterpret_cast<const uint4*>(sm);
    for (int i = threadIdx.x; i < K / 16; i += blockDim.x) dst[i] = src[i];
}

// ---- kernel 2: C = A * B^T on IMMA, limb recombination in the epilogue
__global__ void __launch_bounds__(kGemmThreads, 2) ks_gemm_kernel(KsBatchArgs A) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;          // 2 x 4 warps: 64 x 32 of the 128 x 128 tile each
    const int K = kN * A.level;
    const int m0 = blockIdx.x * kGemmBM, n0 = blockIdx.y * kGemmBN;
    const unsigned char* gA = A.ks_digits + (size_t)m0 * K;
    const unsigned char* gB = A.ksk8 + (size_t)n0 * K;

    // each thread copies 2 chunks of A and 2 of B per stage: chunk id = tid + 256 * h -> row = id >> 2, chunk = id & 3
    auto load_stage = [&](int stage, int kt) {
        unsigned char* sA = sm + stage * 2 * kGemmTileBytes;
        unsigned char* sB = sA + kGemmTileBytes;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int id = tid + 256 * h, row = id >> 2, ch = id & 3;
            const size_t goff = (size_t)row * K + (size_t)kt * kGemmBK + ch * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sA + swz(row, ch))), "l"(gA + goff));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sB + swz(row, ch))), "l"(gB + goff));
        }
    };

    int acc[4][4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[a][b][c] = 0;

    const int KT = K / kGemmBK;
#pragma unroll
    for (int s = 0; s < kGemmStages - 1; s++) {
        if (s < KT) load_stage(s, s);
        asm volatile("cp.async.commit_group;");
    }
    for (int kt = 0; kt < KT; kt++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kGemmStages - 2));
        __syncthreads();
        {   // prefetch the tile that reuses the buffer everybody has just finished reading
            const int nk = kt + kGemmStages - 1;
            if (nk < KT) load_stage(nk % kGemmStages, nk);
            asm volatile("cp.async.commit_group;");
        }
        const unsigned char* sA = sm + (kt % kGemmStages) * 2 * kGemmTileBytes;
        const unsigned char* sB = sA + kGemmTileBytes;
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {                // two k32 steps per 64-byte tile
            uint32_t af[4][4], bf[4][2];
#pragma unroll
            for (int mi = 0; mi < 4; mi++) {
                // ldmatrix.x4: lanes 0-7 rows 0-7 bytes 0-15, 8-15 rows 8-15 bytes 0-15, 16-23 rows 0-7 bytes 16-31, 24-31 rows 8-15 bytes 16-31
                const int row = wm * 64 + mi * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int ch = ks * 2 + (lane >> 4);
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(af[mi][0]), "=r"(af[mi][1]), "=r"(af[mi][2]), "=r"(af[mi][3])
                             : "r"(smem_u32(sA + swz(row, ch))));
            }
#pragma unroll
            for (int np = 0; np < 2; np++) {
                // two n8 tiles per ldmatrix.x4: lanes 0-7 n 0-7 bytes 0-15, 8-15 n 0-7 bytes 16-31, 16-23 n 8-15 bytes 0-15, 24-31 n 8-15 bytes 16-31
                const int row = wn * 32 + np * 16 + (lane & 7) + (lane >> 4) * 8;
                const int ch = ks * 2 + ((lane >> 3) & 1);
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(bf[np * 2][0]), "=r"(bf[np * 2][1]), "=r"(bf[np * 2 + 1][0]), "=r"(bf[np * 2 + 1][1])
                             : "r"(smem_u32(sB + swz(row, ch))));
            }
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                for (int ni = 0; ni < 4; ni++)
                    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+r"(acc[mi][ni][0]), "+r"(acc[mi][ni][1]), "+r"(acc[mi][ni][2]), "+r"(acc[mi][ni][3])
                                 : "r"(af[mi][0]), "r"(af[mi][1]), "r"(af[mi][2]), "r"(af[mi][3]), "r"(bf[ni][0]), "r"(bf[ni][1]));
        }
    }
    asm volatile("cp.async.wait_group 0;");

    // epilogue: an n8 tile is the 8 limbs of one column; thread (g, t) holds limbs 2t, 2t+1 of rows g and g+8
    const int g = lane >> 2, t = lane & 3;
    const int ncol = A.n + 1;
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            const int col = (n0 + wn * 32 + ni * 8) >> 3;
            u64 v0 = ((u64)(uint32_t)acc[mi][ni][0] << (16 * t)) + ((u64)(uint32_t)acc[mi][ni][1] << (16 * t + 8));
            u64 v1 = ((u64)(uint32_t)acc[mi][ni][2] << (16 * t)) + ((u64)(uint32_t)acc[mi][ni][3] << (16 * t + 8));
            v0 += __shfl_xor_sync(0xffffffffu, v0, 1); v0 += __shfl_xor_sync(0xffffffffu, v0, 2);
            v1 += __shfl_xor_sync(0xffffffffu, v1, 1); v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
            if (t == 0 && col < ncol) {
                const u64 corr = A.ksk_corr[col];
                const int r0 = m0 + wm * 64 + mi * 16 + g, r1 = r0 + 8;
                if (r0 < A.B) A.ks_out[(size_t)r0 * ncol + col] = ((col == A.n) ? A.ks_body[r0] : 0ull) - v0 + corr;
                if (r1 < A.B) A.ks_out[(size_t)r1 * ncol + col] = ((col == A.n) ? A.ks_body[r1] : 0ull) - v1 + corr;
            }
        }
}

// ---- key conversion: KSK [N*levels][n+1] u64 -> limb matrix [cols_padded * 8][N*levels] u8 (K-major)
__global__ void ksk_limbs_kernel(const u64* ksk, int K, int ncol, int cols_padded, unsigned char* out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over cols_padded * K, k fastest
    if (idx >= (size_t)cols_padded * K) return;
    const int k = (int)(idx % K), c = (int)(idx / K);
    const u64 v = c < ncol ? ksk[(size_t)k * ncol + c] : 0ull;
#pragma unroll
    for (int l = 0; l < 8; l++) out[((size_t)c * 8 + l) * K + k] = (unsigned char)(v >> (8 * l));
}

int ks_cols_padded(int n) { return ((n + 1) * 8 + kGemmBN - 1) / kGemmBN * kGemmBN / 8; }
size_t ks_digit_rows(size_t B) { return (B + kGemmBM - 1) / kGemmBM * kGemmBM; }

int launch_ksk_limbs(const u64* ksk, int K, int n, unsigned char* out, cudaStream_t s) {
    const int cp = ks_cols_padded(n);
    const size_t total = (size_t)cp * K;
    ksk_limbs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(ksk, K, n + 1, cp, out);
    return 1;
}

cudaError_t keyswitch_mma_configure() {
    cudaError_t e = cudaFuncSetAttribute(ks_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(ks_digits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kN * 8);
}

int launch_keyswitch_mma(const KsBatchArgs& a, cudaStream_t s) {
    if (a.B <= 0) return 0;
    const int K = kN * a.level;
    ks_digits_kernel<<<a.B, 256, K, s>>>(a);
    const dim3 grid((unsigned)(ks_digit_rows(a.B) / kGemmBM), (unsigned)(ks_cols_padded(a.n) * 8 / kGemmBN));
    ks_gemm_kernel<<<grid, kGemmThreads, kGemmSmemBytes, s>>>(a);
    return 2;
}

}  // namespace fhestr
