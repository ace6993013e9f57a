// keyswitch_tc.cu -- K0+K1 on the tensor cores: the LWE keyswitch as an exact u8 x u8 -> s32 GEMM on tcgen05.mma kind::i8.
//
//   out[b][c] = body_b[c == n] - sum_{i, lvl} digit_{b,i,lvl} * KSK[i][lvl][c]          
//
// The decomposed-digit x KSK contraction is GEMM-shaped (M = ciphertexts, K = N * levels = 10 240, N' = 743
// columns), so it belongs on the tensor cores; the CUDA-core u64 kernel (keyswitch.cu) re-streams the 58 MiB key
// once per 8 ciphertexts and is L2-bound at 2.9 TB/s (10.4 ms per 4096 ciphertexts).  Limb split, all integer and exact:
//   A[b][k]      = unsigned digit d' = d + B/2 in [0, B]            (u8; the -B/2 sum KSK correction vector is
//                                                                   the one keyswitch.cu already uses)
//   B[c*8+l][k]  = byte l of KSK[i][lvl][c], k = i * levels + lvl   (u8, K-major; built once at key load)
//   C[b][c*8+l]  = sum_k A B  <= 10 240 * 8 * 255 < 2^25             (s32, no overflow)
//   out[b][c]    = body - sum_l C[b][c*8+l] << 8l + corr[c]          (mod 2^64: bit-exact with the u64 kernel)
// Kernel 1 (ks_digits_kernel) writes the digits (linear combination of arena blocks fused in, K0).  Kernel 2
// (ks_gemm_tc_kernel): one CTA computes a 128 x 256 tile of C (128 ciphertexts x 32 KSK columns):
//   * operands by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle; SASS UTMALDG.2D) into a 4-stage shared-memory ring of
//     128 x 128 B (A) + 256 x 128 B (B) tiles, full / empty mbarriers;
//   * one thread issues tcgen05.mma.cta_group::1.kind::i8 (SASS UTCIMMA; M = 128, N = 256, K = 32 per instruction,
//     four per stage) from shared-memory descriptors; the accumulator is 256 TMEM columns x 128 lanes of s32;
//   * tcgen05.commit (UTCBAR) releases a stage to the producer when its MMAs have read it, and hands the finished
//     accumulator to the epilogue: every warp reads its 32 lanes (tcgen05.ld 32x32b.x32, LDTM: one row per thread, 32
//     columns = the 8 limbs of 4 KSK columns), recombines the limbs in registers -- no shuffles: a row's limbs sit in
//     one thread -- and stores the u64 words.
// Measured, 4096 ciphertexts (profiles/r2b_keyswitch_tc.md): digits + GEMM 0.26 ms against 0.78 ms for round 1's
// mma.sync form (128 x 128 tiles, cp.async + ldmatrix, IMMA.16832; git history), bit-identical output.
#include <cuda.h>

#include "kernels.cuh"

namespace fhestr {

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


constexpr int kTcBM = 128, kTcBN = 256, kTcBK = 128, kTcStages = 4, kTcThreads = 128;
constexpr int kTcABytes = kTcBM * kTcBK, kTcBBytes = kTcBN * kTcBK, kTcStageBytes = kTcABytes + kTcBBytes;
constexpr int kTcBarBytes = 128;                                            // full[4], empty[4], accum, TMEM address
constexpr int kTcSmemBytes = kTcStages * kTcStageBytes + kTcBarBytes + 1024;  // + slack to align the ring to 1 KiB
constexpr int kTcTmemCols = kTcBN;
// instruction descriptor (cute::UMMA::InstrDescriptor): D = s32 (bits 4-5 = 2), A and B unsigned 8 bit (0), both K-major (0),
// N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kTcIdesc = (2u << 4) | ((uint32_t)(kTcBN >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "TC_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra TC_DONE;\n\t"
        "bra TC_WAIT;\n\t"
        "TC_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
// shared-memory matrix descriptor of a K-major tile with 128-byte rows and 128-byte swizzle (cute::UMMA::SmemDescriptor):
// start address >> 4, leading byte offset 1 (unused for swizzled K-major), stride byte offset 1024 >> 4 between 8-row
// groups, version 1 (Blackwell), layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(kTcThreads, 1) ks_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmB, KsBatchArgs A) {
    extern __shared__ unsigned char tc_raw[];
    const uint32_t raw = tc_smem_u32(tc_raw);
    const uint32_t ring = (raw + 1023u) & ~1023u;                      // 1 KiB-aligned: the swizzle atoms are 8 rows x 128 B
    const uint32_t bars = ring + kTcStages * kTcStageBytes;            // full[s] at +8 s, empty[s] at +32 + 8 s, accum at +64
    const uint32_t full0 = bars, empty0 = bars + 32, accum = bars + 64, slot = bars + 72;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * kTcBM, n0 = blockIdx.y * kTcBN;
    const int K = kN * A.level, KT = K / kTcBK;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kTcStages; s++) { tc_mbar_init(full0 + 8 * s, 1); tc_mbar_init(empty0 + 8 * s, 1); }
        tc_mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(kTcTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot) : "memory");

    if (warp == 0 && lane == 0) {
        // producer: one stage = the [m0, +128) x [kt*128, +128) tile of the digits and the [n0, +256) x same tile of the limbs
        for (int kt = 0; kt < KT; kt++) {
            const int s = kt % kTcStages;
            if (kt >= kTcStages) tc_mbar_wait(empty0 + 8 * s, ((kt / kTcStages) - 1) & 1);
            const uint32_t fb = full0 + 8 * s, sa = ring + s * kTcStageBytes, sb = sa + kTcABytes;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"((uint32_t)kTcStageBytes) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(sa), "l"(&tmA), "r"(fb), "r"(kt * kTcBK), "r"(m0) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(sb), "l"(&tmB), "r"(fb), "r"(kt * kTcBK), "r"(n0) : "memory");
        }
    } else if (warp == 1 && lane == 0) {
        // MMA issuer
        for (int kt = 0; kt < KT; kt++) {
            const int s = kt % kTcStages;
            tc_mbar_wait(full0 + 8 * s, (kt / kTcStages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t da = tc_desc(ring + s * kTcStageBytes), db = tc_desc(ring + s * kTcStageBytes + kTcABytes);
#pragma unroll
            for (int k = 0; k < kTcBK / 32; k++) {
                const uint32_t acc_in = (kt > 0 || k > 0) ? 1u : 0u;
                asm volatile(
                    "{\n\t"
                    ".reg .pred p;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
                    "}\n" ::"r"(tmem), "l"(da + 2u * k), "l"(db + 2u * k), "r"(kTcIdesc), "r"(acc_in) : "memory");   // + 32 bytes along K per step
            }
            // the stage may be refilled once these MMAs have read it
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty0 + 8 * s) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(accum) : "memory");
    }

    // epilogue: thread = row 32 warp + lane of the tile; 8 chunks of 32 columns = 4 KSK columns x 8 limbs
    tc_mbar_wait(accum, 0);
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = m0 + warp * 32 + lane;
    const int ncol = A.n + 1;
    const u64 body = row < A.B ? A.ks_body[row] : 0ull;
#pragma unroll 1
    for (int ch = 0; ch < kTcBN / 32; ch++) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
            "tcgen05.wait::ld.sync.aligned;"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ch * 32)) : "memory");
#pragma unroll
        for (int c4 = 0; c4 < 4; c4++) {
            const int col = (n0 + ch * 32 + c4 * 8) >> 3;
            u64 v = 0;
#pragma unroll
            for (int l = 0; l < 8; l++) v += (u64)r[c4 * 8 + l] << (8 * l);
            if (row < A.B && col < ncol) A.ks_out[(size_t)row * ncol + col] = ((col == A.n) ? body : 0ull) - v + A.ksk_corr[col];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTcTmemCols) : "memory");
}

// ---- host side: tensor maps (driver entry point through the runtime, no link against libcuda)
typedef CUresult (*tc_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tc_encode_fn tc_encoder() {
    static tc_encode_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<tc_encode_fn>(p);
    }
    return fn;
}
// [rows][K] u8 row-major, boxes of box_rows x 128 bytes, 128-byte swizzle
static bool tc_make_map(CUtensorMap* m, const void* base, size_t rows, size_t K, uint32_t box_rows) {
    tc_encode_fn enc = tc_encoder();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K};
    const cuuint32_t box[2] = {(cuuint32_t)kTcBK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
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

int ks_cols_padded(int n) { return ((n + 1) * 8 + kTcBN - 1) / kTcBN * kTcBN / 8; }
size_t ks_digit_rows(size_t B) { return (B + kTcBM - 1) / kTcBM * kTcBM; }

int launch_ksk_limbs(const u64* ksk, int K, int n, unsigned char* out, cudaStream_t s) {
    const int cp = ks_cols_padded(n);
    const size_t total = (size_t)cp * K;
    ksk_limbs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(ksk, K, n + 1, cp, out);
    return 1;
}

cudaError_t keyswitch_tc_configure() {
    if (!tc_encoder()) return cudaErrorNotSupported;
    cudaError_t e = cudaFuncSetAttribute(ks_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(ks_digits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kN * 8);
}

// K0+K1: digits, then the GEMM + limb recombination.  Returns the number of kernels launched, 0 when a tensor map
// cannot be encoded (the caller reports the failure: there is no second path).
int launch_keyswitch_tc(const KsBatchArgs& a, cudaStream_t s) {
    if (a.B <= 0) return 0;
    const size_t K = (size_t)kN * a.level;
    const size_t rows = ks_digit_rows((size_t)a.B), limb_rows = (size_t)ks_cols_padded(a.n) * 8;
    CUtensorMap tmA, tmB;
    if (!tc_make_map(&tmA, a.ks_digits, rows, K, kTcBM) || !tc_make_map(&tmB, a.ksk8, limb_rows, K, kTcBN)) return 0;
    ks_digits_kernel<<<a.B, 256, K, s>>>(a);
    const dim3 grid((unsigned)(rows / kTcBM), (unsigned)(limb_rows / kTcBN));
    ks_gemm_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, s>>>(tmA, tmB, a);
    return 2;
}

}  // namespace fhestr
