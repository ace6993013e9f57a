// blind_rotate_wide.cu -- K2+K3+K4 for sm_100a, latency form: ONE PBS per 128-thread CTA, one CTA per SM, used for the
// small dependency levels (fewer jobs than SMs) where the serial chain of 742 CMUX steps is the whole cost.
// The per-thread program is br_wide.cuh.
//
// Shared memory of a CTA (about 221 KiB of the 227 KiB an sm_100 CTA may have):
//   [mod-switched mask 2 KiB | 2 mbarriers | pad]  up to the next 8 KiB-aligned SHARED address
//   [accumulator: 2 polynomials x 2048 words on the 32-bit torus = 16 KiB, each on an 8 KiB boundary (one-LOP3 gather)]
//   [exchange buffer 0: 36 864 B][exchange buffer 1: 34 816 B]
//   [Fourier-key tile ring: 2 x 64 KiB, filled by cp.async.bulk (bulk TMA, SASS: UBLKCP) + mbarrier complete_tx]
// Thread 0 issues the copy of step i+1's tile at the top of step i; every thread waits on the tile's mbarrier
// (parity = use count of that ring slot) right before the GGSW product, so the L2 latency of the key never sits on
// the critical path.  A generic-proxy -> async-proxy fence orders the previous readers of a ring slot (two barriers
// earlier) before the copy that overwrites it.
#include "kernels.cuh"

namespace fhestr {

constexpr int kWAtildeBytes = 2048;
constexpr int kWMbarOff = kWAtildeBytes;          // two 8-byte mbarriers behind the mask
constexpr int kWHeadBytes = kWAtildeBytes + 64;
constexpr int kWAccBytes = 2 * kN * (int)sizeof(acc_t);
constexpr int kWBuf0Bytes = kWBuf0 * (int)sizeof(cplx);
constexpr int kWBuf1Bytes = kWBuf1 * (int)sizeof(cplx);
constexpr int kWKeyBytes = kWKeyTile * (int)sizeof(cplx);   // 65 536
constexpr int kWSharedBase = 0x400;               // the CTA's dynamic window starts here (1 KiB reserved); checked at run time
constexpr int kWPad = (8192 - ((kWSharedBase + kWHeadBytes) & 8191)) & 8191;
constexpr int kWSmemBytes = kWHeadBytes + kWPad + kWAccBytes + kWBuf0Bytes + kWBuf1Bytes + 2 * kWKeyBytes;
static_assert(kWSmemBytes <= 227 * 1024, "one CTA's shared memory");

struct DevWideCtx {
    static constexpr bool kPipelined = true;    // br_wide.cuh: wide_cmux_step_pipe, one named barrier per polynomial
    __device__ __forceinline__ void sync_poly(int p) { asm volatile("bar.sync %0, 128;" ::"r"(p + 1) : "memory"); }
    int tid_;
    acc_t* acc_;
    uint32_t acc_s_;        // shared-space address of polynomial 0's accumulator, 8 KiB aligned (polynomial 1: + 8 KiB)
    cplx* buf0_;
    cplx* buf1_;
    uint16_t* atilde_;
    const cplx* key_smem_;  // ring of two tiles
    uint32_t key_s_;        // shared-space address of the ring
    uint32_t mbar_s_;       // shared-space address of the two mbarriers
    const cplx* bsk_;       // global: [n][kWKeyTile]
    __device__ __forceinline__ int tid() const { return tid_; }
    __device__ __forceinline__ acc_t* acc(int p) { return acc_ + p * kN; }
    __device__ __forceinline__ cplx* buf0() { return buf0_; }
    __device__ __forceinline__ cplx* buf1() { return buf1_; }
    __device__ __forceinline__ uint16_t* atilde() { return atilde_; }
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ acc_t acc_ld_rot(int p, uint32_t x) const {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((acc_s_ + (uint32_t)p * 8192u) | (x & 0x1ffcu)) : "memory");
        return (x & 0x2000u) ? 0u - v : v;
    }
    // one elected thread: arm the slot's mbarrier with the tile size and start the bulk copies (4 x 16 KiB)
    __device__ __forceinline__ void pre_write_sync() {}
    __device__ __forceinline__ void key_release(int) {}
    __device__ __forceinline__ void key_prefetch_current(int) {}
    __device__ __forceinline__ void key_prefetch(int step) {
        if (tid_ != 0) return;
        const uint32_t slot = (uint32_t)step & 1u;
        const uint32_t bar = mbar_s_ + 8u * slot;
        const uint32_t dst = key_s_ + slot * (uint32_t)kWKeyBytes;
        const char* src = reinterpret_cast<const char*>(bsk_ + (size_t)step * kWKeyTile);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)kWKeyBytes) : "memory");
#pragma unroll
        for (int q = 0; q < 4; q++)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst + q * 16384u), "l"(src + q * 16384), "r"(16384u), "r"(bar) : "memory");
    }
    __device__ __forceinline__ const cplx* key_wait(int step) {
        const uint32_t slot = (uint32_t)step & 1u;
        const uint32_t bar = mbar_s_ + 8u * slot;
        const uint32_t parity = ((uint32_t)step >> 1) & 1u;
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "WIDE_WAIT:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
            "@P1 bra WIDE_DONE;\n\t"
            "bra WIDE_WAIT;\n\t"
            "WIDE_DONE:\n\t"
            "}" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
        return key_smem_ + (size_t)slot * kWKeyTile;
    }
};

__device__ __forceinline__ WideConsts load_consts(const WideConsts* tab, int t) {
    WideConsts K;
    const double2* src = reinterpret_cast<const double2*>(tab + t);
    double2* dst = reinterpret_cast<double2*>(&K);
#pragma unroll
    for (int i = 0; i < 18; i++) dst[i] = __ldg(src + i);
    return K;
}

__global__ void __launch_bounds__(kWT, 1) blind_rotate_wide_kernel(BrBatchArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int b = blockIdx.x;
    if (b >= A.B) return;
    const int t = threadIdx.x;
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t pad = (8192u - ((s0 + kWHeadBytes) & 8191u)) & 8191u;
    uint32_t dyn;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if (kWHeadBytes + pad + kWAccBytes + kWBuf0Bytes + kWBuf1Bytes + 2 * kWKeyBytes > dyn) __trap();   // loud, never a wrong result
    DevWideCtx c;
    c.tid_ = t;
    c.atilde_ = reinterpret_cast<uint16_t*>(smem);
    c.mbar_s_ = s0 + kWMbarOff;
    unsigned char* base = smem + kWHeadBytes + pad;
    c.acc_ = reinterpret_cast<acc_t*>(base);
    c.acc_s_ = s0 + kWHeadBytes + pad;
    c.buf0_ = reinterpret_cast<cplx*>(base + kWAccBytes);
    c.buf1_ = reinterpret_cast<cplx*>(base + kWAccBytes + kWBuf0Bytes);
    c.key_smem_ = reinterpret_cast<const cplx*>(base + kWAccBytes + kWBuf0Bytes + kWBuf1Bytes);
    c.key_s_ = c.acc_s_ + kWAccBytes + kWBuf0Bytes + kWBuf1Bytes;
    c.bsk_ = A.bsk_w;
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(c.mbar_s_) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(c.mbar_s_ + 8u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    BrJobView job;
    job.n = A.n;
    job.ks = A.ks + (size_t)b * (A.n + 1);
    const int lut = A.jobs ? A.jobs[b].lut : A.lut_ids[b];
    job.lut = A.luts + (size_t)lut * kN;
    job.post = A.lut_post ? A.lut_post[lut] : 0;
    job.init_acc = A.init_acc ? A.init_acc + (size_t)b * 2 * kN : nullptr;
    job.out_acc = A.out_acc ? A.out_acc + (size_t)b * 2 * kN : nullptr;
    job.out_lwe = A.jobs ? A.arena + (size_t)A.jobs[b].dst * (kN + 1) : nullptr;
    if (A.jobs) {
        job.n_peers = A.n_peers;
        for (int r = 0; r < A.n_peers; r++) job.out_lwe_peer[r] = A.peer_arena[r] + (size_t)A.jobs[b].dst * (kN + 1);
    }
    const WideConsts K = load_consts(A.wide_tab, t);
    wide_thread_main(c, job, K);
}

// ---------------------------------------------------------------------------------------------------------------
// Pair form: TWO PBS per 256-thread CTA (warps 0-3 = PBS A, warps 4-7 = PBS B: every SM sub-partition holds one warp
// of each), one CTA per SM.  Both PBS are at the SAME CMUX step and share ONE 64 KiB key tile in shared memory, filled
// by bulk TMA once per step for both (half the L2 traffic per PBS); everything else is private to a PBS, so the two
// drift inside a step and hide each other's latencies.  Measured (profiles/r2_level_latency.md): 296 PBS in 3.55 ms
// against 4.25 ms for two waves of the single form and 5.0 ms on the throughput kernel; at full load it ends where
// the throughput kernel does (83 k against 86 k PBS/s): all three shapes saturate the shared-memory pipe at 67-68 %.
// Shared memory: per PBS the accumulator (16 KiB) and ONE exchange buffer (36 KiB: a barrier before each rewrite
// instead of the second buffer), one key tile (64 KiB), two mbarriers:
//   full      the tile of this step has landed (TMA complete_tx), parity = step & 1
//   consumed  every thread of both PBS has read its key words of this step (256 arrivals), parity = step & 1
// Thread 0 waits for `consumed` of step i at the top of step i+1 and then starts the copy of tile i+1, which has the
// gather and the whole forward transform (half a step) to land.
constexpr int kW2BufBytes = (kWBuf0Bytes > kWBuf1Bytes ? kWBuf0Bytes : kWBuf1Bytes);
constexpr int kW2PbsBytes = kWAccBytes + kW2BufBytes;                       // 53 248
constexpr int kW2HeadBytes = 2 * kWAtildeBytes + 64;
constexpr int kW2SmemBytes = kW2HeadBytes + 8192 + 2 * kW2PbsBytes + kWKeyBytes;
static_assert(kW2SmemBytes <= 227 * 1024, "one CTA's shared memory");

struct DevWide2Ctx {
    static constexpr bool kPipelined = false;   // one exchange buffer per PBS: the plain step with its extra barriers
    int tid_, grp_, n_groups_;
    acc_t* acc_;
    uint32_t acc_s_;
    cplx* buf_;
    uint16_t* atilde_;
    const cplx* key_smem_;
    uint32_t key_s_, full_s_, consumed_s_;
    const cplx* bsk_;
    __device__ __forceinline__ int tid() const { return tid_; }
    __device__ __forceinline__ acc_t* acc(int p) { return acc_ + p * kN; }
    __device__ __forceinline__ cplx* buf0() { return buf_; }
    __device__ __forceinline__ cplx* buf1() { return buf_; }
    __device__ __forceinline__ uint16_t* atilde() { return atilde_; }
    __device__ __forceinline__ void sync() { asm volatile("bar.sync %0, 128;" ::"r"(grp_ + 1) : "memory"); }
    __device__ __forceinline__ void pre_write_sync() { sync(); }
    __device__ __forceinline__ acc_t acc_ld_rot(int p, uint32_t x) const {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((acc_s_ + (uint32_t)p * 8192u) | (x & 0x1ffcu)) : "memory");
        return (x & 0x2000u) ? 0u - v : v;
    }
    static __device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "W2_WAIT:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
            "@P1 bra W2_DONE;\n\t"
            "bra W2_WAIT;\n\t"
            "W2_DONE:\n\t"
            "}" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    }
    __device__ __forceinline__ void key_prefetch(int) {}
    __device__ __forceinline__ void key_prefetch_current(int step) {
        if (grp_ != 0 || tid_ != 0) return;
        if (step > 0) mbar_wait(consumed_s_, ((uint32_t)(step - 1)) & 1u);   // both PBS are done with the previous tile
        const char* src = reinterpret_cast<const char*>(bsk_ + (size_t)step * kWKeyTile);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_s_), "r"((uint32_t)kWKeyBytes) : "memory");
#pragma unroll
        for (int q = 0; q < 4; q++)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(key_s_ + q * 16384u), "l"(src + q * 16384), "r"(16384u), "r"(full_s_) : "memory");
    }
    __device__ __forceinline__ const cplx* key_wait(int step) {
        mbar_wait(full_s_, (uint32_t)step & 1u);
        return key_smem_;
    }
    __device__ __forceinline__ void key_release(int) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(consumed_s_) : "memory");
    }
};

__global__ void __launch_bounds__(2 * kWT, 1) blind_rotate_wide2_kernel(BrBatchArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int grp = threadIdx.x >> 7, t = threadIdx.x & (kWT - 1);
    const int n_groups = (2 * (int)blockIdx.x + 1 < A.B) ? 2 : 1;
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t pad = (8192u - ((s0 + kW2HeadBytes) & 8191u)) & 8191u;
    uint32_t dyn;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if (kW2HeadBytes + pad + 2 * kW2PbsBytes + kWKeyBytes > dyn) __trap();
    DevWide2Ctx c;
    c.tid_ = t; c.grp_ = grp; c.n_groups_ = n_groups;
    c.atilde_ = reinterpret_cast<uint16_t*>(smem + grp * kWAtildeBytes);
    c.full_s_ = s0 + 2 * kWAtildeBytes;
    c.consumed_s_ = c.full_s_ + 8u;
    unsigned char* base = smem + kW2HeadBytes + pad;            // 8 KiB-aligned shared address
    // [acc A 16K][acc B 16K][buf A][buf B][key tile]: both accumulators keep the 8 KiB alignment of their polynomials
    c.acc_ = reinterpret_cast<acc_t*>(base + grp * kWAccBytes);
    c.acc_s_ = s0 + kW2HeadBytes + pad + grp * kWAccBytes;
    c.buf_ = reinterpret_cast<cplx*>(base + 2 * kWAccBytes + grp * kW2BufBytes);
    c.key_smem_ = reinterpret_cast<const cplx*>(base + 2 * kWAccBytes + 2 * kW2BufBytes);
    c.key_s_ = s0 + kW2HeadBytes + pad + 2 * kWAccBytes + 2 * kW2BufBytes;
    c.bsk_ = A.bsk_w;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(c.full_s_) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(c.consumed_s_), "r"((uint32_t)(n_groups * kWT)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int b = 2 * blockIdx.x + grp;
    if (b >= A.B) return;       // the odd PBS out: its group leaves; `consumed` was armed for one group only

    BrJobView job;
    job.n = A.n;
    job.ks = A.ks + (size_t)b * (A.n + 1);
    const int lut = A.jobs ? A.jobs[b].lut : A.lut_ids[b];
    job.lut = A.luts + (size_t)lut * kN;
    job.post = A.lut_post ? A.lut_post[lut] : 0;
    job.init_acc = A.init_acc ? A.init_acc + (size_t)b * 2 * kN : nullptr;
    job.out_acc = A.out_acc ? A.out_acc + (size_t)b * 2 * kN : nullptr;
    job.out_lwe = A.jobs ? A.arena + (size_t)A.jobs[b].dst * (kN + 1) : nullptr;
    if (A.jobs) {
        job.n_peers = A.n_peers;
        for (int r = 0; r < A.n_peers; r++) job.out_lwe_peer[r] = A.peer_arena[r] + (size_t)A.jobs[b].dst * (kN + 1);
    }
    const WideConsts K = load_consts(A.wide_tab, t);
    wide_thread_main(c, job, K);
}

int launch_blind_rotate_wide2(const BrBatchArgs& a, cudaStream_t s) {
    if (a.B <= 0) return 0;
    blind_rotate_wide2_kernel<<<(a.B + 1) / 2, 2 * kWT, kW2SmemBytes, s>>>(a);
    return 1;
}

// ---- key conversion: one CTA per CMUX step, 128 threads, the four GGSW polynomials one after the other
struct DevWideConvCtx {
    int tid_;
    cplx* buf0_;
    cplx* buf1_;
    __device__ __forceinline__ void pre_write_sync() {}
    __device__ __forceinline__ int tid() const { return tid_; }
    __device__ __forceinline__ cplx* buf0() { return buf0_; }
    __device__ __forceinline__ cplx* buf1() { return buf1_; }
    __device__ __forceinline__ void sync() { __syncthreads(); }
};

__global__ void __launch_bounds__(kWT) bsk_convert_wide_kernel(const u64* bsk_std, const WideConsts* tab, cplx* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    DevWideConvCtx c;
    c.tid_ = threadIdx.x;
    c.buf0_ = reinterpret_cast<cplx*>(smem);
    c.buf1_ = reinterpret_cast<cplx*>(smem + kWBuf0Bytes);
    const WideConsts K = load_consts(tab, threadIdx.x);
    const int i = blockIdx.x;
    for (int row = 0; row < 2; row++)
        for (int col = 0; col < 2; col++)
            wide_bsk_poly_forward(c, bsk_std + (((size_t)i * 2 + row) * 2 + col) * kN, out + (size_t)i * kWKeyTile, row, col, K);
}

cudaError_t blind_rotate_wide_configure() {
    cudaError_t e = cudaFuncSetAttribute(blind_rotate_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(blind_rotate_wide2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kW2SmemBytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(bsk_convert_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWBuf0Bytes + kWBuf1Bytes);
}

int launch_blind_rotate_wide(const BrBatchArgs& a, cudaStream_t s) {
    if (a.B <= 0) return 0;
    blind_rotate_wide_kernel<<<a.B, kWT, kWSmemBytes, s>>>(a);
    return 1;
}

int launch_bsk_convert_wide(const u64* bsk_std, int n, const WideConsts* tab, cplx* out, cudaStream_t s) {
    bsk_convert_wide_kernel<<<n, kWT, kWBuf0Bytes + kWBuf1Bytes, s>>>(bsk_std, tab, out);
    return 1;
}

}  // namespace fhestr
