// blind_rotate.cu -- K2+K3+K4 for sm_100a: modulus switch, blind rotation (742 CMUX steps), sample
// extract, one launch per dependency level.  The per-thread program is br_core.cuh.
//
// Mapping: one PBS = one pair of warps (mask polynomial, body polynomial), one PBS per 64-thread CTA, 4 CTAs
// per SM (255 registers/thread fill the 64K-register file: 8 warps, 2 per SM sub-partition).
// Shared memory per pair: the mod-switched mask (2 KiB), the accumulator 2 x 2048 words on the 32-bit torus (16 KiB,
// each polynomial on an 8 KiB-aligned shared address: 5 KiB of padding behind the mask) and one padded transpose
// matrix per warp (2 x 8448 B) = 40 448 B: 4 CTAs (+ 1 KiB reserved each) still fit the 164 KB carve-out, which
// leaves 92 KB of L1 for the twiddle tables and the BSK tile the 4 CTAs of an SM read at nearly the same time.
// The Fourier BSK (46 MiB for n = 742) stays resident in the 126 MB L2 and is read with 16-byte
// read-only loads, one 64 KiB step tile per CMUX.
#include "kernels.cuh"

namespace fhestr {

constexpr int kAtildeBytes = 2048;  // up to 1024 u16
constexpr int kAccBytes = 2 * kN * (int)sizeof(acc_t);  // 16 KiB: both polynomials on the 32-bit torus
static_assert(FHESTR_BR_SLIM == 1, "the unaligned-accumulator layout of the first round-1 kernel is gone from the tree");
// aligned layout (one PBS per CTA only): [mask 2 KiB][pad][acc0 8 KiB | acc1 8 KiB, each on an 8 KiB-aligned SHARED
// address][two transpose matrices].  The CTA's shared window starts at 0x400 (1 KiB is reserved per CTA), so the
// pad is 5 KiB: 40 448 B per CTA, and 4 x (40 448 + 1 024) = 162 KiB still fits the 164 KiB carve-out.  The kernel
// computes the pad from the real address and traps if the dynamic allocation is too small for it.
constexpr int kSlimSharedBase = 0x400;
constexpr int kSlimPad = (8192 - ((kSlimSharedBase + kAtildeBytes) & 8191)) & 8191;
constexpr int kSlimSmemBytes = kAtildeBytes + kSlimPad + kAccBytes + 2 * kWarpXbufDoubles * 8;  // 40 448 B

struct DevCtx {
    int lane_, poly_, slot_;
    acc_t* acc_;
    double* xbuf_;
    double* xbuf_partner_;
    uint16_t* atilde_;
    uint32_t acc_s_;   // shared-space address of this polynomial's accumulator, 8 KiB aligned
    // word ((x >> 2) mod N) of the accumulator, negated when bit 13 of the byte offset x is set (negacyclic wrap)
    __device__ __forceinline__ acc_t acc_ld_rot(uint32_t x) const {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(acc_s_ | (x & 0x1ffcu)) : "memory");
        return (x & 0x2000u) ? 0u - v : v;
    }
    __device__ __forceinline__ int lane() const { return lane_; }
    __device__ __forceinline__ int poly() const { return poly_; }
    __device__ __forceinline__ acc_t* acc() { return acc_; }
    __device__ __forceinline__ double* xbuf() { return xbuf_; }
    __device__ __forceinline__ double* xbuf_partner() { return xbuf_partner_; }
    __device__ __forceinline__ uint16_t* atilde() { return atilde_; }
    __device__ __forceinline__ void syncwarp() { __syncwarp(); }
    __device__ __forceinline__ void pair_sync() {
        asm volatile("bar.sync %0, 64;" ::"r"(slot_ + 1) : "memory");
    }
    __device__ __forceinline__ void prefetch_l1(const cplx* p) const {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    }
    __device__ __forceinline__ cplx ldg(const cplx* p) const {
        const double2 v = __ldg(reinterpret_cast<const double2*>(p));
        return cplx{v.x, v.y};
    }
};

// P = PBS per CTA, MB = CTAs per SM the register allocation is sized for (launch bound)
template <int P, int MB>
__global__ void __launch_bounds__(64 * P, MB) blind_rotate_kernel(BrBatchArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5;
    const int slot = warp >> 1;
    const int b = blockIdx.x * P + slot;
    if (b >= A.B) return;  // whole pair leaves together; pair barriers are per pair
    DevCtx c;
    c.lane_ = threadIdx.x & 31;
    c.poly_ = warp & 1;
    c.slot_ = slot;
    static_assert(P == 1, "one PBS per CTA");
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t pad = (8192u - ((s0 + kAtildeBytes) & 8191u)) & 8191u;
    uint32_t dyn;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if (kAtildeBytes + pad + kAccBytes + 2 * kWarpXbufDoubles * 8 > dyn) __trap();   // loud, never a wrong result
    unsigned char* base = smem + kAtildeBytes + pad;
    acc_t* acc = reinterpret_cast<acc_t*>(base);
    double* xb = reinterpret_cast<double*>(base + kAccBytes);
    c.acc_ = acc + c.poly_ * kN;
    c.acc_s_ = s0 + kAtildeBytes + pad + c.poly_ * kN * (uint32_t)sizeof(acc_t);
    c.atilde_ = reinterpret_cast<uint16_t*>(smem);
    c.xbuf_ = xb + c.poly_ * kWarpXbufDoubles;
    c.xbuf_partner_ = xb + (1 - c.poly_) * kWarpXbufDoubles;

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
    br_thread_main(c, job, A.bsk, A.tf, A.ti);
}

cudaError_t blind_rotate_configure() {
    return cudaFuncSetAttribute(blind_rotate_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlimSmemBytes);
}

// Measured (r1, 4096 PBS): one PBS per 64-thread CTA, 4 CTAs per SM at 255 registers is the fastest shape; independent
// CTAs drift out of phase and overlap their FP64 and shared-memory phases, while 2 or 4 PBS per CTA ran in lockstep
// (1.3-1.5x slower), and sizing the register allocation for 5-6 CTAs per SM (168 registers) made every warp ~1.6x
// slower for 1.5x the warps (net 0.8x).  Those variants are gone from the tree (git history: round 1).
int launch_blind_rotate(const BrBatchArgs& a, cudaStream_t s) {
    if (a.B <= 0) return 0;
    blind_rotate_kernel<1, 4><<<a.B, 64, kSlimSmemBytes, s>>>(a);
    return 1;
}

}  // namespace fhestr
