// blind_rotate.cu -- K2+K3+K4 for sm_100a: modulus switch, blind rotation (742 CMUX steps), sample
// extract, one launch per dependency level.  The per-thread program is br_core.cuh.
//
// Mapping: one PBS = one pair of warps (mask polynomial, body polynomial); a CTA carries P pairs
// (P = 4 fills the 64K-register file at 255 registers/thread: 8 warps, 2 per SM sub-partition).
// Shared memory per pair: accumulator 2 x 2048 u64 (32 KiB) + two padded transpose buffers
// (2 x 8448 B) + the mod-switched mask (2 KiB) = 51 712 B; 4 pairs = 202 KiB of the 227 KiB.
// The Fourier BSK (46 MiB for n = 742) stays resident in the 126 MB L2 and is read with 16-byte
// read-only loads, one 64 KiB step tile per CMUX.
#include "kernels.cuh"

namespace fhestr {

constexpr int kAtildeBytes = 2048;  // up to 1024 u16
constexpr int kPairSmemBytes = 2 * kN * 8 + 2 * kXbufDoubles * 8 + kAtildeBytes;

struct DevCtx {
    int lane_, poly_, slot_;
    u64* acc_;
    double* xbuf_;
    double* xbuf_partner_;
    uint16_t* atilde_;
    __device__ __forceinline__ int lane() const { return lane_; }
    __device__ __forceinline__ int poly() const { return poly_; }
    __device__ __forceinline__ u64* acc() { return acc_; }
    __device__ __forceinline__ double* xbuf() { return xbuf_; }
    __device__ __forceinline__ double* xbuf_partner() { return xbuf_partner_; }
    __device__ __forceinline__ uint16_t* atilde() { return atilde_; }
    __device__ __forceinline__ void syncwarp() { __syncwarp(); }
    __device__ __forceinline__ void pair_sync() {
        asm volatile("bar.sync %0, 64;" ::"r"(slot_ + 1) : "memory");
    }
    __device__ __forceinline__ cplx ldg(const cplx* p) const {
        const double2 v = __ldg(reinterpret_cast<const double2*>(p));
        return cplx{v.x, v.y};
    }
};

template <int P>
__global__ void __launch_bounds__(64 * P, 1) blind_rotate_kernel(BrBatchArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5;
    const int slot = warp >> 1;
    const int b = blockIdx.x * P + slot;
    if (b >= A.B) return;  // whole pair leaves together; pair barriers are per pair
    unsigned char* base = smem + (size_t)slot * kPairSmemBytes;
    u64* acc = reinterpret_cast<u64*>(base);
    double* xb = reinterpret_cast<double*>(base + 2 * kN * 8);
    DevCtx c;
    c.lane_ = threadIdx.x & 31;
    c.poly_ = warp & 1;
    c.slot_ = slot;
    c.acc_ = acc + c.poly_ * kN;
    c.xbuf_ = xb + c.poly_ * kXbufDoubles;
    c.xbuf_partner_ = xb + (1 - c.poly_) * kXbufDoubles;
    c.atilde_ = reinterpret_cast<uint16_t*>(base + 2 * kN * 8 + 2 * kXbufDoubles * 8);

    BrJobView job;
    job.n = A.n;
    job.ks = A.ks + (size_t)b * (A.n + 1);
    const int lut = A.jobs ? A.jobs[b].lut : A.lut_ids[b];
    job.lut = A.luts + (size_t)lut * kN;
    job.init_acc = A.init_acc ? A.init_acc + (size_t)b * 2 * kN : nullptr;
    job.out_acc = A.out_acc ? A.out_acc + (size_t)b * 2 * kN : nullptr;
    job.out_lwe = A.jobs ? A.arena + (size_t)A.jobs[b].dst * (kN + 1) : nullptr;
    br_thread_main(c, job, A.bsk, A.tf, A.ti);
}

cudaError_t blind_rotate_configure() {
    cudaError_t e;
    e = cudaFuncSetAttribute(blind_rotate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1 * kPairSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(blind_rotate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kPairSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(blind_rotate_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kPairSmemBytes);
    return e;
}

int launch_blind_rotate(const BrBatchArgs& a, int pbs_per_cta, cudaStream_t s) {
    if (a.B <= 0) return 0;
    int P = pbs_per_cta;
    if (P != 1 && P != 2 && P != 4) P = 1;  // measured: independent 64-thread CTAs drift out of phase and overlap FP64 with shared-memory phases; P=4 runs in lockstep and is 1.47x slower
    const int grid = (a.B + P - 1) / P;
    switch (P) {
        case 1: blind_rotate_kernel<1><<<grid, 64, 1 * kPairSmemBytes, s>>>(a); break;
        case 2: blind_rotate_kernel<2><<<grid, 128, 2 * kPairSmemBytes, s>>>(a); break;
        default: blind_rotate_kernel<4><<<grid, 256, 4 * kPairSmemBytes, s>>>(a); break;
    }
    return 1;
}

}  // namespace fhestr
