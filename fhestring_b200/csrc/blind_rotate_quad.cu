// blind_rotate_quad.cu -- K2+K3+K4 for sm_100a, four warps per PBS (br_quad.cuh): modulus switch, blind rotation
// (742 CMUX steps), sample extract, one launch per dependency level.
//
// Mapping: one PBS = one 128-thread CTA = 2 polynomials x 2 warps; 4 CTAs per SM at 128 registers per thread =
// 16 warps per SM, the four warps of a PBS on the four SM sub-partitions.
// Shared memory per PBS: accumulator 2 x 2048 words on the 32-bit torus (16 KiB) + one padded exchange buffer per
// polynomial (2 x 17 408 B) + the mod-switched mask (2 KiB) = 53 248 B.
#include "br_quad.cuh"
#include "kernels.cuh"

namespace fhestr {

constexpr int kQuadAccBytes = 2 * kN * (int)sizeof(acc_t);
constexpr int kQuadExchBytes = kQExchCplx * (int)sizeof(cplx);
constexpr int kQuadSmemBytes = kQuadAccBytes + 2 * kQuadExchBytes + 2048;

struct QuadDevCtx {
    int tau_, poly_;
    acc_t* acc_;
    cplx* exch_;
    cplx* exch_partner_;
    uint16_t* atilde_;
    __device__ __forceinline__ int tau() const { return tau_; }
    __device__ __forceinline__ int poly() const { return poly_; }
    __device__ __forceinline__ acc_t* acc() { return acc_; }
    __device__ __forceinline__ cplx* exch() { return exch_; }
    __device__ __forceinline__ const cplx* exch_partner() { return exch_partner_; }
    __device__ __forceinline__ uint16_t* atilde() { return atilde_; }
    __device__ __forceinline__ void syncwarp() { __syncwarp(); }
    __device__ __forceinline__ void poly_sync() {
        if (poly_ == 0) asm volatile("bar.sync 1, 64;" ::: "memory");
        else asm volatile("bar.sync 2, 64;" ::: "memory");
    }
    __device__ __forceinline__ void cta_sync() { __syncthreads(); }
    __device__ __forceinline__ cplx ldg(const cplx* p) const {
        const double2 v = __ldg(reinterpret_cast<const double2*>(p));
        return cplx{v.x, v.y};
    }
};

__global__ void __launch_bounds__(128, 4) blind_rotate_quad_kernel(BrBatchArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5;
    QuadDevCtx c;
    c.poly_ = warp >> 1;
    c.tau_ = threadIdx.x & 63;
    c.acc_ = reinterpret_cast<acc_t*>(smem) + c.poly_ * kN;
    cplx* ex = reinterpret_cast<cplx*>(smem + kQuadAccBytes);
    c.exch_ = ex + c.poly_ * kQExchCplx;
    c.exch_partner_ = ex + (1 - c.poly_) * kQExchCplx;
    c.atilde_ = reinterpret_cast<uint16_t*>(smem + kQuadAccBytes + 2 * kQuadExchBytes);

    BrJobView job;
    job.n = A.n;
    job.ks = A.ks + (size_t)b * (A.n + 1);
    const int lut = A.jobs ? A.jobs[b].lut : A.lut_ids[b];
    job.lut = A.luts + (size_t)lut * kN;
    job.init_acc = A.init_acc ? A.init_acc + (size_t)b * 2 * kN : nullptr;
    job.out_acc = A.out_acc ? A.out_acc + (size_t)b * 2 * kN : nullptr;
    job.out_lwe = A.jobs ? A.arena + (size_t)A.jobs[b].dst * (kN + 1) : nullptr;
    quad_thread_main(c, job, A.bsk_q, A.qt);
}

cudaError_t blind_rotate_quad_configure() {
    return cudaFuncSetAttribute(blind_rotate_quad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kQuadSmemBytes);
}

int launch_blind_rotate_quad(const BrBatchArgs& a, cudaStream_t s) {
    if (a.B <= 0) return 0;
    blind_rotate_quad_kernel<<<a.B, 128, kQuadSmemBytes, s>>>(a);
    return 1;
}

// ---- K6 for this layout: one 64-thread CTA per GGSW polynomial
struct QuadConvCtx {
    int tau_;
    cplx* exch_;
    __device__ __forceinline__ int tau() const { return tau_; }
    __device__ __forceinline__ cplx* exch() { return exch_; }
    __device__ __forceinline__ void syncwarp() { __syncwarp(); }
    __device__ __forceinline__ void poly_sync() { __syncthreads(); }
    __device__ __forceinline__ cplx ldg(const cplx* p) const {
        const double2 v = __ldg(reinterpret_cast<const double2*>(p));
        return cplx{v.x, v.y};
    }
};
__global__ void __launch_bounds__(64) bsk_convert_quad_kernel(const u64* bsk_std, int n_polys, QuadTables tb, cplx* out) {
    __shared__ __align__(16) cplx ex[kQExchCplx];
    const int q = blockIdx.x;  // polynomial index = (step*2 + row)*2 + col
    if (q >= n_polys) return;
    QuadConvCtx c{(int)threadIdx.x, ex};
    const int step = q >> 2, row = (q >> 1) & 1, col = q & 1;
    quad_bsk_poly_forward(c, bsk_std + (size_t)q * kN, out + (size_t)step * kQBskStepElems, row, col, tb);
}
int launch_bsk_convert_quad(const u64* bsk_std, int n, const QuadTables& tb, cplx* out, cudaStream_t s) {
    bsk_convert_quad_kernel<<<n * 4, 64, 0, s>>>(bsk_std, n * 4, tb, out);
    return 1;
}

}  // namespace fhestr
