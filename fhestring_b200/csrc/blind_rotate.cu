// blind_rotate.cu -- K2+K3+K4 for sm_100a: modulus switch, blind rotation (742 CMUX steps), sample
// extract, one launch per dependency level.  The per-thread program is br_core.cuh.
//
// Mapping: one PBS = one pair of warps (mask polynomial, body polynomial), one PBS per 64-thread CTA, 4 CTAs
// per SM (launch bound 4: ptxas may use 255 registers per thread, the rolled step needs 168 and spends the rest on
// loads in flight).
// Shared memory per pair: the mod-switched mask (2 KiB), the accumulator 2 x 2048 words on the 32-bit torus (16 KiB,
// each polynomial on an 8 KiB-aligned shared address: 4 KiB of padding behind the mask) and one padded transpose
// matrix of complex words per warp (2 x 16 896 B) = 56 320 B: 4 CTAs (+ 1 KiB reserved each) take 224 of the SM's
// 228 KB (carve-out 100 %: since the twiddles left for tensor memory the L1 served 11 % of the loads).
// Tensor memory per CTA: 128 columns (4 CTAs = the SM's 512).  Warp w owns TMEM lanes 32 w .. 32 w + 31; thread t keeps
// its 32 inter-pass twiddles tf[k1*32 + t] (128 words) in the 128 columns of its lane: written once per PBS with
// tcgen05.st, read 8 twiddles at a time with tcgen05.ld (SASS STTM / LDTM) -- a lane-private scratch file next to the
// registers that costs no LSU wavefronts (the 64 LDG.128 per warp-step they replace were 17 % of this kernel's).
// The Fourier BSK (46 MiB for n = 742) stays resident in the 126 MB L2 and is read with 16-byte
// read-only loads, one 64 KiB step tile per CMUX.
#include "kernels.cuh"

namespace fhestr {

constexpr int kAtildeBytes = 2048;  // up to 1024 u16
constexpr int kAccBytes = 2 * kN * (int)sizeof(acc_t);  // 16 KiB: both polynomials on the 32-bit torus
// layout: [two transpose matrices][mask 2 KiB][pad][acc0 8 KiB | acc1 8 KiB, each on an 8 KiB-aligned SHARED address].
// The CTA's shared window starts at 0x400 (1 KiB is reserved per CTA), so the pad is 4 KiB: 56 320 B per CTA, and
// 4 x (56 320 + 1 024) = 224 KiB of the SM's 228.  The kernel computes the pad from the real address and traps if the
// dynamic allocation is too small for it.
constexpr int kSharedBase = 0x400;
constexpr int kCtasPerSmForSmem = 4;
constexpr int kXbufBytes = 2 * kWarpXbufDoubles * 8;   // one padded matrix per warp
constexpr int kFrontBytes = kXbufBytes + kAtildeBytes;
constexpr int kPadBytes = (8192 - ((kSharedBase + kFrontBytes) & 8191)) & 8191;
constexpr int kSmemBytes = kFrontBytes + kPadBytes + kAccBytes;  // 56 320 B
static_assert(kCtasPerSmForSmem * (kSmemBytes + 1024) <= 228 * 1024, "four CTAs per SM");
constexpr int kCtasPerSm = 4;
constexpr int kTmemCols = 128;   // = words of a lane's twiddle set; 4 CTAs per SM use all 512 columns
static_assert(kCtasPerSm * kTmemCols <= 512, "tensor memory of one SM");
static_assert(kTwChunks * 32 == kTmemCols, "four 32-column chunks per lane");

struct DevCtx {
    int lane_, poly_;
    acc_t* acc_;
    double* xbuf_;
    double* xbuf_partner_;
    uint16_t* atilde_;
    uint32_t acc_s_;     // shared-space address of this polynomial's accumulator, 8 KiB aligned
    uint32_t tmem_tw_;   // TMEM address of this warp's lanes, column 0 of the CTA's allocation
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
    __device__ __forceinline__ void pair_sync() { asm volatile("bar.sync 1, 64;" ::: "memory"); }
    __device__ __forceinline__ cplx ldg(const cplx* p) const {
        const double2 v = __ldg(reinterpret_cast<const double2*>(p));
        return cplx{v.x, v.y};
    }
    // twiddles in tensor memory: words [32 ch, 32 ch + 32) of this lane's 128 = twiddles 8 ch .. 8 ch + 7
    __device__ __forceinline__ void tw_ld(int ch, uint32_t (&r)[32], const cplx*) const {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(tmem_tw_ + 32u * (uint32_t)ch));
    }
    // the loaded registers may be read only after tcgen05.wait::ld; passing them through the statement keeps the
    // compiler from moving a use above it
    __device__ __forceinline__ void tw_wait(uint32_t (&r)[32]) const {
        asm volatile("tcgen05.wait::ld.sync.aligned;"
            : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
              "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
              "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
              "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]));
    }
    __device__ __forceinline__ void tw_st(int ch, const uint32_t (&r)[32]) const {
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
            :: "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
               "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
               "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
               "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]),
               "r"(tmem_tw_ + 32u * (uint32_t)ch) : "memory");
    }
    static __device__ __forceinline__ double tw_word(uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }
};

__global__ void __launch_bounds__(64, kCtasPerSm) blind_rotate_kernel(BrBatchArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.x;
    if (b >= A.B) return;   // the whole CTA leaves together, before it owns any tensor memory
    DevCtx c;
    c.lane_ = threadIdx.x & 31;
    c.poly_ = warp;
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t pad = (8192u - ((s0 + kFrontBytes) & 8191u)) & 8191u;
    uint32_t dyn;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if (kFrontBytes + pad + kAccBytes > dyn || A.n + 1 > kAtildeBytes / 2) __trap();   // loud, never a wrong result
    double* xb = reinterpret_cast<double*>(smem);
    c.atilde_ = reinterpret_cast<uint16_t*>(smem + kXbufBytes);
    const uint32_t acc_off = kFrontBytes + pad + c.poly_ * kN * (uint32_t)sizeof(acc_t);
    c.acc_ = reinterpret_cast<acc_t*>(smem + acc_off);
    c.acc_s_ = s0 + acc_off;
    c.xbuf_ = xb + c.poly_ * kWarpXbufDoubles;
    c.xbuf_partner_ = xb + (1 - c.poly_) * kWarpXbufDoubles;

    // tensor memory: warp 0 allocates the CTA's columns (the address lands in shared memory: the mask buffer is free
    // until br_thread_main fills it), every thread writes its 32 twiddles into the columns of its own lane
    {
        uint32_t* slot_addr = reinterpret_cast<uint32_t*>(c.atilde_);
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(slot_addr)), "n"(kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tbase = *reinterpret_cast<volatile uint32_t*>(slot_addr);
        c.tmem_tw_ = tbase + ((uint32_t)warp << 21);   // lane field: bits 16 and up, 32 lanes per warp
        __syncthreads();                               // everyone has read the address before the mask is written
#pragma unroll
        for (int ch = 0; ch < kTwChunks; ch++) {
            uint32_t r[32];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const double2 w = __ldg(reinterpret_cast<const double2*>(A.tf + (ch * 8 + j) * 32 + c.lane_));
                r[4 * j] = (uint32_t)__double2loint(w.x); r[4 * j + 1] = (uint32_t)__double2hiint(w.x);
                r[4 * j + 2] = (uint32_t)__double2loint(w.y); r[4 * j + 3] = (uint32_t)__double2hiint(w.y);
            }
            c.tw_st(ch, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }

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
    br_thread_main(c, job, A.bsk, A.tf);

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // both warps are done with their columns
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem_tw_ & 0xffffu), "n"(kTmemCols) : "memory");
}

cudaError_t blind_rotate_configure() {
    cudaError_t e = cudaFuncSetAttribute(blind_rotate_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(blind_rotate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
}

// Shapes measured against this one and not kept (profiles/r1_final2_ab_variants.md, r2_phase_log.md,
// r2_compact_tmem.md; git history): 2 or 4 PBS per CTA (run in lockstep, 1.2-1.5x slower, with or without an FP64
// baton between them), a TMA-filled shared-memory key ring, 5 and 6 CTAs per SM (the rolled step fits 168 registers
// without spills and six CTAs run, but no faster than four), round 1's straight-line step.
int launch_blind_rotate(const BrBatchArgs& a, cudaStream_t s) {
    if (a.B <= 0) return 0;
    blind_rotate_kernel<<<a.B, 64, kSmemBytes, s>>>(a);
    return 1;
}

}  // namespace fhestr
