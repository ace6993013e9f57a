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
// PBS per CTA.  1: four independent 64-thread CTAs per SM.  4 (FHESTR_BR_QUAD): ONE 256-thread CTA per SM whose PBS
// s and s ^ kPartnerXor sit on the same two SM sub-partitions (warp w runs on sub-partition w mod 4) and hand the FP64
// pipe to each other through named barriers (br_core.cuh: FHESTR_BR_BATON).
#ifndef FHESTR_BR_QUAD
#define FHESTR_BR_QUAD 0
#endif
#ifndef FHESTR_BR_PARTNER_XOR
#define FHESTR_BR_PARTNER_XOR 2
#endif
// start-up stagger (ns) of the second PBS of a baton pair / of the odd pair, baton off: phase offset by delay only
#ifndef FHESTR_BR_STAGGER_NS
#define FHESTR_BR_STAGGER_NS 0
#endif
#ifndef FHESTR_BR_STAGGER2_NS
#define FHESTR_BR_STAGGER2_NS 0
#endif
// FHESTR_BR_PBS = PBS per CTA (1, 2 or 4; FHESTR_BR_QUAD = 4).  With FHESTR_BR_RING the CTA's PBS share one key ring.
#ifndef FHESTR_BR_PBS
#define FHESTR_BR_PBS (FHESTR_BR_QUAD ? 4 : 1)
#endif
// ring slots of 8 KiB (FHESTR_BR_RING): two 2-PBS CTAs per SM leave 40 KiB free; with a power of two that divides the
// 8 chunks of a step, slot and mbarrier parity of a chunk do not depend on the step (no registers, no arithmetic)
#ifndef FHESTR_BR_RING_SLOTS
#define FHESTR_BR_RING_SLOTS 4
#endif
// FHESTR_BR_CTAS: PBS per SM the register allocation is sized for (4: 255 registers per thread; 5: 200)
#ifndef FHESTR_BR_CTAS
#define FHESTR_BR_CTAS 4
#endif
constexpr int kCtasPerSm = FHESTR_BR_CTAS / FHESTR_BR_PBS;
constexpr int kPbsPerCta = FHESTR_BR_PBS;
constexpr int kRingSlots = FHESTR_BR_RING ? FHESTR_BR_RING_SLOTS : 0;
constexpr int kChunkBytes = kKeyChunkElems * (int)sizeof(cplx);     // 8 KiB
constexpr int kPieceBytes = kKeyPieceElems * (int)sizeof(cplx);     // 4 KiB: one GGSW row of a chunk
constexpr int kMbarBytes = 256;                                     // full[slots] + consumed[slots], 8 B each
constexpr int kPartnerXor = FHESTR_BR_PARTNER_XOR;
constexpr int kXbufBytes = 2 * kWarpXbufDoubles * 8;   // per PBS: one padded matrix per warp
// layout: [P x transpose matrices][P x mask][mbarriers][pad][P x (acc0 | acc1)][key ring], the accumulators on 8 KiB-aligned
// SHARED addresses
constexpr bool kAccAligned = FHESTR_BR_CTAS <= 5;   // six CTAs per SM have no room for the alignment pad: the gather adds instead of or-ing
constexpr int kMaskBytes = (FHESTR_BR_RING || !kAccAligned) ? 1536 : kAtildeBytes;    // n + 1 <= 768 mask words when the ring needs the room
constexpr int kSlimFront = kPbsPerCta * (kXbufBytes + kMaskBytes) + (FHESTR_BR_RING ? kMbarBytes : 0);
constexpr int kSlimPad = kAccAligned ? (8192 - ((kSlimSharedBase + kSlimFront) & 8191)) & 8191 : 0;
constexpr int kSlimSmemBytes = kSlimFront + kSlimPad + kPbsPerCta * kAccBytes + kRingSlots * kChunkBytes;  // 39 936 B for one PBS without the ring
static_assert(kSlimSmemBytes <= 227 * 1024, "one CTA's shared memory");

#ifdef FHESTR_BR_PHASELOG
// timing experiment only: clock of every step start (and of the product stage) of the first four PBS on SMs 0..7
constexpr int kLogSms = 8, kLogSteps = 1024;
__device__ unsigned g_log_claim[kLogSms];
constexpr int kLogMarks = 18;
__device__ unsigned long long g_log[kLogSms][4][kLogMarks][kLogSteps];
__device__ unsigned g_log_meta[kLogSms][4][4];
__device__ unsigned long long g_cta_log[8192][4];   // per PBS: smid, globaltimer at start, at loop end, at exit
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif
struct DevCtx {
    int log_slot_ = -1, log_sm_ = 0, log_step_ = 0;
    int log_b_ = -1;
    __device__ __forceinline__ void log_mark(int which) {
#ifdef FHESTR_BR_PHASELOG
        if (which < 0) { if (log_b_ >= 0) g_cta_log[log_b_][2] = gtimer(); return; }
        if (log_slot_ >= 0 && log_step_ < kLogSteps) {
            g_log[log_sm_][log_slot_][which][log_step_] = clock64();
            if (which == 9) log_step_++;
        }
#endif
    }
    int lane_, poly_, slot_;
    // key ring (FHESTR_BR_RING)
    // shared-space addresses of full[slots], consumed[slots] and the ring: compile-time constants (the kernel traps if
    // its dynamic window does not start at kSlimSharedBase), so that the ring costs the loop no registers
    static constexpr uint32_t full_s_ = kSlimSharedBase + kPbsPerCta * (kXbufBytes + kMaskBytes);
    static constexpr uint32_t cons_s_ = full_s_ + 8u * kRingSlots;
    static constexpr uint32_t ring_s_ = kSlimSharedBase + kSlimFront + kSlimPad + kPbsPerCta * kAccBytes;
    int warp_ = 0, n_ = 0;
    const char* bsk_bytes_ = nullptr;
    static __device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "BR_WAIT:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
            "@P1 bra BR_DONE;\n\t"
            "bra BR_WAIT;\n\t"
            "BR_DONE:\n\t"
            "}" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    }
    // chunk cc = step * 8 + j lives in slot cc mod slots during use (cc div slots) of that slot
    static constexpr bool kStaticSlots = kRingSlots > 0 && (kRingSlots & (kRingSlots - 1)) == 0 && kKeyChunks % (kRingSlots ? kRingSlots : 1) == 0;
    static __device__ __forceinline__ uint32_t slot_of(int step, int j) {
        if (kStaticSlots) return (uint32_t)j & (uint32_t)(kRingSlots - 1);
        return ((uint32_t)step * kKeyChunks + j) % (uint32_t)(kRingSlots ? kRingSlots : 1);
    }
    static __device__ __forceinline__ uint32_t parity_of(int step, int j) {
        if (kStaticSlots) return ((uint32_t)j / (uint32_t)(kRingSlots ? kRingSlots : 1)) & 1u;   // (step * 8 + j) / slots, 8 / slots even
        return (((uint32_t)step * kKeyChunks + j) / (uint32_t)(kRingSlots ? kRingSlots : 1)) & 1u;
    }
    __device__ __forceinline__ void key_issue(uint32_t cc) const {     // one thread
        const uint32_t slot = cc % (uint32_t)(kRingSlots ? kRingSlots : 1);
        const uint32_t bar = full_s_ + 8u * slot, dst = ring_s_ + slot * (uint32_t)kChunkBytes;
        const char* src = bsk_bytes_ + (size_t)(cc / kKeyChunks) * (kBskStepElems * sizeof(cplx)) + (size_t)(cc % kKeyChunks) * kPieceBytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)kChunkBytes) : "memory");
#pragma unroll
        for (int row = 0; row < 2; row++)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst + row * (uint32_t)kPieceBytes), "l"(src + (size_t)row * (kBskStepElems / 2) * sizeof(cplx)),
                           "r"((uint32_t)kPieceBytes), "r"(bar) : "memory");
    }
    __device__ __forceinline__ uint32_t key_wait(int step, int j) const {
        mbar_wait(full_s_ + 8u * slot_of(step, j), parity_of(step, j));
        return ring_s_ + slot_of(step, j) * (uint32_t)kChunkBytes;
    }
    __device__ __forceinline__ cplx key_ld(uint32_t kc, int row, int r, int col, int k1) const {
        double x, y;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y)
                     : "r"(kc + (uint32_t)(((row * kKeyChunkRows + r) * 2 + col) * 32 + k1) * 16u));
        return cplx{x, y};
    }
    __device__ __forceinline__ void key_done(int step, int j) const {
        __syncwarp();
        if (lane_ == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(cons_s_ + 8u * slot_of(step, j)) : "memory");
    }
    // the warp whose turn it is refills the slot of chunk (step, j) with the chunk `slots` further on, once every
    // warp of the CTA has left it
    __device__ __forceinline__ void key_duty(int step, int j) const {
        const uint32_t cc = (uint32_t)step * kKeyChunks + j, slots = kRingSlots ? kRingSlots : 1;
        if (lane_ == 0 && ((uint32_t)j & 1u) == (uint32_t)warp_ && cc + slots < (uint32_t)n_ * kKeyChunks) {   // the two warps of the CTA's first PBS take turns
            mbar_wait(cons_s_ + 8u * slot_of(step, j), parity_of(step, j));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            key_issue(cc + slots);
        }
    }
    acc_t* acc_;
    double* xbuf_;
    double* xbuf_partner_;
    uint16_t* atilde_;
    uint32_t acc_s_;   // shared-space address of this polynomial's accumulator, 8 KiB aligned
    int bar_in_ = 0, bar_out_ = 0;   // named barriers of the FP64 baton (0 = no partner: no baton)
    bool leader_ = false;
    // word ((x >> 2) mod N) of the accumulator, negated when bit 13 of the byte offset x is set (negacyclic wrap)
    __device__ __forceinline__ acc_t acc_ld_rot(uint32_t x) const {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(kAccAligned ? (acc_s_ | (x & 0x1ffcu)) : (acc_s_ + (x & 0x1ffcu))) : "memory");
        if (FHESTR_BR_ABLATE == 7) return v;
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
    // FP64 baton: wait for the partner PBS to leave its FP64 stretch / tell it that this one has left its own
    __device__ __forceinline__ void fp_acquire() {
        if (FHESTR_BR_BATON != 0 && bar_in_) asm volatile("bar.sync %0, 128;" ::"r"(bar_in_));
    }
    __device__ __forceinline__ void fp_release() {
        if (FHESTR_BR_BATON != 0 && bar_out_) asm volatile("bar.arrive %0, 128;" ::"r"(bar_out_));
    }
    __device__ __forceinline__ void fp_start() {   // the follower hands the leader its first turn
        if (FHESTR_BR_BATON != 0 && bar_out_ && !leader_) asm volatile("bar.arrive %0, 128;" ::"r"(bar_out_));
    }
    __device__ __forceinline__ void fp_finish() {  // the leader takes the follower's last hand-over
        if (FHESTR_BR_BATON != 0 && bar_in_ && leader_) asm volatile("bar.sync %0, 128;" ::"r"(bar_in_));
    }
    // twiddles in tensor memory (FHESTR_BR_TMEM_TW): words [32 ch, 32 ch + 32) of this lane's 128 = twiddles 8 ch .. 8 ch + 7
    uint32_t tmem_tw_ = 0;   // TMEM address of this warp's lanes, column 0 of the CTA's allocation
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
    if (!FHESTR_BR_RING && b >= A.B) return;
    DevCtx c;
    c.lane_ = threadIdx.x & 31;
    c.poly_ = warp & 1;
    c.slot_ = slot;
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
    constexpr uint32_t front = P * (kXbufBytes + kMaskBytes) + (FHESTR_BR_RING ? kMbarBytes : 0);
    const uint32_t pad = kAccAligned ? (8192u - ((s0 + front) & 8191u)) & 8191u : 0u;
    uint32_t dyn;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if (front + pad + P * kAccBytes + kRingSlots * kChunkBytes > dyn || A.n + 1 > kMaskBytes / 2) __trap();   // loud, never a wrong result
    double* xb = reinterpret_cast<double*>(smem + slot * kXbufBytes);
    c.atilde_ = reinterpret_cast<uint16_t*>(smem + P * kXbufBytes + slot * kMaskBytes);
    const uint32_t acc_off = front + pad + slot * kAccBytes + c.poly_ * kN * (uint32_t)sizeof(acc_t);
    c.acc_ = reinterpret_cast<acc_t*>(smem + acc_off);
    c.acc_s_ = s0 + acc_off;
    c.xbuf_ = xb + c.poly_ * kWarpXbufDoubles;
    c.xbuf_partner_ = xb + (1 - c.poly_) * kWarpXbufDoubles;
#if FHESTR_BR_RING
    {
        const int pbs_here = min(P, A.B - (int)blockIdx.x * P);
        if (s0 != (uint32_t)kSlimSharedBase) __trap();
        c.warp_ = warp;
        c.n_ = A.n;
        c.bsk_bytes_ = reinterpret_cast<const char*>(A.bsk);
        if (threadIdx.x == 0) {
            for (int q = 0; q < kRingSlots; q++) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(c.full_s_ + 8u * q) : "memory");
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(c.cons_s_ + 8u * q), "r"((uint32_t)(2 * pbs_here)) : "memory");
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (int q = 0; q < kRingSlots && q < A.n * kKeyChunks; q++) c.key_issue((uint32_t)q);
        }
        __syncthreads();   // every warp of the CTA is still here: absent PBS leave below
    }
#endif
    if (b >= A.B) return;  // whole pair leaves together; pair barriers are per pair, batons only between present PBS
    if (P > 1) {
        const int partner = slot ^ kPartnerXor;
        c.leader_ = slot < partner;
        if (FHESTR_BR_BATON != 0 && blockIdx.x * P + partner < A.B) {
            // pair index among the P/2 baton pairs: the slot with the partner bit cleared, compacted
            const int lo = c.leader_ ? slot : partner;
            const int pair = kPartnerXor == 1 ? (lo >> 1) : (kPartnerXor == 2 ? (lo & 1) : lo);
            const int to_follower = P + 1 + 2 * pair, to_leader = P + 2 + 2 * pair;
            c.bar_in_ = c.leader_ ? to_leader : to_follower;
            c.bar_out_ = c.leader_ ? to_follower : to_leader;
        }
        if (FHESTR_BR_STAGGER_NS > 0 && !c.leader_) __nanosleep(FHESTR_BR_STAGGER_NS);
        if (FHESTR_BR_STAGGER2_NS > 0 && ((c.leader_ ? slot : partner) & (kPartnerXor == 1 ? 2 : 1))) __nanosleep(FHESTR_BR_STAGGER2_NS);
    }

#if FHESTR_BR_TMEM_TW
    // 128 TMEM columns per CTA (4 CTAs per SM = all 512): warp w owns lanes 32 w .. 32 w + 31, one twiddle word per column
    {
        uint32_t* slot_addr = reinterpret_cast<uint32_t*>(c.atilde_);   // the mask buffer is free until br_thread_main fills it
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"((uint32_t)__cvta_generic_to_shared(slot_addr)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tbase = *reinterpret_cast<volatile uint32_t*>(slot_addr);
        c.tmem_tw_ = tbase + ((uint32_t)(warp & 3) << 21);   // lane field: bits 16.., 32 lanes per warp
        __syncthreads();                                      // everyone has read the address before the mask is written
#pragma unroll
        for (int ch = 0; ch < 4; ch++) {
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
#endif
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
#ifdef FHESTR_BR_PHASELOG
    if (threadIdx.x % 64 == 0) {
        unsigned smid, warpid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
        if (smid < kLogSms) {
            const unsigned k = atomicAdd(&g_log_claim[smid], 1u);
            if (k < 4) {
                c.log_slot_ = (int)k; c.log_sm_ = (int)smid;
                g_log_meta[smid][k][0] = warpid; g_log_meta[smid][k][1] = blockIdx.x; g_log_meta[smid][k][2] = slot;
            }
        }
    }
#endif
    c.fp_start();
#ifdef FHESTR_BR_PHASELOG
    if (threadIdx.x % 64 == 0 && b < 8192) {
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_cta_log[b][0] = smid; g_cta_log[b][1] = gtimer();
        c.log_b_ = b;
    }
#endif
    br_thread_main(c, job, A.bsk, A.tf, A.ti);
#ifdef FHESTR_BR_PHASELOG
    if (threadIdx.x % 64 == 0 && b < 8192) g_cta_log[b][3] = gtimer();
#endif
#if FHESTR_BR_TMEM_TW
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // both warps are done with their columns
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(c.tmem_tw_ & 0xffffu) : "memory");
#endif
}

cudaError_t blind_rotate_configure() {
    if (FHESTR_BR_RING || kCtasPerSm > 4) {   // two 2-PBS CTAs need (nearly) all of the SM's shared memory; the key no longer goes through L1
        constexpr int ctas = kCtasPerSm;
        constexpr int pct = (100 * ctas * (kSlimSmemBytes + 1024) + 228 * 1024 - 1) / (228 * 1024);
        cudaError_t e = cudaFuncSetAttribute(blind_rotate_kernel<kPbsPerCta, kCtasPerSm>, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
        if (e != cudaSuccess) return e;
    }
    return cudaFuncSetAttribute(blind_rotate_kernel<kPbsPerCta, kCtasPerSm>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlimSmemBytes);
}

// Measured (r1, 4096 PBS): one PBS per 64-thread CTA, 4 CTAs per SM at 255 registers is the fastest shape; independent
// CTAs drift out of phase and overlap their FP64 and shared-memory phases, while 2 or 4 PBS per CTA ran in lockstep
// (1.3-1.5x slower), and sizing the register allocation for 5-6 CTAs per SM (168 registers) made every warp ~1.6x
// slower for 1.5x the warps (net 0.8x).  Those variants are gone from the tree (git history: round 1).
int launch_blind_rotate(const BrBatchArgs& a, cudaStream_t s) {
    if (a.B <= 0) return 0;
    blind_rotate_kernel<kPbsPerCta, kCtasPerSm><<<(a.B + kPbsPerCta - 1) / kPbsPerCta, 64 * kPbsPerCta, kSlimSmemBytes, s>>>(a);
    return 1;
}

}  // namespace fhestr

#ifdef FHESTR_BR_PHASELOG
extern "C" __attribute__((visibility("default"))) int fhestr_debug_phase_log(unsigned long long* log, unsigned* meta, int reset) {
    using namespace fhestr;
    if (log && reset != 2) cudaMemcpyFromSymbol(log, g_log, sizeof(g_log));
    if (meta) cudaMemcpyFromSymbol(meta, g_log_meta, sizeof(g_log_meta));
    if (reset == 2 && log) cudaMemcpyFromSymbol(log, g_cta_log, sizeof(g_cta_log));
    if (reset) { unsigned z[kLogSms] = {}; cudaMemcpyToSymbol(g_log_claim, z, sizeof(z)); }
    return kLogSms * 4 * kLogMarks * kLogSteps;
}
#endif
