// br_core.cuh -- per-thread body of the blind-rotation CMUX step (GGSW x GLWE external product with a
// register/shared-memory negacyclic f64 FFT), written once and compiled twice:
//   * by nvcc into the sm_100a kernel (blind_rotate.cu), Ctx = device warp context;
//   * by g++ into the host emulation used by the CPU tests (tests/emu/br_emu.cpp), Ctx = std::thread
//     + std::barrier context, so the exact index logic is checked against the oracle without a GPU.
//
// Replaces (reference side): the tfhe-rs blind rotation behind every PBS issued from
// /root/reference/src/ciphertext/fheasciichar.rs:36-102.
//
// Geometry.  N = 2048, k = 1, one decomposition level of 23 bits.  A polynomial p is folded to M = 1024
// complex points c_n = p_n + i p_{n+M}; its negacyclic spectrum is X_k = sum_n c_n zeta^n W^{nk},
// zeta = exp(i pi / N), W = exp(-2 pi i / M)  (so X_k = p(y_k), y_k = exp(i pi (1-4k)/N), y_k^N = -1).
// Four-step 32 x 32:  n = 32 n1 + n2,  k = k1 + 32 k2.
//   pass 1 (lane = n2, registers n1 -> k1): merged-twist 32-point transform (fft32_fwd_p1)
//   twiddle Tf(k1,n2) = exp(i pi n2 (1-4 k1)/N), transpose through shared memory
//   pass 2 (lane = k1, registers n2 -> k2): plain DFT-32 (fft32_fwd_p2)
// ONE WARP owns one polynomial: 32 complex points per lane live in registers, the only exchange inside
// a transform is one 32x32 transpose (real and imaginary parts through two 8.25 KB padded buffers), and
// only __syncwarp is needed.  The two warps of a PBS (mask polynomial, body polynomial) swap one
// spectrum per step through the same buffers (pair barrier).  The inverse mirrors the forward and ends
// in the layout the next step's forward starts from, so the new accumulator words stay in registers.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define FHE_HD __host__ __device__ __forceinline__
#else
#define FHE_HD inline
#endif
#include "fft32_gen.cuh"

// Compile-time switches of the step (A/B-measured on one B200, profiles/r1_final2_ab_variants.md; build.py --variant
// builds any other setting as a second library for FHESTR_ENGINE_LIB):
//   FHESTR_BR_SLIM=1 (default)  the accumulator buffers sit on 8 KiB-aligned shared addresses so the rotation gather
//       forms each address with ONE logic op, and the signed digit keeps -2^22 instead of mapping it to +2^22 (the
//       same torus value, |digit| unchanged).  The integer prologue is instruction-fetch bound
//       (profiles/r1_br_stall_breakdown.md): 8 instead of 12 instructions per coefficient, loop 3 812 -> 3 554.
//       0 = the first round-1 kernel (also the only form with 2 or 4 PBS per CTA).
//   FHESTR_BR_CVT_FP64=4 (default)  every 4th torus conversion of the epilogue runs on the FP64 pipe, see below.
//   FHESTR_BR_PREFETCH=8, FHESTR_BR_I2F_FP64=0: measured, no gain beyond noise.
//   FHESTR_BR_L1PF=0: NOT MEASURED YET (round-2 candidate).  1 = the key rows of a half-step that are not prefetched
//       into registers are requested into L1 ahead of the pair barrier (prefetch.global.L1, no registers): the two
//       FMAs that wait longest in the product stage wait on exactly those words coming from L2
//       (profiles/r1_br_stall_breakdown.md).
#ifndef FHESTR_BR_SLIM
#define FHESTR_BR_SLIM 1
#endif
// FHESTR_BR_ABLATE (timing experiments only, results are wrong): 1 = no 32-point codelets, 2 = no transposes and no
// spectrum exchange (barriers kept), 3 = no key loads, 4 = no inter-pass twiddle loads, 5 = no accumulator gather /
// store -- where the throughput kernel's step goes when its parts are removed one at a time (profiles/r2_br_ablation.md);
// with FHESTR_BR_COMPACT: 6 = no twists along n1 (loop 31.8 KB), 7 = also no sign handling in the gather (29.8 KB)
#ifndef FHESTR_BR_ABLATE
#define FHESTR_BR_ABLATE 0
#endif

// FHESTR_BR_BATON (bit mask, 0 = off): FP64 baton between the two PBS of a CTA whose warps sit on the same SM
// sub-partitions.  Bit 0 = forward pass 1 (+ twiddle), bit 1 = forward pass 2, bit 2 = GGSW product + inverse pass 1
// (+ twiddle), bit 3 = inverse pass 2.  A stretch whose bit is set is entered through Ctx::fp_acquire() and left through
// Ctx::fp_release(): the two PBS take strict turns, so that one is in a shared-memory / integer stretch while the other
// owns the FP64 pipe (independent CTAs only reach that by chance; in lockstep the step is 1.3-1.5 x slower).
#ifndef FHESTR_BR_BATON
#define FHESTR_BR_BATON 0
#endif

// FHESTR_BR_RING=1: the Fourier key reaches the product stage through a shared-memory ring filled by bulk TMA
// (Ctx::key_wait / key_ld / key_done / key_duty) instead of 16-byte loads from L2: a step's 64 KiB tile is 8 chunks
// of 4 spectrum rows (8 KiB: the rows of both GGSW rows), consumed in order by every warp of the CTA.
#ifndef FHESTR_BR_RING
#define FHESTR_BR_RING 0
#endif
// FHESTR_BR_ONE_TWIDDLE=1: the inverse transform multiplies by conj(tf) AFTER its transpose (lane = n2, register k1:
// the same coalesced rows of the same table the forward reads) instead of by the transposed copy ti before it --
// same values, same products, half the twiddle footprint in L1.  Off: ptxas then keeps the 32 twiddle loads in flight
// across the start of the last codelet and spills (200-550 bytes per thread).
#ifndef FHESTR_BR_ONE_TWIDDLE
#define FHESTR_BR_ONE_TWIDDLE 0
#endif

// FHESTR_BR_ACC_SMEM=1: no register copy of the warp's own 64 accumulator words between steps; the prologue reads them
// back from shared memory (64 more LDS.32 per warp-step).  64 registers fewer: what the ring form needs to stay free
// of spills (its mbarrier waits split the step into many blocks).
#ifndef FHESTR_BR_ACC_SMEM
#define FHESTR_BR_ACC_SMEM FHESTR_BR_RING
#endif

// FHESTR_BR_TMEM_TW=1: a lane's 32 inter-pass twiddles tf[k1*32 + lane] live in TENSOR MEMORY (128 columns of the
// CTA's TMEM allocation, lane-private: tcgen05.st once per PBS, tcgen05.ld 8 twiddles at a time) instead of being
// re-read from L1 / L2 in every transform: the 64 LDG.128 per warp-step are 17 % of the kernel's LSU wavefronts, its
// busiest unit, and TMEM has its own data path.  The inverse multiplies after its transpose (ONE_TWIDDLE), so one set
// of values serves both directions.  Ctx::tw_ld / tw_wait.
#ifndef FHESTR_BR_TMEM_TW
#define FHESTR_BR_TMEM_TW 0
#endif
#if FHESTR_BR_TMEM_TW
#undef FHESTR_BR_ONE_TWIDDLE
#define FHESTR_BR_ONE_TWIDDLE 1
#endif

// FHESTR_BR_COMPACT=1: the CMUX step as a ROLLED loop of four passes over ONE copy of the plain DFT-32 codelet
// (cmux_step_compact below) instead of four different straight-line codelets: the step's code shrinks from 57.7 KB to
// about the size of the SM's 32 KB instruction cache (profiles/r2_phase_log.md: the straight-line loop streams from
// the GPC-level instruction cache, which is 91 % busy).  The inverse passes are the forward codelet on swapped
// real / imaginary parts, the twist along n1 is applied outside the codelet (kFft32Twist*).  Same transform, same
// spectrum layout, same key.
#ifndef FHESTR_BR_COMPACT
#define FHESTR_BR_COMPACT 0
#endif

namespace fhestr {

constexpr int kKeyChunkRows = 4;                       // spectrum rows k2 per ring chunk
constexpr int kKeyChunks = 32 / kKeyChunkRows;         // chunks per CMUX step
constexpr int kKeyChunkElems = 2 * kKeyChunkRows * 2 * 32;   // complex words: [GGSW row][k2 in chunk][col][k1] = 8 KiB
constexpr int kKeyPieceElems = kKeyChunkRows * 2 * 32;       // one GGSW row of a chunk: contiguous in the global layout

constexpr int kBatonPhases = ((FHESTR_BR_BATON >> 0) & 1) + ((FHESTR_BR_BATON >> 1) & 1) + ((FHESTR_BR_BATON >> 2) & 1) + ((FHESTR_BR_BATON >> 3) & 1);
template <class Ctx> FHE_HD void baton_in(Ctx& c, int bit) { if ((FHESTR_BR_BATON >> bit) & 1) c.fp_acquire(); }
template <class Ctx> FHE_HD void baton_out(Ctx& c, int bit) { if ((FHESTR_BR_BATON >> bit) & 1) c.fp_release(); }

typedef unsigned long long u64;
typedef long long i64;

constexpr int kN = 2048;          // polynomial size
constexpr int kM = 1024;          // complex points
constexpr int kXPad = 33;         // transpose row stride (doubles): conflict-free 64-bit column reads
constexpr int kXbufDoubles = 32 * kXPad;  // 1056 doubles = 8448 B: one padded 32 x 32 matrix
constexpr int kWarpXbufDoubles = kXbufDoubles;  // per warp: ONE matrix (shared-memory carve-out 164 KB instead of 228 KB: 92 KB of L1 for twiddles and the BSK tile)
constexpr int kPbsBaseLog = 23;
#ifndef FHESTR_BR_PREFETCH
#define FHESTR_BR_PREFETCH 8
#endif
constexpr int kBskPrefetch = FHESTR_BR_PREFETCH;    // key rows per half-step requested ahead of the pair barrier (8 = 64 registers)
// FHESTR_BR_CVT_FP64 = m > 0: every m-th torus conversion of the epilogue runs on the FP64 pipe (four FP64 instructions,
// bit-identical to the F2I) instead of the conversion unit, which sustains one F2I.S64 per 8 cycles per sub-partition
#ifndef FHESTR_BR_L1PF
#define FHESTR_BR_L1PF 0
#endif
#ifndef FHESTR_BR_CVT_FP64
#define FHESTR_BR_CVT_FP64 4
#endif

struct alignas(16) cplx { double x, y; };

FHE_HD cplx cmul(cplx a, cplx b) {
    cplx r;
    r.x = fma(a.x, b.x, -(a.y * b.y));
    r.y = fma(a.x, b.y, a.y * b.x);
    return r;
}

// A.6 modulus switch to 2N = 4096, reduced to [0, 2N)
FHE_HD uint32_t modswitch_2N(u64 x) {
    u64 t = x >> (64 - 12 - 1);
    t += t & 1;
    t >>= 1;
    return (uint32_t)t & (2 * kN - 1);
}

// The accumulator lives on the 32-bit torus: acc_t = the top 32 bits of tfhe-rs' u64 torus words.  The
// decomposition only ever looks at the top 23 bits (+ one rounding bit) of a difference, and what each CMUX
// adds is known to about 2^-25 of the torus (f64 FFT round-off), so the 2^-33 rounding of a 32-bit word is
// two orders of magnitude below the noise the step already has -- and it halves the accumulator's shared
// memory (16 KiB per PBS instead of 32), its traffic and the integer work of every step.
typedef uint32_t acc_t;

FHE_HD acc_t acc_from_u64(u64 x) { return (acc_t)((x + (1ull << 31)) >> 32); }
FHE_HD u64 acc_to_u64(acc_t a) { return (u64)a << 32; }

// A.4 signed decomposition, one level of 23 bits, of a 32-bit torus difference: digit in (-2^22, 2^22]
FHE_HD double digit23(acc_t x) {
    int32_t d = ((int32_t)(x + (1u << 8))) >> 9;
    if (d == -(1 << 22)) d = (1 << 22);
    return (double)d;
}

// slim variant: the tie x = 2^31 - 256 .. 2^31 - 1 stays -2^22 (same torus value as +2^22, two instructions fewer)
// FHESTR_BR_I2F_FP64=1: the int -> double conversion of the digit as one DADD on the FP64 pipe instead of an I2F.F64 on
// the conversion unit: the biased digit (x + 2^8 + 2^31) >> 9 in [0, 2^23) is dropped into the low mantissa word of
// 2^52 and 2^52 + 2^22 is subtracted.  Exact, same value.
#ifndef FHESTR_BR_I2F_FP64
#define FHESTR_BR_I2F_FP64 0
#endif
FHE_HD double digit23_slim(acc_t x) {
#if FHESTR_BR_I2F_FP64
    const uint32_t biased = (x + ((1u << 8) + (1u << 31))) >> 9;
#ifdef __CUDA_ARCH__
    return __hiloint2double(0x43300000, (int)biased) - 4503599631564800.0;   // 2^52 + 2^22
#else
    const u64 bits = 0x4330000000000000ull | biased;
    double d;
    __builtin_memcpy(&d, &bits, 8);
    return d - 4503599631564800.0;
#endif
#else
    return (double)(((int32_t)(x + (1u << 8))) >> 9);
#endif
}

// coefficient j of X^e * P (negacyclic), e in [0, 2N); generic word type
template <class T>
FHE_HD T rot_coef(const T* P, int j, int e) {
    const int q = (j - e) & (2 * kN - 1);
    const T v = P[q & (kN - 1)];
    return (q & kN) ? (T)0 - v : v;
}

// x in units of 2^-32 turns -> round(x) mod 2^32 (the Fourier BSK carries the 2^-32 / M scale).  |x| is about
// 2^57 (digits 2^22 x key 2^31 x sqrt(4096) terms) and the 64-bit conversion saturates only beyond 2^63, a
// 40-sigma event, so ONE round-to-nearest conversion and a truncation replace the usual x - rint(x) sequence.
FHE_HD acc_t torus32_from_double(double x) {
#ifdef __CUDA_ARCH__
    return (acc_t)(unsigned long long)__double2ll_rn(x);
#else
    return (acc_t)(u64)(i64)llrint(x);
#endif
}

// The same value on the FP64 pipe: r = rint(x 2^-32) by the 1.5 * 2^52 trick, x - r 2^32 is exact and lies in
// [-2^31, 2^31], and adding 1.5 * 2^52 once more leaves round-to-nearest-even(x) mod 2^32 in the low mantissa word.
FHE_HD acc_t torus32_from_double_fp64(double x) {
    const double M = 6755399441055744.0;   // 1.5 * 2^52
    const double r = fma(x, 1.0 / 4294967296.0, M) - M;
    const double u = fma(r, -4294967296.0, x) + M;
#ifdef __CUDA_ARCH__
    return (acc_t)__double2loint(u);
#else
    u64 bits;
    __builtin_memcpy(&bits, &u, 8);
    return (acc_t)bits;
#endif
}
FHE_HD acc_t torus32_conv(double x, int idx) {
#if FHESTR_BR_CVT_FP64 > 0
    if (idx % FHESTR_BR_CVT_FP64 == 0) return torus32_from_double_fp64(x);
#endif
    (void)idx;
    return torus32_from_double(x);
}

// 32x32 transpose of one double per (lane, register) through the warp's padded buffer, real parts then imaginary
template <class Ctx>
FHE_HD void transpose32_half(Ctx& c, double (&v)[32]) {
    const int t = c.lane();
    double* buf = c.xbuf();
#if FHESTR_BR_ABLATE != 2
#pragma unroll
    for (int r = 0; r < 32; r++) buf[r * kXPad + t] = v[r];
#endif
    c.syncwarp();
#if FHESTR_BR_ABLATE != 2
#pragma unroll
    for (int r = 0; r < 32; r++) v[r] = buf[t * kXPad + r];
#endif
    c.syncwarp();
}
template <class Ctx>
FHE_HD void transpose32(Ctx& c, double (&re)[32], double (&im)[32]) {
    transpose32_half(c, re);
    transpose32_half(c, im);
}

// forward transform of the 32 complex points held by this lane (lane = n2, register n1) into the
// spectrum layout (lane = k1, register k2).  tf[k1*32 + n2].
template <class Ctx>
FHE_HD void forward1024(Ctx& c, double (&re)[32], double (&im)[32], const cplx* tf) {
    const int t = c.lane();
    c.log_mark(1);
    baton_in(c, 0);
#if FHESTR_BR_ABLATE != 1
    fft32_fwd_p1(re, im);
#endif
#if FHESTR_BR_TMEM_TW
    {
        // 8 twiddles (32 words) per tcgen05.ld; the next chunk is in flight while this one is used
        uint32_t wa[32], wb[32];
        c.tw_ld(0, wa, tf);
        c.tw_wait(wa);
#pragma unroll
        for (int ch = 0; ch < 4; ch++) {
            uint32_t (&cur)[32] = (ch & 1) ? wb : wa;
            uint32_t (&nxt)[32] = (ch & 1) ? wa : wb;
            if (ch < 3) c.tw_ld(ch + 1, nxt, tf);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int k1 = ch * 8 + j;
                const cplx w = cplx{c.tw_word(cur[4 * j], cur[4 * j + 1]), c.tw_word(cur[4 * j + 2], cur[4 * j + 3])};
                const cplx z = cmul(cplx{re[k1], im[k1]}, w);
                re[k1] = z.x; im[k1] = z.y;
            }
            if (ch < 3) c.tw_wait(nxt);
        }
    }
#else
#pragma unroll
    for (int k1 = 0; k1 < 32; k1++) {
#if FHESTR_BR_ABLATE == 4
        const cplx w = cplx{0.8 + 0.001 * k1, 0.6};
#else
        const cplx w = c.ldg(tf + k1 * 32 + t);
#endif
        const cplx z = cmul(cplx{re[k1], im[k1]}, w);
        re[k1] = z.x; im[k1] = z.y;
    }
#endif
    baton_out(c, 0);
    c.log_mark(2);
    transpose32(c, re, im);
    c.log_mark(3);
    baton_in(c, 1);
#if FHESTR_BR_ABLATE != 1
    fft32_fwd_p2(re, im);
#endif
    baton_out(c, 1);
}

// inverse: (lane = k1, register k2) -> (lane = n2, register n1).  ti[n2*32 + k1] = conj(Tf(k1,n2)) = conj(tf[k1*32 + n2]).
template <class Ctx>
FHE_HD void inverse1024(Ctx& c, double (&re)[32], double (&im)[32], const cplx* tf, const cplx* ti) {
    const int t = c.lane();
#if FHESTR_BR_ABLATE != 1
    fft32_inv_p1(re, im);
#endif
#if !FHESTR_BR_ONE_TWIDDLE
#pragma unroll
    for (int n2 = 0; n2 < 32; n2++) {
        const cplx w = c.ldg(ti + n2 * 32 + t);
        const cplx z = cmul(cplx{re[n2], im[n2]}, w);
        re[n2] = z.x; im[n2] = z.y;
    }
#endif
    baton_out(c, 2);
    c.log_mark(6);
    transpose32(c, re, im);
    c.log_mark(7);
    baton_in(c, 3);
#if FHESTR_BR_ONE_TWIDDLE
    // after the transpose: lane = n2, register k1 -- conj(Tf(k1, n2)) = conj(tf[k1*32 + n2])
#if FHESTR_BR_TMEM_TW
    {
        uint32_t wa[32], wb[32];
        c.tw_ld(0, wa, tf);
        c.tw_wait(wa);
#pragma unroll
        for (int ch = 0; ch < 4; ch++) {
            uint32_t (&cur)[32] = (ch & 1) ? wb : wa;
            uint32_t (&nxt)[32] = (ch & 1) ? wa : wb;
            if (ch < 3) c.tw_ld(ch + 1, nxt, tf);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int k1 = ch * 8 + j;
                const cplx w = cplx{c.tw_word(cur[4 * j], cur[4 * j + 1]), c.tw_word(cur[4 * j + 2], cur[4 * j + 3])};
                const double zr = fma(re[k1], w.x, im[k1] * w.y);
                const double zi = fma(-re[k1], w.y, im[k1] * w.x);
                re[k1] = zr; im[k1] = zi;
            }
            if (ch < 3) c.tw_wait(nxt);
        }
    }
#else
#pragma unroll
    for (int k1 = 0; k1 < 32; k1++) {
#if FHESTR_BR_ABLATE == 4
        const cplx w = cplx{0.8 + 0.001 * k1, 0.6};
#else
        const cplx w = c.ldg(tf + k1 * 32 + t);
#endif
        const double zr = fma(re[k1], w.x, im[k1] * w.y);
        const double zi = fma(-re[k1], w.y, im[k1] * w.x);   // = cmul(x, conj w), operation for operation
        re[k1] = zr; im[k1] = zi;
    }
#endif
#endif
#if FHESTR_BR_ABLATE != 1
    fft32_inv_p2(re, im);
#endif
    baton_out(c, 3);
}

// Fourier BSK layout (engine-private, produced once by the key-conversion kernel):
//   g[ ((((step*2 + row)*32 + k2)*2 + col)*32 + k1 ]   complex f64, k = k1 + 32 k2
// scaled by 2^-32 / M so that the inverse transform directly yields units of 2^-32 turns (acc_t ulps).
constexpr int kBskStepElems = 2 * 32 * 2 * 32;  // 4096 complex = 64 KiB per CMUX step
FHE_HD int bsk_index(int row, int k2, int col, int k1) { return ((row * 32 + k2) * 2 + col) * 32 + k1; }

// One CMUX step for this warp's polynomial:  ACC += GGSW (x) (X^e * ACC - ACC).
//   a[0..31]  = ACC[32 n1 + lane], a[32..63] = ACC[32 n1 + lane + M]  (registers, in/out)
//   c.acc()   = this polynomial's accumulator in shared memory (same values), updated on exit
template <class Ctx>
FHE_HD void cmux_step(Ctx& c, acc_t (&a)[64], int e, int step, const cplx* g, const cplx* tf, const cplx* ti) {
    const int t = c.lane();
    const int p = c.poly();
    acc_t* acc = c.acc();
    double re[32], im[32];
    c.log_mark(0);
    // rotate, subtract, decompose
#if FHESTR_BR_SLIM
    // byte offset of coefficient (t - e) mod 2N in the 2N-word negacyclic extension; row n1 adds 128 bytes.  Bit 13
    // of the running offset is the sign, bits 2..12 the word inside the 8 KiB buffer (c.acc_ld_rot)
    const uint32_t x0 = (uint32_t)((t - e) & (2 * kN - 1)) << 2;
#pragma unroll
    for (int n1 = 0; n1 < 32; n1++) {
#if FHESTR_BR_ABLATE == 5
        re[n1] = digit23_slim((x0 * 2654435761u + n1) - a[n1]);
        im[n1] = digit23_slim((x0 * 40503u + n1) - a[32 + n1]);
#elif FHESTR_BR_ACC_SMEM
        re[n1] = digit23_slim(c.acc_ld_rot(x0 + 128u * n1) - acc[32 * n1 + t]);
        im[n1] = digit23_slim(c.acc_ld_rot(x0 + 128u * n1 + 4096u) - acc[32 * n1 + t + kM]);
#else
        re[n1] = digit23_slim(c.acc_ld_rot(x0 + 128u * n1) - a[n1]);
        im[n1] = digit23_slim(c.acc_ld_rot(x0 + 128u * n1 + 4096u) - a[32 + n1]);
#endif
    }
#else
#pragma unroll
    for (int n1 = 0; n1 < 32; n1++) {
        const int j = 32 * n1 + t;
        re[n1] = digit23(rot_coef(acc, j, e) - a[n1]);
        im[n1] = digit23(rot_coef(acc, j + kM, e) - a[32 + n1]);
    }
#endif
    forward1024(c, re, im, tf);
    c.log_mark(4);
    // Fourier-domain GGSW product.  The two warps swap their spectra through the transpose buffers (1024 complex
    // points fit in the two matrices) and each forms ITS output polynomial completely:
    //     out_p = D_p * G[p][p] + D_(1-p) * G[1-p][p]
    // the second product accumulates with FMAs, so the step costs 8 FP64 operations per point instead of 10
    cplx* xo = reinterpret_cast<cplx*>(c.xbuf());
    const cplx* xp = reinterpret_cast<const cplx*>(c.xbuf_partner());
#if FHESTR_BR_RING
    (void)g;
#pragma unroll
    for (int half = 0; half < 2; half++) {   // 16 spectrum rows at a time: the buffer holds 528 complex points
#pragma unroll
        for (int q = 0; q < 16; q++) xo[q * 32 + t] = cplx{re[half * 16 + q], im[half * 16 + q]};
        c.log_mark(10 + half * 4);
        c.pair_sync();
        c.log_mark(11 + half * 4);
        if (half == 1) baton_in(c, 2);
#pragma unroll
        for (int jc = 0; jc < 16 / kKeyChunkRows; jc++) {
            const int j = half * (16 / kKeyChunkRows) + jc;      // chunk of this step
            const auto kc = c.key_wait(step, j);
#pragma unroll
            for (int r = 0; r < kKeyChunkRows; r++) {
                const int q = jc * kKeyChunkRows + r, k2 = half * 16 + q;
                const cplx gs = c.key_ld(kc, p, r, p, t);
                const cplx go = c.key_ld(kc, 1 - p, r, p, t);
                const cplx v = xp[q * 32 + t];
                const cplx s = cmul(cplx{re[k2], im[k2]}, gs);
                re[k2] = fma(v.x, go.x, fma(-v.y, go.y, s.x));
                im[k2] = fma(v.x, go.y, fma(v.y, go.x, s.y));
            }
            c.key_done(step, j);
            if (jc > 0) c.key_duty(step, j - 1);     // refill the slot of the chunk before: everyone has left it by now
        }
        c.log_mark(12 + half * 4);
        c.pair_sync();
        c.key_duty(step, half * (16 / kKeyChunkRows) + 16 / kKeyChunkRows - 1);
        c.log_mark(13 + half * 4);
    }
#else
#pragma unroll
    for (int half = 0; half < 2; half++) {   // 16 spectrum rows at a time: the buffer holds 528 complex points
#pragma unroll
        for (int q = 0; q < 16; q++)
            if (FHESTR_BR_ABLATE != 2) xo[q * 32 + t] = cplx{re[half * 16 + q], im[half * 16 + q]};
        // the first key words of this half do not depend on the partner: request them BEFORE the barrier so that
        // their L2 latency overlaps the wait
        cplx gsv[kBskPrefetch], gov[kBskPrefetch];
#pragma unroll
        for (int q = 0; q < kBskPrefetch; q++) {
#if FHESTR_BR_ABLATE == 3
            gsv[q] = cplx{1e-9 * q, 2e-9};
            gov[q] = cplx{3e-9, 1e-9 * q};
#else
            gsv[q] = c.ldg(g + bsk_index(p, half * 16 + q, p, t));
            gov[q] = c.ldg(g + bsk_index(1 - p, half * 16 + q, p, t));
#endif
        }
#if FHESTR_BR_L1PF
#pragma unroll
        for (int q = kBskPrefetch; q < 16; q++) {
            c.prefetch_l1(g + bsk_index(p, half * 16 + q, p, t));
            c.prefetch_l1(g + bsk_index(1 - p, half * 16 + q, p, t));
        }
#endif
        c.log_mark(10 + half * 4);
        c.pair_sync();
        c.log_mark(11 + half * 4);
        if (half == 1) baton_in(c, 2);
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int k2 = half * 16 + q;
#if FHESTR_BR_ABLATE == 3
            const cplx gs = cplx{1e-9 * q, 2e-9}, go = cplx{3e-9, 1e-9 * q};
#else
            const cplx gs = q < kBskPrefetch ? gsv[q < kBskPrefetch ? q : 0] : c.ldg(g + bsk_index(p, k2, p, t));
            const cplx go = q < kBskPrefetch ? gov[q < kBskPrefetch ? q : 0] : c.ldg(g + bsk_index(1 - p, k2, p, t));
#endif
#if FHESTR_BR_ABLATE == 2
            const cplx v = cplx{re[k2] * 0.5, im[k2] * 0.25};
#else
            const cplx v = xp[q * 32 + t];
#endif
            const cplx s = cmul(cplx{re[k2], im[k2]}, gs);
            re[k2] = fma(v.x, go.x, fma(-v.y, go.y, s.x));
            im[k2] = fma(v.x, go.y, fma(v.y, go.x, s.y));
        }
        c.log_mark(12 + half * 4);
        c.pair_sync();
        c.log_mark(13 + half * 4);
    }
#endif
    c.log_mark(5);
    inverse1024(c, re, im, tf, ti);
    c.log_mark(8);
    // accumulate into the torus accumulator; keep the new words in registers for the next step
#pragma unroll
    for (int n1 = 0; n1 < 32; n1++) {
        const int j = 32 * n1 + t;
#if FHESTR_BR_ABLATE == 5
        a[n1] = a[n1] + torus32_conv(re[n1], n1);
        a[32 + n1] = a[32 + n1] + torus32_conv(im[n1], n1);
#elif FHESTR_BR_ACC_SMEM
        acc[j] = acc[j] + torus32_conv(re[n1], n1);
        acc[j + kM] = acc[j + kM] + torus32_conv(im[n1], n1);
#else
        a[n1] = acc[j] + torus32_conv(re[n1], n1);
        a[32 + n1] = acc[j + kM] + torus32_conv(im[n1], n1);
        acc[j] = a[n1];
        acc[j + kM] = a[32 + n1];
#endif
    }
    c.syncwarp();
    c.log_mark(9);
}

#ifdef __CUDA_ARCH__
#define FHESTR_TWIST(i) kFft32TwistDev[i]
#else
#define FHESTR_TWIST(i) kFft32TwistHost[i]
#endif

// FHESTR_BR_TWREG = K (compact loop only): the first K of a lane's 32 inter-pass twiddles tf[k1*32 + lane] stay in
// registers for the whole blind rotation (4 K registers); the inverse multiplies AFTER its transpose (lane = n2,
// register k1: the values the forward uses), so one set serves both directions and 2 K of the step's 64 twiddle loads
// go away.  FHESTR_BR_AREG=1: the accumulation adds to the register copy a[] instead of re-reading the accumulator
// from shared memory (64 LDS fewer per warp-step, 64 registers live across the passes).  Both trade the registers the
// compact loop frees (168 instead of 255) for shared-memory / L1 wavefronts: the LSU pipe is this kernel's busiest unit.
#ifndef FHESTR_BR_TWREG
#define FHESTR_BR_TWREG 0
#endif
#ifndef FHESTR_BR_AREG
#define FHESTR_BR_AREG 0
#endif
// FHESTR_BR_TW_EARLY=1 (with FHESTR_BR_TMEM_TW): the first twiddle chunk is requested from tensor memory BEFORE the
// pass whose results it multiplies (32 more live registers across the codelet), so its latency is never exposed
#ifndef FHESTR_BR_TW_EARLY
#define FHESTR_BR_TW_EARLY 0
#endif
// FHESTR_BR_KEY_EARLY=E (compact loop): the first E register-prefetched key rows of the product's first half are
// requested BEFORE forward pass 2 instead of after it, so that their L2 latency hides behind the codelet (8 E registers
// live across the loop: the rolled loop cannot scope them to one iteration)
#ifndef FHESTR_BR_KEY_EARLY
#define FHESTR_BR_KEY_EARLY 0
#endif
constexpr int kTwReg = FHESTR_BR_TWREG;

// multiply the 32 points by the inter-pass twiddles of this lane (registers for k1 < kTwReg, L1 / L2 beyond)
template <class Ctx>
FHE_HD void twiddle32(Ctx& c, double (&A)[32], double (&B)[32], const cplx* tf, const cplx (&twr)[kTwReg > 0 ? kTwReg : 1], uint32_t (&wa)[32]) {
    const int t = c.lane();
#if FHESTR_BR_TMEM_TW
    (void)twr; (void)t;
    uint32_t wb[32];
#if !FHESTR_BR_TW_EARLY
    c.tw_ld(0, wa, tf);
#endif
    c.tw_wait(wa);
#pragma unroll
    for (int ch = 0; ch < 4; ch++) {
        uint32_t (&cur)[32] = (ch & 1) ? wb : wa;
        uint32_t (&nxt)[32] = (ch & 1) ? wa : wb;
        if (ch < 3) c.tw_ld(ch + 1, nxt, tf);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int r = ch * 8 + j;
            const cplx w = cplx{c.tw_word(cur[4 * j], cur[4 * j + 1]), c.tw_word(cur[4 * j + 2], cur[4 * j + 3])};
            const cplx z = cmul(cplx{A[r], B[r]}, w);
            A[r] = z.x; B[r] = z.y;
        }
        if (ch < 3) c.tw_wait(nxt);
    }
#else
#pragma unroll
    for (int r = 0; r < 32; r++) {
        const cplx w = r < kTwReg ? twr[r < kTwReg ? r : 0] : c.ldg(tf + r * 32 + t);
        const cplx z = cmul(cplx{A[r], B[r]}, w);
        A[r] = z.x; B[r] = z.y;
    }
#endif
}

// The same CMUX step with its four 32-point passes rolled into one loop (FHESTR_BR_COMPACT).  (A, B) hold
// (real, imaginary) parts during the forward passes and (imaginary, real) during the inverse ones: the inverse DFT of x
// is swap(DFT(swap(x))), and a multiplication by conj(w) in the swapped domain is a multiplication by w, so the passes
// and the twiddle multiplication of both directions are the SAME instructions.
//   before the loop: gather + digits + twist exp(+i pi n1/64)
//   pass 0: DFT over n1, twiddle tf, transpose
//   pass 1: DFT over n2, spectrum exchange + GGSW product (results written swapped)
//   pass 2: DFT over k2 (inverse), transpose, twiddle tf (lane = n2, register k1 again)
//   pass 3: DFT over k1 (inverse)
//   after the loop: untwist exp(-i pi n1/64), accumulate
template <class Ctx>
FHE_HD void cmux_step_compact(Ctx& c, acc_t (&a)[64], int e, int step, const cplx* g, const cplx* tf, const cplx (&twr)[kTwReg > 0 ? kTwReg : 1]) {
    const int t = c.lane();
    const int p = c.poly();
    acc_t* acc = c.acc();
    double A[32], B[32];
    (void)step;
    // the gather consumes the register copy a[] BEFORE the rolled loop and the accumulation rewrites it AFTER it, so
    // that a[] is dead across the loop (inside it, the compiler would have to keep it alive over every pass)
    c.log_mark(0);
    {
        const uint32_t x0 = (uint32_t)((t - e) & (2 * kN - 1)) << 2;
#pragma unroll
        for (int n1 = 0; n1 < 32; n1++) {
            const double dr = digit23_slim(c.acc_ld_rot(x0 + 128u * n1) - a[n1]);
            const double di = digit23_slim(c.acc_ld_rot(x0 + 128u * n1 + 4096u) - a[32 + n1]);
            if (n1 == 0 || FHESTR_BR_ABLATE >= 6) { A[n1] = dr; B[n1] = di; }   // ABLATE 6, 7: timing without the twists
            else {
                const double cs = FHESTR_TWIST(2 * n1), sn = FHESTR_TWIST(2 * n1 + 1);
                A[n1] = fma(dr, cs, -(di * sn));
                B[n1] = fma(dr, sn, di * cs);
            }
        }
    }
    c.log_mark(1);
    // FHESTR_BR_COMPACT=1: four iterations over one codelet copy; 2: two iterations (forward, inverse) over two copies
    uint32_t w0[32];
#if FHESTR_BR_KEY_EARLY
    constexpr int kKeyEarly = FHESTR_BR_KEY_EARLY < kBskPrefetch ? FHESTR_BR_KEY_EARLY : kBskPrefetch;   // rows requested early
    cplx gs0[kKeyEarly], go0[kKeyEarly];
#endif
#if FHESTR_BR_COMPACT == 2
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
#if FHESTR_BR_TMEM_TW && FHESTR_BR_TW_EARLY
        c.tw_ld(0, w0, tf);
#endif
        fft32_fwd_p2(A, B);
        if (h == 0) {
            twiddle32(c, A, B, tf, twr, w0);
            c.log_mark(2);
            transpose32(c, A, B);
            c.log_mark(3);
        } else {
            c.log_mark(6);
            transpose32(c, A, B);
            c.log_mark(7);
            twiddle32(c, A, B, tf, twr, w0);
        }
        fft32_fwd_p2(A, B);
        if (h == 0) {
            c.log_mark(4);
            cplx* xo = reinterpret_cast<cplx*>(c.xbuf());
            const cplx* xp = reinterpret_cast<const cplx*>(c.xbuf_partner());
#pragma unroll
            for (int half = 0; half < 2; half++) {
#pragma unroll
                for (int q = 0; q < 16; q++) xo[q * 32 + t] = cplx{A[half * 16 + q], B[half * 16 + q]};
                cplx gsv[kBskPrefetch], gov[kBskPrefetch];
#pragma unroll
                for (int q = 0; q < kBskPrefetch; q++) {
                    gsv[q] = c.ldg(g + bsk_index(p, half * 16 + q, p, t));
                    gov[q] = c.ldg(g + bsk_index(1 - p, half * 16 + q, p, t));
                }
                c.log_mark(10 + half * 4);
                c.pair_sync();
                c.log_mark(11 + half * 4);
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const int k2 = half * 16 + q;
                    const cplx gs = q < kBskPrefetch ? gsv[q < kBskPrefetch ? q : 0] : c.ldg(g + bsk_index(p, k2, p, t));
                    const cplx go = q < kBskPrefetch ? gov[q < kBskPrefetch ? q : 0] : c.ldg(g + bsk_index(1 - p, k2, p, t));
                    const cplx v = xp[q * 32 + t];
                    const cplx s = cmul(cplx{A[k2], B[k2]}, gs);
                    const double zr = fma(v.x, go.x, fma(-v.y, go.y, s.x));
                    const double zi = fma(v.x, go.y, fma(v.y, go.x, s.y));
                    A[k2] = zi; B[k2] = zr;      // swapped from here on
                }
                c.log_mark(12 + half * 4);
                c.pair_sync();
                c.log_mark(13 + half * 4);
            }
            c.log_mark(5);
        }
    }
#else
#pragma unroll 1
    for (int it = 0; it < 4; it++) {
#if FHESTR_BR_TMEM_TW && FHESTR_BR_TW_EARLY
        if ((it & 1) == 0) c.tw_ld(0, w0, tf);
#endif
#if FHESTR_BR_KEY_EARLY
        if (it == 1) {
#pragma unroll
            for (int q = 0; q < kKeyEarly; q++) {
                gs0[q] = c.ldg(g + bsk_index(p, q, p, t));
                go0[q] = c.ldg(g + bsk_index(1 - p, q, p, t));
            }
        }
#endif
        fft32_fwd_p2(A, B);
        if (it == 0) {
            twiddle32(c, A, B, tf, twr, w0);
            c.log_mark(2);
            transpose32(c, A, B);
            c.log_mark(3);
        } else if (it == 2) {
            c.log_mark(6);
            transpose32(c, A, B);
            c.log_mark(7);
            twiddle32(c, A, B, tf, twr, w0);
        } else if (it == 1) {
            c.log_mark(4);
            cplx* xo = reinterpret_cast<cplx*>(c.xbuf());
            const cplx* xp = reinterpret_cast<const cplx*>(c.xbuf_partner());
#pragma unroll
            for (int half = 0; half < 2; half++) {
#pragma unroll
                for (int q = 0; q < 16; q++) xo[q * 32 + t] = cplx{A[half * 16 + q], B[half * 16 + q]};
                cplx gsv[kBskPrefetch], gov[kBskPrefetch];
#pragma unroll
                for (int q = 0; q < kBskPrefetch; q++) {
#if FHESTR_BR_KEY_EARLY
                    if (half == 0 && q < kKeyEarly) { gsv[q] = gs0[q < kKeyEarly ? q : 0]; gov[q] = go0[q < kKeyEarly ? q : 0]; continue; }
#endif
                    gsv[q] = c.ldg(g + bsk_index(p, half * 16 + q, p, t));
                    gov[q] = c.ldg(g + bsk_index(1 - p, half * 16 + q, p, t));
                }
#if FHESTR_BR_L1PF
#pragma unroll
                for (int q = kBskPrefetch; q < 16; q++) {
                    c.prefetch_l1(g + bsk_index(p, half * 16 + q, p, t));
                    c.prefetch_l1(g + bsk_index(1 - p, half * 16 + q, p, t));
                }
#endif
                c.log_mark(10 + half * 4);
                c.pair_sync();
                c.log_mark(11 + half * 4);
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const int k2 = half * 16 + q;
                    const cplx gs = q < kBskPrefetch ? gsv[q < kBskPrefetch ? q : 0] : c.ldg(g + bsk_index(p, k2, p, t));
                    const cplx go = q < kBskPrefetch ? gov[q < kBskPrefetch ? q : 0] : c.ldg(g + bsk_index(1 - p, k2, p, t));
                    const cplx v = xp[q * 32 + t];
                    const cplx s = cmul(cplx{A[k2], B[k2]}, gs);
                    const double zr = fma(v.x, go.x, fma(-v.y, go.y, s.x));
                    const double zi = fma(v.x, go.y, fma(v.y, go.x, s.y));
                    A[k2] = zi; B[k2] = zr;      // swapped from here on
                }
                c.log_mark(12 + half * 4);
                c.pair_sync();
                c.log_mark(13 + half * 4);
            }
            c.log_mark(5);
        }
    }
#endif
    c.log_mark(8);
#pragma unroll
    for (int n1 = 0; n1 < 32; n1++) {
        const int j = 32 * n1 + t;
        double zr = B[n1], zi = A[n1];
        if (n1 != 0 && FHESTR_BR_ABLATE < 6) {
            const double cs = FHESTR_TWIST(2 * n1), sn = FHESTR_TWIST(2 * n1 + 1);
            zr = fma(B[n1], cs, A[n1] * sn);
            zi = fma(A[n1], cs, -(B[n1] * sn));
        }
#if FHESTR_BR_AREG
        a[n1] += torus32_conv(zr, n1);
        a[32 + n1] += torus32_conv(zi, n1);
#else
        a[n1] = acc[j] + torus32_conv(zr, n1);
        a[32 + n1] = acc[j + kM] + torus32_conv(zi, n1);
#endif
        acc[j] = a[n1];
        acc[j + kM] = a[32 + n1];
    }
    c.syncwarp();
    c.log_mark(9);
}

// Forward transform of one standard-domain GGSW polynomial (key conversion, once per key)
template <class Ctx>
FHE_HD void bsk_poly_forward(Ctx& c, const u64* poly, cplx* out_step, int row, int col, const cplx* tf) {
    const int t = c.lane();
    const double sc = 1.0 / (4294967296.0 * (double)kM);   // 2^-64 (u64 -> turns) * 2^32 (turns -> acc_t ulps) / M
    double re[32], im[32];
#pragma unroll
    for (int n1 = 0; n1 < 32; n1++) {
        const int j = 32 * n1 + t;
        re[n1] = (double)(i64)poly[j] * sc;
        im[n1] = (double)(i64)poly[j + kM] * sc;
    }
    forward1024(c, re, im, tf);
#pragma unroll
    for (int k2 = 0; k2 < 32; k2++) out_step[bsk_index(row, k2, col, t)] = cplx{re[k2], im[k2]};
}

// ---------------------------------------------------------------------------------------------
// Whole blind rotation for one PBS, executed by a pair of warps (poly 0 = mask, poly 1 = body).
// Fuses the modulus switch (A.6) in front and the sample extraction (A.8) behind.
struct BrJobView {
    const u64* ks;        // [n+1] keyswitched LWE (small key), u64
    const u64* lut;       // [N] body polynomial of the trivial GLWE accumulator
    const u64* init_acc;  // optional [2][N]: start from this GLWE instead of X^{-b~} * LUT (test hook)
    u64* out_lwe;         // optional [N+1]: sample-extracted LWE under the big key
    u64* out_acc;         // optional [2][N]: raw accumulator (test hook)
    int n;                // number of CMUX steps (small LWE dimension)
    // multi-GPU: the same block of the PEER arenas (NVLink P2P stores); the sample-extract epilogue writes the
    // result everywhere at once, so no collective follows the kernel
    u64* out_lwe_peer[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int n_peers = 0;
    u64 post = 0;         // added to the body of the extracted LWE (half-step LUTs: + delta/2, see fhestr_lut_register)
};

constexpr int kMaxPeers = 7;

template <class Ctx>
FHE_HD void br_thread_main(Ctx& c, const BrJobView& job, const cplx* bsk, const cplx* tf, const cplx* ti) {
    const int t = c.lane();
    const int p = c.poly();
    const int n = job.n;
    uint16_t* at = c.atilde();
    for (int idx = p * 32 + t; idx <= n; idx += 64) at[idx] = (uint16_t)modswitch_2N(job.ks[idx]);
    c.pair_sync();
    acc_t* acc = c.acc();
    acc_t a[64];
    {
        const int e0 = (2 * kN - (int)at[n]) & (2 * kN - 1);
#pragma unroll
        for (int m = 0; m < 64; m++) {
            const int j = 32 * m + t;  // m < 32: j = 32 n1 + t ; m >= 32: j = 32 n1 + t + M
            u64 v;
            if (job.init_acc) v = job.init_acc[p * kN + j];
            else v = (p == 1) ? rot_coef(job.lut, j, e0) : (u64)0;
            a[m] = acc_from_u64(v);   // LUT words are multiples of 2^59: exact
            acc[j] = a[m];
        }
    }
    c.syncwarp();
#if FHESTR_BR_COMPACT
    cplx twr[kTwReg > 0 ? kTwReg : 1];
#pragma unroll
    for (int r = 0; r < kTwReg; r++) twr[r] = c.ldg(tf + r * 32 + t);
    (void)ti;
#endif
    for (int i = 0; i < n; i++) {
        const int e = at[i];
        if (e == 0) {          // X^0 * ACC - ACC == 0: the external product contributes exactly nothing
            for (int q = 0; q < kBatonPhases; q++) { c.fp_acquire(); c.fp_release(); }   // keep the partner's turns
#if FHESTR_BR_RING
#pragma unroll 1
            for (int j = 0; j < kKeyChunks; j++) {      // keep the CTA's key ring turning
                (void)c.key_wait(i, j);
                c.key_done(i, j);
                if (j % (16 / kKeyChunkRows) != 0) c.key_duty(i, j - 1);
                if (j % (16 / kKeyChunkRows) == 16 / kKeyChunkRows - 1) c.key_duty(i, j);
            }
#endif
            continue;
        }
#if FHESTR_BR_COMPACT
        cmux_step_compact(c, a, e, i, bsk + (size_t)i * kBskStepElems, tf, twr);
#else
        cmux_step(c, a, e, i, bsk + (size_t)i * kBskStepElems, tf, ti);
#endif
    }
    c.fp_finish();
    c.log_mark(-1);
    if (job.out_acc) {
#pragma unroll
        for (int m = 0; m < 64; m++) job.out_acc[p * kN + 32 * m + t] = acc_to_u64(FHESTR_BR_ACC_SMEM ? acc[32 * m + t] : a[m]);
    }
    if (job.out_lwe) {
        if (p == 0) {
#pragma unroll
            for (int m = 0; m < 64; m++) {
                const int j = 32 * m + t;
                const u64 w = acc_to_u64((j == 0) ? acc[0] : (acc_t)0 - acc[kN - j]);
                job.out_lwe[j] = w;
                for (int r = 0; r < job.n_peers; r++) job.out_lwe_peer[r][j] = w;
            }
        } else if (t == 0) {
            job.out_lwe[kN] = acc_to_u64(acc[0]) + job.post;
            for (int r = 0; r < job.n_peers; r++) job.out_lwe_peer[r][kN] = acc_to_u64(acc[0]) + job.post;
        }
    }
}

// Twiddle tables shared by the engine and the host emulation:
//   tf[k1*32 + n2] = exp(+i pi n2 (1-4 k1) / N),  ti[n2*32 + k1] = conj of the same value (the compact loop reads tf
//   only: its inverse passes run on swapped real / imaginary parts and multiply after the transpose)
inline void make_twiddles(cplx* tf, cplx* ti) {
    for (int k1 = 0; k1 < 32; k1++)
        for (int n2 = 0; n2 < 32; n2++) {
            const long double ang = 3.14159265358979323846264338327950288L * (long double)(n2 * (1 - 4 * k1)) / (long double)kN;
            const double cr = (double)cosl(ang), ci = (double)sinl(ang);
            tf[k1 * 32 + n2] = cplx{cr, ci};
            ti[n2 * 32 + k1] = cplx{cr, -ci};
        }
}

}  // namespace fhestr
