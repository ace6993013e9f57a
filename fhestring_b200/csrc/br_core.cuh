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
//   twist along n1: c_n *= exp(i pi n1 / 64)                          (constant bank, kFft32Twist*)
//   pass 1 (lane = n2, registers n1 -> k1): plain DFT-32               (fft32_dft)
//   twiddle Tf(k1,n2) = exp(i pi n2 (1-4 k1)/N), transpose through shared memory
//   pass 2 (lane = k1, registers n2 -> k2): plain DFT-32               (fft32_dft)
// ONE WARP owns one polynomial: 32 complex points per lane live in registers, the only exchange inside
// a transform is one 32x32 transpose of complex words through a 16.5 KB padded buffer, and
// only __syncwarp is needed.  The two warps of a PBS (mask polynomial, body polynomial) swap one
// spectrum per step through the same buffers (pair barrier).
//
// The CMUX step is a ROLLED loop of four passes over ONE copy of the DFT-32 codelet (cmux_step): the inverse
// DFT of x is swap(DFT(swap(x))) and a multiplication by conj(w) in the swapped domain is a multiplication by w,
// so both directions run the same pass and the same twiddle code, and the inverse multiplies AFTER its transpose
// (lane = n2, register k1 again), so that one set of 32 twiddles per lane serves both directions.  Those
// twiddles live in TENSOR MEMORY (Ctx::tw_ld: 128 lane-private TMEM columns, written once per PBS) instead of
// being re-read from L1 / L2 in every transform.  Measured against round 1's straight-line step with four different
// codelets and twiddles from L1 (profiles/r2_compact_tmem.md): 43.3 instead of 46.5 ms per 4096 PBS; the rolled loop
// needs 168 instead of 255 registers, which is what pays for the early twiddle request and the deeper key prefetch.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define FHE_HD __host__ __device__ __forceinline__
#else
#define FHE_HD inline
#endif
#include "fft32_gen.cuh"

// Compile-time knobs of the step (A/B-measured on one B200, profiles/r2_compact_tmem.md; build.py --variant builds any
// other setting as a second library for FHESTR_ENGINE_LIB):
//   FHESTR_BR_PREFETCH = key rows per half-step requested ahead of the pair barrier (6 = 48 registers)
//   FHESTR_BR_CVT_FP64 = m > 0: every m-th torus conversion of the epilogue runs on the FP64 pipe (four FP64
//       instructions, bit-identical to the F2I) instead of the conversion unit; 0 = none (round 1's kernel: 4)
#ifndef FHESTR_BR_PREFETCH
#define FHESTR_BR_PREFETCH 6
#endif
#ifndef FHESTR_BR_CVT_FP64
#define FHESTR_BR_CVT_FP64 0
#endif

namespace fhestr {

typedef unsigned long long u64;
typedef long long i64;

constexpr int kN = 2048;          // polynomial size
constexpr int kM = 1024;          // complex points
constexpr int kXPad = 33;         // transpose row stride (complex words): conflict-free 128-bit column reads
constexpr int kXbufDoubles = 32 * kXPad;  // 1056 words
constexpr int kWarpXbufDoubles = 2 * kXbufDoubles;  // per warp: ONE matrix of 32 x 33 complex words = 16 896 B
constexpr int kPbsBaseLog = 23;
constexpr int kBskPrefetch = FHESTR_BR_PREFETCH;
constexpr int kTwChunks = 4;      // a lane's 32 twiddles = 128 words are read 8 twiddles (32 words) at a time

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
// memory (16 KiB per PBS instead of 32), its traffic and the integer work of every step (A/B:
// profiles/r2_k3_accuracy.md).
typedef uint32_t acc_t;

FHE_HD acc_t acc_from_u64(u64 x) { return (acc_t)((x + (1ull << 31)) >> 32); }
FHE_HD u64 acc_to_u64(acc_t a) { return (u64)a << 32; }

// A.4 signed decomposition, one level of 23 bits, of a 32-bit torus difference: digit in [-2^22, 2^22); the tie
// x = 2^31 - 256 .. 2^31 - 1 stays -2^22 (the same torus value as +2^22, |digit| unchanged)
FHE_HD double digit23(acc_t x) { return (double)(((int32_t)(x + (1u << 8))) >> 9); }

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

#ifdef __CUDA_ARCH__
#define FHESTR_TWIST(i) kFft32TwistDev[i]
#else
#define FHESTR_TWIST(i) kFft32TwistHost[i]
#endif

// 32x32 transpose of one complex point per (lane, register) through the warp's padded buffer: 16-byte words, row stride
// 33 words, so both the row-wise stores and the column-wise loads are conflict-free per quarter-warp.  (Up to the first
// half of round 2 the real and imaginary parts went through an 8.25 KB matrix one after the other: twice the
// shared-memory instructions and warp barriers for the same wavefronts; 43.27 -> 42.88 ms per 4096 PBS.)
template <class Ctx>
FHE_HD void transpose32(Ctx& c, double (&re)[32], double (&im)[32]) {
    const int t = c.lane();
    cplx* buf = reinterpret_cast<cplx*>(c.xbuf());
#pragma unroll
    for (int r = 0; r < 32; r++) buf[r * kXPad + t] = cplx{re[r], im[r]};
    c.syncwarp();
#pragma unroll
    for (int r = 0; r < 32; r++) { const cplx v = buf[t * kXPad + r]; re[r] = v.x; im[r] = v.y; }
    c.syncwarp();
}

// Multiply the 32 points of this lane by its inter-pass twiddles tf[r*32 + lane], r = 0..31.  The words come from the
// context 8 twiddles at a time (Ctx::tw_ld requests chunk ch into 32 registers, Ctx::tw_wait makes them readable): on
// the device from tensor memory, the next chunk in flight while this one is used; the first chunk was requested by the
// caller before the pass whose results it multiplies, so its latency is never exposed.
template <class Ctx>
FHE_HD void twiddle32(Ctx& c, double (&A)[32], double (&B)[32], const cplx* tf, uint32_t (&wa)[32]) {
    uint32_t wb[32];
    c.tw_wait(wa);
#pragma unroll
    for (int ch = 0; ch < kTwChunks; ch++) {
        uint32_t (&cur)[32] = (ch & 1) ? wb : wa;
        uint32_t (&nxt)[32] = (ch & 1) ? wa : wb;
        if (ch + 1 < kTwChunks) c.tw_ld(ch + 1, nxt, tf);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int r = ch * 8 + j;
            const cplx w = cplx{c.tw_word(cur[4 * j], cur[4 * j + 1]), c.tw_word(cur[4 * j + 2], cur[4 * j + 3])};
            const cplx z = cmul(cplx{A[r], B[r]}, w);
            A[r] = z.x; B[r] = z.y;
        }
        if (ch + 1 < kTwChunks) c.tw_wait(nxt);
    }
}

// Fourier BSK layout (engine-private, produced once by the key-conversion kernel):
//   g[ ((((step*2 + row)*32 + k2)*2 + col)*32 + k1 ]   complex f64, k = k1 + 32 k2
// scaled by 2^-32 / M so that the inverse transform directly yields units of 2^-32 turns (acc_t ulps).
constexpr int kBskStepElems = 2 * 32 * 2 * 32;  // 4096 complex = 64 KiB per CMUX step
FHE_HD int bsk_index(int row, int k2, int col, int k1) { return ((row * 32 + k2) * 2 + col) * 32 + k1; }

// One CMUX step for this warp's polynomial:  ACC += GGSW (x) (X^e * ACC - ACC).
//   a[0..31]  = ACC[32 n1 + lane], a[32..63] = ACC[32 n1 + lane + M]  (registers, in/out)
//   c.acc()   = this polynomial's accumulator in shared memory (same values), updated on exit
// (A, B) hold (real, imaginary) parts during the forward passes and (imaginary, real) during the inverse ones.
//   before the loop: gather + digits + twist exp(+i pi n1/64)
//   pass 0: DFT over n1, twiddle, transpose
//   pass 1: DFT over n2, spectrum exchange + GGSW product (results written swapped)
//   pass 2: DFT over k2 (inverse), transpose, twiddle (lane = n2, register k1 again)
//   pass 3: DFT over k1 (inverse)
//   after the loop: untwist exp(-i pi n1/64), accumulate
template <class Ctx>
FHE_HD void cmux_step(Ctx& c, acc_t (&a)[64], int e, const cplx* g, const cplx* tf) {
    const int t = c.lane();
    const int p = c.poly();
    acc_t* acc = c.acc();
    double A[32], B[32];
    // The gather consumes the register copy a[] BEFORE the rolled loop and the accumulation rewrites it AFTER it, so
    // that a[] is dead across the loop (inside it, the compiler would have to keep it alive over every pass).
    // Byte offset of coefficient (t - e) mod 2N in the 2N-word negacyclic extension; row n1 adds 128 bytes.  Bit 13
    // of the running offset is the sign, bits 2..12 the word inside the 8 KiB buffer (c.acc_ld_rot).
    {
        const uint32_t x0 = (uint32_t)((t - e) & (2 * kN - 1)) << 2;
#pragma unroll
        for (int n1 = 0; n1 < 32; n1++) {
            const double dr = digit23(c.acc_ld_rot(x0 + 128u * n1) - a[n1]);
            const double di = digit23(c.acc_ld_rot(x0 + 128u * n1 + 4096u) - a[32 + n1]);
            if (n1 == 0) { A[0] = dr; B[0] = di; }
            else {
                const double cs = FHESTR_TWIST(2 * n1), sn = FHESTR_TWIST(2 * n1 + 1);
                A[n1] = fma(dr, cs, -(di * sn));
                B[n1] = fma(dr, sn, di * cs);
            }
        }
    }
    uint32_t w0[32];
#pragma unroll 1
    for (int it = 0; it < 4; it++) {
        if ((it & 1) == 0) c.tw_ld(0, w0, tf);
        fft32_dft(A, B);
        if (it == 0) {
            twiddle32(c, A, B, tf, w0);
            transpose32(c, A, B);
        } else if (it == 2) {
            transpose32(c, A, B);
            twiddle32(c, A, B, tf, w0);
        } else if (it == 1) {
            // Fourier-domain GGSW product.  The two warps swap their spectra through the transpose buffers, 16 spectrum
            // rows at a time (the second half's key rows are requested while the first half multiplies), and each
            // forms ITS output polynomial completely:
            //     out_p = D_p * G[p][p] + D_(1-p) * G[1-p][p]
            // the second product accumulates with FMAs, so the step costs 8 FP64 operations per point instead of 10
            cplx* xo = reinterpret_cast<cplx*>(c.xbuf());
            const cplx* xp = reinterpret_cast<const cplx*>(c.xbuf_partner());
#pragma unroll
            for (int half = 0; half < 2; half++) {
#pragma unroll
                for (int q = 0; q < 16; q++) xo[q * 32 + t] = cplx{A[half * 16 + q], B[half * 16 + q]};
                // the first key words of this half do not depend on the partner: request them BEFORE the barrier so
                // that their L2 latency overlaps the wait
                cplx gsv[kBskPrefetch], gov[kBskPrefetch];
#pragma unroll
                for (int q = 0; q < kBskPrefetch; q++) {
                    gsv[q] = c.ldg(g + bsk_index(p, half * 16 + q, p, t));
                    gov[q] = c.ldg(g + bsk_index(1 - p, half * 16 + q, p, t));
                }
                c.pair_sync();
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
                c.pair_sync();
            }
        }
    }
    // accumulate into the torus accumulator; keep the new words in registers for the next step's gather
#pragma unroll
    for (int n1 = 0; n1 < 32; n1++) {
        const int j = 32 * n1 + t;
        double zr = B[n1], zi = A[n1];
        if (n1 != 0) {
            const double cs = FHESTR_TWIST(2 * n1), sn = FHESTR_TWIST(2 * n1 + 1);
            zr = fma(B[n1], cs, A[n1] * sn);
            zi = fma(A[n1], cs, -(B[n1] * sn));
        }
        a[n1] = acc[j] + torus32_conv(zr, n1);
        a[32 + n1] = acc[j + kM] + torus32_conv(zi, n1);
        acc[j] = a[n1];
        acc[j + kM] = a[32 + n1];
    }
    c.syncwarp();
}

// forward transform of the 32 complex points held by this lane (lane = n2, register n1) into the
// spectrum layout (lane = k1, register k2): the same passes as the first half of cmux_step
template <class Ctx>
FHE_HD void forward1024(Ctx& c, double (&re)[32], double (&im)[32], const cplx* tf) {
#pragma unroll
    for (int n1 = 1; n1 < 32; n1++) {
        const double cs = FHESTR_TWIST(2 * n1), sn = FHESTR_TWIST(2 * n1 + 1);
        const double xr = re[n1], xi = im[n1];
        re[n1] = fma(xr, cs, -(xi * sn));
        im[n1] = fma(xr, sn, xi * cs);
    }
    uint32_t w0[32];
    c.tw_ld(0, w0, tf);
    fft32_dft(re, im);
    twiddle32(c, re, im, tf, w0);
    transpose32(c, re, im);
    fft32_dft(re, im);
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
FHE_HD void br_thread_main(Ctx& c, const BrJobView& job, const cplx* bsk, const cplx* tf) {
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
    for (int i = 0; i < n; i++) {
        const int e = at[i];
        if (e == 0) continue;   // X^0 * ACC - ACC == 0: the external product contributes exactly nothing
        cmux_step(c, a, e, bsk + (size_t)i * kBskStepElems, tf);
    }
    if (job.out_acc) {
#pragma unroll
        for (int m = 0; m < 64; m++) job.out_acc[p * kN + 32 * m + t] = acc_to_u64(a[m]);
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

// Twiddle table shared by the engine and the host emulation:  tf[k1*32 + n2] = exp(+i pi n2 (1-4 k1) / N)
inline void make_twiddles(cplx* tf) {
    for (int k1 = 0; k1 < 32; k1++)
        for (int n2 = 0; n2 < 32; n2++) {
            const long double ang = 3.14159265358979323846264338327950288L * (long double)(n2 * (1 - 4 * k1)) / (long double)kN;
            tf[k1 * 32 + n2] = cplx{(double)cosl(ang), (double)sinl(ang)};
        }
}

}  // namespace fhestr
