// br_quad.cuh -- blind rotation with FOUR warps per PBS (two per polynomial), written once and compiled twice
// like br_core.cuh: by nvcc into the sm_100a kernel (blind_rotate_quad.cu) and by g++ into the host emulation
// (tests/emu/br_quad_emu.cpp).
//
// Why: with one warp per polynomial (br_core.cuh) a thread carries 32 complex points = 128 data registers, the
// register file holds 8 warps per SM and every warp runs latency-bound at ~0.2 IPC (r1 ncu: FP64 pipe 48 %
// busy, shared-memory pipe 55 %, the two never overlapped).  Here a thread carries 16 points, fits 128
// registers, and 16 warps per SM hide each other's DFMA / shared-memory / L2 latencies; the four warps of a PBS
// sit on the four SM sub-partitions, so one PBS also finishes a step in about half the time.
//
// Transform.  Folded polynomial c_n = p_n + i p_{n+M}, M = 1024; spectrum X_k = sum_n c_n zeta^n W^{nk}
// (zeta = exp(i pi/N), W = exp(-2 pi i/M)).  With n = 64 n1 + 4 m + q and k = k1 + 16 (j + 16 r):
//     zeta^n W^{nk} = [D(n1) W16^{n1 k1}] . T(k1, 4m+q) . W16^{m j} . W64^{q j} . W4^{q r}
//     D(n1) = exp(i pi n1/32),  T(k1, n2) = exp(i pi n2 (1 - 4 k1)/N)
// 64 threads tau per polynomial, 16 registers each:
//   A   thread tau = n2,            registers n1 -> k1 : fft16_twisted (D merged), x T(k1, n2)
//   X1  exchange through shared memory (rows k1, padded to 68 points: conflict-free both ways)
//   B   thread tau = 4 k1 + q,      registers m  -> j  : fft16_plain
//   X2  exchange inside groups of 4 lanes (same rows, __syncwarp only)
//   C   thread tau = 4 k1 + s,      registers (jj, q) -> (jj, r), j = s + 4 jj : x W64^{q j}, 4-point DFT
// The inverse is the TRANSPOSED algorithm (the kernel matrix is applied from the k side: C', X2', B', X1', A')
// in the swapped domain swap(a + ib) = b + ia, where conjugation is free:  swap(sum_k Z_k conj(Phi_kn)) =
// sum_k swap(Z)_k Phi_kn  -- so it uses the same forward codelets and tables.  scripts/proto/quad_fft_proto.py
// is the numpy statement of exactly this index algebra.
#pragma once
#include "br_core.cuh"
#include "fft16_gen.cuh"

namespace fhestr {

constexpr int kQRow = 68;                        // exchange row stride (complex points): 64 + 4
constexpr int kQExchCplx = 16 * kQRow;           // 1088 complex = 17 408 B per polynomial
constexpr int kQBskStepElems = 2 * 2 * 16 * 64;  // 4096 complex = 64 KiB per CMUX step
// Fourier BSK layout of this kernel: g[((row*2 + col)*16 + reg)*64 + tau], reg = 4 jj + r,
// spectrum index k = k1 + 16 (s + 4 jj + 16 r) with tau = 4 k1 + s; scaled by 2^-32 / M like br_core's
FHE_HD int qbsk_index(int row, int col, int reg, int tau) { return ((row * 2 + col) * 16 + reg) * 64 + tau; }

struct QuadTables {
    const cplx* tq;    // [k1*64 + n2]   = T(k1, n2)
    const cplx* tqt;   // [m*64 + tau]   = T(tau >> 2, 4 m + (tau & 3))
    const cplx* w64;   // [j*4 + q]      = exp(-2 pi i q j / 64)
};

inline void make_quad_tables(cplx* tq, cplx* tqt, cplx* w64) {
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int k1 = 0; k1 < 16; k1++)
        for (int n2 = 0; n2 < 64; n2++) {
            const long double ang = pi * (long double)(n2 * (1 - 4 * k1)) / (long double)kN;
            const cplx v{(double)cosl(ang), (double)sinl(ang)};
            tq[k1 * 64 + n2] = v;
            tqt[(n2 >> 2) * 64 + (4 * k1 + (n2 & 3))] = v;   // m = n2 >> 2, q = n2 & 3, tau = 4 k1 + q
        }
    for (int j = 0; j < 16; j++)
        for (int q = 0; q < 4; q++) {
            const long double ang = -2.0L * pi * (long double)(q * j) / 64.0L;
            w64[j * 4 + q] = cplx{(double)cosl(ang), (double)sinl(ang)};
        }
}

// 4-point DFT with W4 = -i: out[r] = sum_q in[q] (-i)^{q r}
FHE_HD void dft4(double& r0, double& i0, double& r1, double& i1, double& r2, double& i2, double& r3, double& i3) {
    const double ar = r0 + r2, ai = i0 + i2, br = r0 - r2, bi = i0 - i2;
    const double cr = r1 + r3, ci = i1 + i3, dr = r1 - r3, di = i1 - i3;
    r0 = ar + cr; i0 = ai + ci;
    r2 = ar - cr; i2 = ai - ci;
    r1 = br + di; i1 = bi - dr;    // b + (-i) d
    r3 = br - di; i3 = bi + dr;    // b + (+i) d
}

// stages A (after the caller filled re/im in layout A) .. C: spectrum in registers [4 jj + r]
template <class Ctx>
FHE_HD void quad_forward(Ctx& c, double (&re)[16], double (&im)[16], const QuadTables& tb) {
    const int tau = c.tau();
    const int k1p = tau >> 2, q = tau & 3;
    cplx* E = c.exch();
    fft16_twisted(re, im);
#pragma unroll
    for (int k1 = 0; k1 < 16; k1++) {
        const cplx z = cmul(cplx{re[k1], im[k1]}, c.ldg(tb.tq + k1 * 64 + tau));
        E[k1 * kQRow + tau] = z;
    }
    c.poly_sync();
#pragma unroll
    for (int m = 0; m < 16; m++) {
        const cplx z = E[k1p * kQRow + 4 * m + q];
        re[m] = z.x; im[m] = z.y;
    }
    fft16_plain(re, im);
    c.syncwarp();  // rows k1p of this warp are read by this warp only: safe to overwrite them now
#pragma unroll
    for (int j = 0; j < 16; j++) E[k1p * kQRow + q * 17 + j] = cplx{re[j], im[j]};
    c.syncwarp();
    const int s = q;
#pragma unroll
    for (int jj = 0; jj < 4; jj++) {
        const int j = s + 4 * jj;
#pragma unroll
        for (int qq = 0; qq < 4; qq++) {
            cplx z = E[k1p * kQRow + qq * 17 + j];
            if (qq) z = cmul(z, c.ldg(tb.w64 + j * 4 + qq));
            re[4 * jj + qq] = z.x; im[4 * jj + qq] = z.y;
        }
        dft4(re[4 * jj], im[4 * jj], re[4 * jj + 1], im[4 * jj + 1], re[4 * jj + 2], im[4 * jj + 2], re[4 * jj + 3], im[4 * jj + 3]);
    }
}

// the transposed algorithm: registers [4 jj + r] (thread tau = 4 k1 + s) -> layout A (thread n2, registers n1);
// the caller must have separated earlier uses of the exchange buffer with a barrier
template <class Ctx>
FHE_HD void quad_transposed(Ctx& c, double (&re)[16], double (&im)[16], const QuadTables& tb) {
    const int tau = c.tau();
    const int k1p = tau >> 2, q = tau & 3;
    const int s = q;
    cplx* E = c.exch();
#pragma unroll
    for (int jj = 0; jj < 4; jj++) {
        const int j = s + 4 * jj;
        dft4(re[4 * jj], im[4 * jj], re[4 * jj + 1], im[4 * jj + 1], re[4 * jj + 2], im[4 * jj + 2], re[4 * jj + 3], im[4 * jj + 3]);
#pragma unroll
        for (int qq = 0; qq < 4; qq++) {
            cplx z{re[4 * jj + qq], im[4 * jj + qq]};
            if (qq) z = cmul(z, c.ldg(tb.w64 + j * 4 + qq));
            E[k1p * kQRow + qq * 17 + j] = z;
        }
    }
    c.syncwarp();
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const cplx z = E[k1p * kQRow + q * 17 + j];
        re[j] = z.x; im[j] = z.y;
    }
    fft16_plain(re, im);
    c.syncwarp();
#pragma unroll
    for (int m = 0; m < 16; m++) {
        const cplx z = cmul(cplx{re[m], im[m]}, c.ldg(tb.tqt + m * 64 + tau));
        E[k1p * kQRow + 4 * m + q] = z;
    }
    c.poly_sync();
#pragma unroll
    for (int k1 = 0; k1 < 16; k1++) {
        const cplx z = E[k1 * kQRow + tau];
        re[k1] = z.x; im[k1] = z.y;
    }
    fft16_plain_twist(re, im);
}

// One CMUX step for the 64 threads of one polynomial:  ACC += GGSW (x) (X^e * ACC - ACC)
template <class Ctx>
FHE_HD void quad_cmux_step(Ctx& c, int e, const cplx* g, const QuadTables& tb) {
    const int tau = c.tau();
    const int p = c.poly();
    acc_t* acc = c.acc();
    double re[16], im[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) {
        const int j = 64 * n1 + tau;
        re[n1] = digit23(rot_coef(acc, j, e) - acc[j]);
        im[n1] = digit23(rot_coef(acc, j + kM, e) - acc[j + kM]);
    }
    quad_forward(c, re, im, tb);
    // Fourier-domain GGSW product: our spectrum times row p of the GGSW; the column-(1-p) half goes to the
    // partner polynomial's thread with the same (tau, register) through our exchange buffer
    cplx* X = c.exch();
    const cplx* Xp = c.exch_partner();
    double ore[16], oim[16];
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const cplx d{re[r], im[r]};
        const cplx o = cmul(d, c.ldg(g + qbsk_index(p, 1 - p, r, tau)));
        const cplx sf = cmul(d, c.ldg(g + qbsk_index(p, p, r, tau)));
        ore[r] = o.x; oim[r] = o.y;
        re[r] = sf.x; im[r] = sf.y;
    }
    c.poly_sync();   // the other warp of this polynomial has finished stage C reads of its rows
#pragma unroll
    for (int r = 0; r < 16; r++) X[r * 64 + tau] = cplx{ore[r], oim[r]};
    c.cta_sync();
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const cplx v = Xp[r * 64 + tau];
        const double sx = re[r] + v.x, sy = im[r] + v.y;
        re[r] = sy; im[r] = sx;   // swapped domain from here on
    }
    c.cta_sync();    // the partner has read our buffer: it may be reused
    quad_transposed(c, re, im, tb);
    // un-swap (real part = im, imaginary part = re), round to the 32-bit torus, accumulate
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) {
        const int j = 64 * n1 + tau;
        acc[j] += torus32_from_double(im[n1]);
        acc[j + kM] += torus32_from_double(re[n1]);
    }
    c.poly_sync();   // the new accumulator is visible to both warps before the next rotation reads it
}

// Forward transform of one standard-domain GGSW polynomial by 64 threads (key conversion, once per key)
template <class Ctx>
FHE_HD void quad_bsk_poly_forward(Ctx& c, const u64* poly, cplx* out_step, int row, int col, const QuadTables& tb) {
    const int tau = c.tau();
    const double sc = 1.0 / (4294967296.0 * (double)kM);
    double re[16], im[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) {
        const int j = 64 * n1 + tau;
        re[n1] = (double)(i64)poly[j] * sc;
        im[n1] = (double)(i64)poly[j + kM] * sc;
    }
    quad_forward(c, re, im, tb);
#pragma unroll
    for (int r = 0; r < 16; r++) out_step[qbsk_index(row, col, r, tau)] = cplx{re[r], im[r]};
    c.poly_sync();   // the exchange buffer is free for the next polynomial
}

// Whole blind rotation of one PBS by 128 threads (2 polynomials x 64 threads); fuses the modulus switch in
// front and the sample extraction behind (same contract as br_thread_main)
template <class Ctx>
FHE_HD void quad_thread_main(Ctx& c, const BrJobView& job, const cplx* bsk, const QuadTables& tb) {
    const int tau = c.tau();
    const int p = c.poly();
    const int n = job.n;
    uint16_t* at = c.atilde();
    for (int idx = p * 64 + tau; idx <= n; idx += 128) at[idx] = (uint16_t)modswitch_2N(job.ks[idx]);
    c.cta_sync();
    acc_t* acc = c.acc();
    {
        const int e0 = (2 * kN - (int)at[n]) & (2 * kN - 1);
        for (int m = 0; m < 32; m++) {
            const int j = 64 * m + tau;
            u64 v;
            if (job.init_acc) v = job.init_acc[p * kN + j];
            else v = (p == 1) ? rot_coef(job.lut, j, e0) : (u64)0;
            acc[j] = acc_from_u64(v);
        }
    }
    c.poly_sync();
    for (int i = 0; i < n; i++) {
        const int e = at[i];
        if (e == 0) continue;  // X^0 * ACC - ACC == 0: the external product contributes exactly nothing
        quad_cmux_step(c, e, bsk + (size_t)i * kQBskStepElems, tb);
    }
    if (job.out_acc)
        for (int m = 0; m < 32; m++) job.out_acc[p * kN + 64 * m + tau] = acc_to_u64(acc[64 * m + tau]);
    if (job.out_lwe) {
        if (p == 0) {
            for (int m = 0; m < 32; m++) {
                const int j = 64 * m + tau;
                job.out_lwe[j] = acc_to_u64((j == 0) ? acc[0] : (acc_t)0 - acc[kN - j]);
            }
        } else if (tau == 0) {
            job.out_lwe[kN] = acc_to_u64(acc[0]);
        }
    }
}

}  // namespace fhestr
