// br_wide.cuh -- the LATENCY form of the blind rotation: one PBS spread over 128 threads (one warp on each of the
// four sub-partitions of an SM), used for dependency levels with fewer jobs than the GPU has SMs, where the
// throughput kernel (br_core.cuh: one PBS per pair of warps, 742 x 6.6 us on an idle SM) leaves the machine idle.
// Written once and compiled twice like br_core.cuh: by nvcc into blind_rotate_wide.cu and by g++ into the host
// emulation (tests/emu/br_wide_emu.cpp, 128 std::threads + std::barrier) that checks the index logic on the CPU.
//
// Replaces (reference side): the same tfhe-rs blind rotation behind every PBS issued from
// /root/reference/src/ciphertext/fheasciichar.rs:36-102 -- the serial dependency chain of
// /root/reference/src/server_key/mod.rs:151-182 (contains: window matches -> AND -> OR tree) is what makes the
// single-PBS latency matter.
//
// Transform.  A polynomial p of degree < 2048 is folded to c(x) = sum_{n<1024} (p_n + i p_{n+1024}) x^n and evaluated
// at the 1024 roots of x^1024 = i (each is a root of x^2048 = -1; the other half are their conjugates).  The
// evaluation is a recursive reduction  c mod (x^m - rho)  ->  c mod (x^(m/2) -/+ sqrt(rho))  (butterfly
// (a, b) -> (a + r b, a - r b), r = sqrt(rho)), ten levels, grouped 3 + 3 + 1 + 3:
//   stage 1  thread n0 = t holds coefficients 128 k + n0 (k = 0..7), modulus angle pi/2 for every thread
//            -> coefficient n0 of the eight residues  R_j mod (x^128 - rho_j)
//   exchange 1 (shared memory): thread (j, q) = (t >> 4, t & 15) collects coefficients 16 m + q of R_j
//   stage 2  -> coefficient q of the residues  S_jl mod (x^16 - rho_jl)
//   exchange 2: thread (jl, g) = (t >> 1, t & 1) reads all 16 coefficients of S_jl and forms its half
//            v_k = w_k +/- r w_(k+8)  (the one level that would otherwise need a third exchange: both threads
//            of a pair read the same 16 words, the second read is a broadcast)
//   stage 3  -> the eight values of c at the roots below that node: spectrum point (t, u)
// Every thread holds the 8 spectrum points (t, u) of BOTH polynomials of the GLWE accumulator, so the GGSW
// product needs no exchange at all (the throughput kernel swaps spectra between its two warps), and the two
// independent polynomials give a single warp the instruction-level parallelism that hides its own latencies.
// The inverse mirrors the forward (un-normalised: every level doubles, 2^10 = the 1/M the Fourier key carries).
// The spectrum ORDER is private to this kernel: the key is transformed by the same code (wide_bsk_poly_forward).
//
// The 64 KiB Fourier-key tile of a step is streamed into shared memory by ONE bulk-TMA copy per step
// (cp.async.bulk + mbarrier, double-buffered: the tile of step i+1 lands while step i computes), which is the
// north-star design for the key ("BSK streamed via TMA into SMEM"); in the throughput kernel the four PBS of an
// SM are at different steps, so there the tile comes through L1 instead.
#pragma once
#include "br_core.cuh"

namespace fhestr {

constexpr int kWT = 128;                       // threads per PBS
constexpr int kWKeyTile = 4 * 8 * kWT;         // complex words per step: [row*2+col][u][t] = 64 KiB
constexpr int kWX1 = 1024;                     // exchange 1: [p][j*128 + n0]
constexpr int kWX2 = 64 * 17;                  // exchange 2: [p][jl*17 + q]       (row stride 17: conflict-free)
constexpr int kWX2I = 64 * 18;                 // inverse exchange 2: [p][jl*18 + g*9 + k]
constexpr int kWBuf0 = 2 * kWX2I;              // complex words of buffer 0 (exchange 1 forward, exchange 2 inverse)
constexpr int kWBuf1 = 2 * kWX2;               // buffer 1 (exchange 2 forward, exchange 1 inverse)
static_assert(kWBuf0 >= 2 * kWX1 && kWBuf1 >= 2 * kWX1, "both buffers hold a full exchange-1 matrix");

// per-thread constants of the transform (device: loaded once into registers; 18 complex words per thread)
struct WideConsts {
    cplx tw2[4];   // stage 2 twiddles of thread (j, q): rA, rB, rC, rC * exp(i pi/4) for the modulus angle of R_j
    cplx tw3[4];   // stage 3 twiddles of thread (jl, g)
    cplx rg;       // +/- sqrt(rho_jl): the half level of thread (jl, g)
    cplx kap[8];   // inverse half level of thread (j, q): 1 for q < 8, conj(sqrt(rho_jl)) for q >= 8, l = 0..7
    double sig;    // +1 for q < 8, -1 for q >= 8
    double pad_;
};
static_assert(sizeof(WideConsts) == 18 * 16, "WideConsts is 18 complex words");

// angle offsets of the eight children of a radix-8 stage, in units of pi/4:  theta_out[o] = theta/8 + kWDelta[o] pi/4
FHE_HD int wide_delta(int o) {
    const int d[8] = {0, 4, 2, 6, 1, 5, 3, 7};
    return d[o];
}

// forward butterfly (a, b) -> (a + w b, a - w b): 6 FMAs (a - w b = 2a - (a + w b))
FHE_HD void w_bfly(double& ar, double& ai, double& br, double& bi, double wr, double wi) {
    const double x = fma(wr, br, ar);
    const double pr = fma(-wi, bi, x);
    const double y = fma(wr, bi, ai);
    const double pi = fma(wi, br, y);
    br = fma(2.0, ar, -pr);
    bi = fma(2.0, ai, -pi);
    ar = pr;
    ai = pi;
}
// inverse (un-normalised): (p, q) -> (p + q, conj(w) (p - q)),  w = the forward twiddle
FHE_HD void w_ibfly(double& pr, double& pi, double& qr, double& qi, double wr, double wi) {
    const double dr = pr - qr, di = pi - qi;
    pr = pr + qr;
    pi = pi + qi;
    qr = fma(wr, dr, wi * di);
    qi = fma(wr, di, -(wi * dr));
}

// radix-8 stage: three levels on v[0..7]; tw = {rA, rB, rC, rC2}; i*w = (-w.y, w.x)
FHE_HD void w_eval8_fwd(double (&re)[8], double (&im)[8], const cplx (&tw)[4]) {
#pragma unroll
    for (int k = 0; k < 4; k++) w_bfly(re[k], im[k], re[k + 4], im[k + 4], tw[0].x, tw[0].y);
#pragma unroll
    for (int k = 0; k < 2; k++) {
        w_bfly(re[k], im[k], re[k + 2], im[k + 2], tw[1].x, tw[1].y);
        w_bfly(re[4 + k], im[4 + k], re[6 + k], im[6 + k], -tw[1].y, tw[1].x);
    }
    w_bfly(re[0], im[0], re[1], im[1], tw[2].x, tw[2].y);
    w_bfly(re[2], im[2], re[3], im[3], -tw[2].y, tw[2].x);
    w_bfly(re[4], im[4], re[5], im[5], tw[3].x, tw[3].y);
    w_bfly(re[6], im[6], re[7], im[7], -tw[3].y, tw[3].x);
}
FHE_HD void w_eval8_inv(double (&re)[8], double (&im)[8], const cplx (&tw)[4]) {
    w_ibfly(re[0], im[0], re[1], im[1], tw[2].x, tw[2].y);
    w_ibfly(re[2], im[2], re[3], im[3], -tw[2].y, tw[2].x);
    w_ibfly(re[4], im[4], re[5], im[5], tw[3].x, tw[3].y);
    w_ibfly(re[6], im[6], re[7], im[7], -tw[3].y, tw[3].x);
#pragma unroll
    for (int k = 0; k < 2; k++) {
        w_ibfly(re[k], im[k], re[k + 2], im[k + 2], tw[1].x, tw[1].y);
        w_ibfly(re[4 + k], im[4 + k], re[6 + k], im[6 + k], -tw[1].y, tw[1].x);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) w_ibfly(re[k], im[k], re[k + 4], im[k + 4], tw[0].x, tw[0].y);
}

// stage-1 twiddles: modulus angle pi/2 for every thread -> literals (constant bank on the device)
FHE_HD void wide_tw1(cplx (&tw)[4]) {
    tw[0] = cplx{0.70710678118654752440, 0.70710678118654752440};     // exp(i pi/4)
    tw[1] = cplx{0.92387953251128675613, 0.38268343236508977173};     // exp(i pi/8)
    tw[2] = cplx{0.98078528040323044913, 0.19509032201612826785};     // exp(i pi/16)
    tw[3] = cplx{0.55557023301960222474, 0.83146961230254523708};     // exp(i 5 pi/16)
}

// Forward transform of the 8 folded points (re + i im) this thread holds for each of P polynomials
// (stage-1 layout: register k = coefficient 128 k + t) into the spectrum layout (point (t, u) in register u).
template <int P, class Ctx>
FHE_HD void wide_forward(Ctx& c, double (&re)[P][8], double (&im)[P][8], const WideConsts& K) {
    const int t = c.tid();
    cplx tw1[4];
    wide_tw1(tw1);
#pragma unroll
    for (int p = 0; p < P; p++) w_eval8_fwd(re[p], im[p], tw1);
    cplx* b0 = c.buf0();
#pragma unroll
    for (int p = 0; p < P; p++)
#pragma unroll
        for (int j = 0; j < 8; j++) b0[p * kWX1 + j * 128 + t] = cplx{re[p][j], im[p][j]};
    c.sync();
    {
        const int base = (t >> 4) * 128 + (t & 15);
#pragma unroll
        for (int p = 0; p < P; p++)
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const cplx v = b0[p * kWX1 + base + 16 * m];
                re[p][m] = v.x;
                im[p][m] = v.y;
            }
    }
#pragma unroll
    for (int p = 0; p < P; p++) w_eval8_fwd(re[p], im[p], K.tw2);
    cplx* b1 = c.buf1();
    c.pre_write_sync();   // single-buffer contexts: every thread has read exchange 1 before anyone overwrites it
    {
        const int base = (t >> 4) * 8 * 17 + (t & 15);
#pragma unroll
        for (int p = 0; p < P; p++)
#pragma unroll
            for (int l = 0; l < 8; l++) b1[p * kWX2 + base + l * 17] = cplx{re[p][l], im[p][l]};
    }
    c.sync();
    {
        const int base = (t >> 1) * 17;
#pragma unroll
        for (int p = 0; p < P; p++)
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const cplx lo = b1[p * kWX2 + base + k];
                const cplx hi = b1[p * kWX2 + base + k + 8];
                re[p][k] = fma(K.rg.x, hi.x, fma(-K.rg.y, hi.y, lo.x));
                im[p][k] = fma(K.rg.x, hi.y, fma(K.rg.y, hi.x, lo.y));
            }
    }
#pragma unroll
    for (int p = 0; p < P; p++) w_eval8_fwd(re[p], im[p], K.tw3);
}

// Inverse: spectrum layout -> stage-1 layout, un-normalised (x 1024).  Uses buffer 0 then buffer 1, so a forward
// that follows (next step) may start writing buffer 0 only after one more barrier (the accumulator barrier).
template <int P, class Ctx>
FHE_HD void wide_inverse(Ctx& c, double (&re)[P][8], double (&im)[P][8], const WideConsts& K) {
    const int t = c.tid();
#pragma unroll
    for (int p = 0; p < P; p++) w_eval8_inv(re[p], im[p], K.tw3);
    cplx* b0 = c.buf0();
    c.pre_write_sync();
    {
        const int base = (t >> 1) * 18 + (t & 1) * 9;
#pragma unroll
        for (int p = 0; p < P; p++)
#pragma unroll
            for (int k = 0; k < 8; k++) b0[p * kWX2I + base + k] = cplx{re[p][k], im[p][k]};
    }
    c.sync();
    {
        const int base = (t >> 4) * 8 * 18 + (t & 7);
#pragma unroll
        for (int p = 0; p < P; p++)
#pragma unroll
            for (int l = 0; l < 8; l++) {
                const cplx lo = b0[p * kWX2I + base + l * 18];
                const cplx hi = b0[p * kWX2I + base + l * 18 + 9];
                const double dr = fma(K.sig, hi.x, lo.x), di = fma(K.sig, hi.y, lo.y);
                re[p][l] = fma(K.kap[l].x, dr, -(K.kap[l].y * di));
                im[p][l] = fma(K.kap[l].x, di, K.kap[l].y * dr);
            }
    }
#pragma unroll
    for (int p = 0; p < P; p++) w_eval8_inv(re[p], im[p], K.tw2);
    cplx* b1 = c.buf1();
    c.pre_write_sync();
    {
        const int base = (t >> 4) * 128 + (t & 15);
#pragma unroll
        for (int p = 0; p < P; p++)
#pragma unroll
            for (int m = 0; m < 8; m++) b1[p * kWX1 + base + 16 * m] = cplx{re[p][m], im[p][m]};
    }
    c.sync();
#pragma unroll
    for (int p = 0; p < P; p++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const cplx v = b1[p * kWX1 + j * 128 + t];
            re[p][j] = v.x;
            im[p][j] = v.y;
        }
    cplx tw1[4];
    wide_tw1(tw1);
#pragma unroll
    for (int p = 0; p < P; p++) w_eval8_inv(re[p], im[p], tw1);
}

FHE_HD int wide_key_index(int row, int col, int u, int t) { return ((row * 2 + col) * 8 + u) * kWT + t; }

// One CMUX step:  ACC += GGSW (x) (X^e ACC - ACC), both polynomials in this thread.
//   a[p][k] = ACC_p[128 k + t], a[p][8 + k] = ACC_p[128 k + t + 1024]  (registers, in/out; the same words are in
//   shared memory for the rotated reads of the other threads).
template <class Ctx>
FHE_HD void wide_cmux_step(Ctx& c, acc_t (&a)[2][16], int e, int step, const WideConsts& K) {
    const int t = c.tid();
    double re[2][8], im[2][8];
    // byte offset of coefficient (t - e) mod 2N in the negacyclic extension; register k adds 128 words, the folded
    // partner 1024 words; bit 13 is the sign (acc_ld_rot)
    const uint32_t x0 = (uint32_t)((t - e) & (2 * kN - 1)) << 2;
#pragma unroll
    for (int p = 0; p < 2; p++)
#pragma unroll
        for (int k = 0; k < 8; k++) {
            re[p][k] = digit23(c.acc_ld_rot(p, x0 + 512u * k) - a[p][k]);
            im[p][k] = digit23(c.acc_ld_rot(p, x0 + 512u * k + 4096u) - a[p][8 + k]);
        }
    wide_forward<2>(c, re, im, K);
    const cplx* key = c.key_wait(step);   // this step's tile (device: shared memory, after the mbarrier wait)
#pragma unroll
    for (int u = 0; u < 8; u++) {
        const cplx g00 = key[wide_key_index(0, 0, u, t)], g01 = key[wide_key_index(0, 1, u, t)];
        const cplx g10 = key[wide_key_index(1, 0, u, t)], g11 = key[wide_key_index(1, 1, u, t)];
        const double d0r = re[0][u], d0i = im[0][u], d1r = re[1][u], d1i = im[1][u];
        re[0][u] = fma(d1r, g10.x, fma(-d1i, g10.y, fma(d0r, g00.x, -(d0i * g00.y))));
        im[0][u] = fma(d1r, g10.y, fma(d1i, g10.x, fma(d0r, g00.y, d0i * g00.x)));
        re[1][u] = fma(d1r, g11.x, fma(-d1i, g11.y, fma(d0r, g01.x, -(d0i * g01.y))));
        im[1][u] = fma(d1r, g11.y, fma(d1i, g11.x, fma(d0r, g01.y, d0i * g01.x)));
    }
    c.key_release(step);   // shared-tile contexts: this thread is done with the step's key tile
    wide_inverse<2>(c, re, im, K);
#pragma unroll
    for (int p = 0; p < 2; p++) {
        acc_t* acc = c.acc(p);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            a[p][k] += torus32_conv(re[p][k], k);
            a[p][8 + k] += torus32_conv(im[p][k], k);
            acc[128 * k + t] = a[p][k];
            acc[128 * k + t + kM] = a[p][8 + k];
        }
    }
    c.sync();
}

// ---------------------------------------------------------------------------------------------------------------
// The same CMUX step as a two-stream software pipeline (contexts with one barrier PER POLYNOMIAL: c.sync_poly(p)).
// The plain step above runs both polynomials through every exchange behind ONE barrier, so all four warps of a PBS are
// always in the same phase and the shared-memory pipe and the FP64 pipe take turns (ncu: 57 % + 47 % of the step).
// Here polynomial 1 runs half a phase behind polynomial 0: every batch of shared-memory stores is followed by an FP64
// block of the OTHER polynomial before its barrier, and every batch of loads by one before its consumer, so the two
// pipes overlap.  Same arithmetic, same order of operations per polynomial: bit-identical results.
FHE_HD void w_eval8_fwd_ab(double (&re)[8], double (&im)[8], const cplx (&tw)[4]) {
#pragma unroll
    for (int k = 0; k < 4; k++) w_bfly(re[k], im[k], re[k + 4], im[k + 4], tw[0].x, tw[0].y);
#pragma unroll
    for (int k = 0; k < 2; k++) {
        w_bfly(re[k], im[k], re[k + 2], im[k + 2], tw[1].x, tw[1].y);
        w_bfly(re[4 + k], im[4 + k], re[6 + k], im[6 + k], -tw[1].y, tw[1].x);
    }
}
FHE_HD void w_eval8_fwd_c(double (&re)[8], double (&im)[8], const cplx (&tw)[4]) {
    w_bfly(re[0], im[0], re[1], im[1], tw[2].x, tw[2].y);
    w_bfly(re[2], im[2], re[3], im[3], -tw[2].y, tw[2].x);
    w_bfly(re[4], im[4], re[5], im[5], tw[3].x, tw[3].y);
    w_bfly(re[6], im[6], re[7], im[7], -tw[3].y, tw[3].x);
}
FHE_HD void w_eval8_inv_c(double (&re)[8], double (&im)[8], const cplx (&tw)[4]) {
    w_ibfly(re[0], im[0], re[1], im[1], tw[2].x, tw[2].y);
    w_ibfly(re[2], im[2], re[3], im[3], -tw[2].y, tw[2].x);
    w_ibfly(re[4], im[4], re[5], im[5], tw[3].x, tw[3].y);
    w_ibfly(re[6], im[6], re[7], im[7], -tw[3].y, tw[3].x);
}
FHE_HD void w_eval8_inv_ba(double (&re)[8], double (&im)[8], const cplx (&tw)[4]) {
#pragma unroll
    for (int k = 0; k < 2; k++) {
        w_ibfly(re[k], im[k], re[k + 2], im[k + 2], tw[1].x, tw[1].y);
        w_ibfly(re[4 + k], im[4 + k], re[6 + k], im[6 + k], -tw[1].y, tw[1].x);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) w_ibfly(re[k], im[k], re[k + 4], im[k + 4], tw[0].x, tw[0].y);
}

template <class Ctx>
FHE_HD void wide_cmux_step_pipe(Ctx& c, acc_t (&a)[2][16], int e, int step, const WideConsts& K) {
    const int t = c.tid();
    double re[2][8], im[2][8];
    cplx tw1[4];
    wide_tw1(tw1);
    cplx* b0 = c.buf0();
    cplx* b1 = c.buf1();
    const uint32_t x0 = (uint32_t)((t - e) & (2 * kN - 1)) << 2;
    const int i1 = (t >> 4) * 128 + (t & 15);            // exchange 1, reader side (and inverse exchange 1, writer side)
    const int i2w = (t >> 4) * 8 * 17 + (t & 15);        // exchange 2, writer
    const int i2r = (t >> 1) * 17;                       // exchange 2, reader
    const int i3w = (t >> 1) * 18 + (t & 1) * 9;         // inverse exchange 2, writer
    const int i3r = (t >> 4) * 8 * 18 + (t & 7);         // inverse exchange 2, reader

#define W_GATHER(p)                                                                              \
    _Pragma("unroll") for (int k = 0; k < 8; k++) {                                              \
        re[p][k] = digit23(c.acc_ld_rot(p, x0 + 512u * k) - a[p][k]);                       \
        im[p][k] = digit23(c.acc_ld_rot(p, x0 + 512u * k + 4096u) - a[p][8 + k]);           \
    }
#define W_S1(p) _Pragma("unroll") for (int j = 0; j < 8; j++) b0[p * kWX1 + j * 128 + t] = cplx{re[p][j], im[p][j]};
#define W_L1(p) _Pragma("unroll") for (int m = 0; m < 8; m++) { const cplx v = b0[p * kWX1 + i1 + 16 * m]; re[p][m] = v.x; im[p][m] = v.y; }
#define W_S2(p) _Pragma("unroll") for (int l = 0; l < 8; l++) b1[p * kWX2 + i2w + l * 17] = cplx{re[p][l], im[p][l]};
#define W_L2(p)                                                                                  \
    _Pragma("unroll") for (int k = 0; k < 8; k++) {                                              \
        const cplx lo = b1[p * kWX2 + i2r + k];                                                  \
        const cplx hi = b1[p * kWX2 + i2r + k + 8];                                              \
        re[p][k] = fma(K.rg.x, hi.x, fma(-K.rg.y, hi.y, lo.x));                                  \
        im[p][k] = fma(K.rg.x, hi.y, fma(K.rg.y, hi.x, lo.y));                                   \
    }
#define W_S3(p) _Pragma("unroll") for (int k = 0; k < 8; k++) b0[p * kWX2I + i3w + k] = cplx{re[p][k], im[p][k]};
#define W_L3(p)                                                                                  \
    _Pragma("unroll") for (int l = 0; l < 8; l++) {                                              \
        const cplx lo = b0[p * kWX2I + i3r + l * 18];                                            \
        const cplx hi = b0[p * kWX2I + i3r + l * 18 + 9];                                        \
        const double dr = fma(K.sig, hi.x, lo.x), di = fma(K.sig, hi.y, lo.y);                   \
        re[p][l] = fma(K.kap[l].x, dr, -(K.kap[l].y * di));                                      \
        im[p][l] = fma(K.kap[l].x, di, K.kap[l].y * dr);                                         \
    }
#define W_S4(p) _Pragma("unroll") for (int m = 0; m < 8; m++) b1[p * kWX1 + i1 + 16 * m] = cplx{re[p][m], im[p][m]};
#define W_L4(p) _Pragma("unroll") for (int j = 0; j < 8; j++) { const cplx v = b1[p * kWX1 + j * 128 + t]; re[p][j] = v.x; im[p][j] = v.y; }
#define W_ACC(p)                                                                                 \
    {                                                                                            \
        acc_t* acc = c.acc(p);                                                                   \
        _Pragma("unroll") for (int k = 0; k < 8; k++) {                                          \
            a[p][k] += torus32_conv(re[p][k], k);                                                \
            a[p][8 + k] += torus32_conv(im[p][k], k);                                            \
            acc[128 * k + t] = a[p][k];                                                          \
            acc[128 * k + t + kM] = a[p][8 + k];                                                 \
        }                                                                                        \
    }

    // ---- forward
    c.sync_poly(0);                      // accumulator 0 of the previous step is complete
    W_GATHER(0)
    w_eval8_fwd(re[0], im[0], tw1);
    W_S1(0)
    c.sync_poly(1);
    W_GATHER(1)
    w_eval8_fwd_ab(re[1], im[1], tw1);
    c.sync_poly(0); W_L1(0)
    w_eval8_fwd_c(re[1], im[1], tw1);
    W_S1(1)
    w_eval8_fwd_ab(re[0], im[0], K.tw2);
    c.sync_poly(1); W_L1(1)
    w_eval8_fwd_c(re[0], im[0], K.tw2);
    W_S2(0)
    w_eval8_fwd_ab(re[1], im[1], K.tw2);
    c.sync_poly(0); W_L2(0)
    w_eval8_fwd_c(re[1], im[1], K.tw2);
    W_S2(1)
    w_eval8_fwd_ab(re[0], im[0], K.tw3);
    c.sync_poly(1); W_L2(1)
    w_eval8_fwd_c(re[0], im[0], K.tw3);
    w_eval8_fwd(re[1], im[1], K.tw3);
    // ---- GGSW product (both polynomials of every spectrum point are in this thread)
    const cplx* key = c.key_wait(step);
#pragma unroll
    for (int u = 0; u < 8; u++) {
        const cplx g00 = key[wide_key_index(0, 0, u, t)], g01 = key[wide_key_index(0, 1, u, t)];
        const cplx g10 = key[wide_key_index(1, 0, u, t)], g11 = key[wide_key_index(1, 1, u, t)];
        const double d0r = re[0][u], d0i = im[0][u], d1r = re[1][u], d1i = im[1][u];
        re[0][u] = fma(d1r, g10.x, fma(-d1i, g10.y, fma(d0r, g00.x, -(d0i * g00.y))));
        im[0][u] = fma(d1r, g10.y, fma(d1i, g10.x, fma(d0r, g00.y, d0i * g00.x)));
        re[1][u] = fma(d1r, g11.x, fma(-d1i, g11.y, fma(d0r, g01.x, -(d0i * g01.y))));
        im[1][u] = fma(d1r, g11.y, fma(d1i, g11.x, fma(d0r, g01.y, d0i * g01.x)));
    }
    c.key_release(step);
    // ---- inverse
    w_eval8_inv(re[0], im[0], K.tw3);
    W_S3(0)
    w_eval8_inv_c(re[1], im[1], K.tw3);
    c.sync_poly(0); W_L3(0)
    w_eval8_inv_ba(re[1], im[1], K.tw3);
    W_S3(1)
    w_eval8_inv_c(re[0], im[0], K.tw2);
    c.sync_poly(1); W_L3(1)
    w_eval8_inv_ba(re[0], im[0], K.tw2);
    W_S4(0)
    w_eval8_inv_c(re[1], im[1], K.tw2);
    c.sync_poly(0); W_L4(0)
    w_eval8_inv_ba(re[1], im[1], K.tw2);
    W_S4(1)
    w_eval8_inv_c(re[0], im[0], tw1);
    c.sync_poly(1); W_L4(1)
    w_eval8_inv_ba(re[0], im[0], tw1);
    W_ACC(0)
    w_eval8_inv(re[1], im[1], tw1);
    W_ACC(1)
#undef W_GATHER
#undef W_S1
#undef W_L1
#undef W_S2
#undef W_L2
#undef W_S3
#undef W_L3
#undef W_S4
#undef W_L4
#undef W_ACC
}

// Forward transform of one standard-domain GGSW polynomial into tile position (row, col) (key conversion, once)
template <class Ctx>
FHE_HD void wide_bsk_poly_forward(Ctx& c, const u64* poly, cplx* out_step, int row, int col, const WideConsts& K) {
    const int t = c.tid();
    const double sc = 1.0 / (4294967296.0 * (double)kM);   // 2^-64 (u64 -> turns) * 2^32 (turns -> acc_t ulps) / M
    double re[1][8], im[1][8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        re[0][k] = (double)(i64)poly[128 * k + t] * sc;
        im[0][k] = (double)(i64)poly[128 * k + t + kM] * sc;
    }
    wide_forward<1>(c, re, im, K);
#pragma unroll
    for (int u = 0; u < 8; u++) out_step[wide_key_index(row, col, u, t)] = cplx{re[0][u], im[0][u]};
    c.sync();   // the exchange buffers are reused by the next polynomial
}

// Whole blind rotation of one PBS by 128 threads, mod-switch in front and sample extract behind (as br_thread_main)
template <class Ctx>
FHE_HD void wide_thread_main(Ctx& c, const BrJobView& job, const WideConsts& K) {
    const int t = c.tid();
    const int n = job.n;
    uint16_t* at = c.atilde();
    for (int idx = t; idx <= n; idx += kWT) at[idx] = (uint16_t)modswitch_2N(job.ks[idx]);
    c.key_prefetch(0);
    c.sync();
    acc_t a[2][16];
    {
        const int e0 = (2 * kN - (int)at[n]) & (2 * kN - 1);
#pragma unroll
        for (int p = 0; p < 2; p++) {
            acc_t* acc = c.acc(p);
#pragma unroll
            for (int m = 0; m < 16; m++) {
                const int j = 128 * (m & 7) + t + (m >> 3) * kM;
                u64 v;
                if (job.init_acc) v = job.init_acc[p * kN + j];
                else v = (p == 1) ? rot_coef(job.lut, j, e0) : (u64)0;
                a[p][m] = acc_from_u64(v);
                acc[j] = a[p][m];
            }
        }
    }
    c.sync();
    for (int i = 0; i < n; i++) {
        if (i + 1 < n) c.key_prefetch(i + 1);   // ring contexts: next step's tile
        c.key_prefetch_current(i);              // single-tile contexts: this step's tile (everyone is done with the last)
        // e == 0 is NOT skipped here (the key pipeline stays in step): every digit is exactly 0, so is the product
        if constexpr (Ctx::kPipelined) wide_cmux_step_pipe(c, a, at[i], i, K);
        else wide_cmux_step(c, a, at[i], i, K);
    }
    if constexpr (Ctx::kPipelined) c.sync();   // the pipelined step leaves its last accumulator stores un-fenced
    if (job.out_acc) {
#pragma unroll
        for (int p = 0; p < 2; p++)
#pragma unroll
            for (int m = 0; m < 16; m++) job.out_acc[p * kN + 128 * (m & 7) + t + (m >> 3) * kM] = acc_to_u64(a[p][m]);
    }
    if (job.out_lwe) {
        const acc_t* acc0 = c.acc(0);
#pragma unroll
        for (int m = 0; m < 16; m++) {
            const int j = 128 * m + t;
            const u64 w = acc_to_u64((j == 0) ? acc0[0] : (acc_t)0 - acc0[kN - j]);
            job.out_lwe[j] = w;
            for (int r = 0; r < job.n_peers; r++) job.out_lwe_peer[r][j] = w;
        }
        if (t == 0) {
            const u64 b = acc_to_u64(c.acc(1)[0]) + job.post;
            job.out_lwe[kN] = b;
            for (int r = 0; r < job.n_peers; r++) job.out_lwe_peer[r][kN] = b;
        }
    }
}

// Per-thread constant table, shared by the engine and the host emulation.  Angles are kept as exact multiples of
// pi / 2^14 so that every twiddle is one cosl/sinl of an exactly representable argument.
inline void make_wide_consts(WideConsts* tab /* [kWT] */) {
    const long double kPi = 3.14159265358979323846264338327950288L;
    auto cis = [&](long double units /* of pi/16384 */) {
        const long double ang = kPi * units / 16384.0L;
        return cplx{(double)cosl(ang), (double)sinl(ang)};
    };
    auto fill_tw = [&](cplx (&tw)[4], long double th) {   // modulus angle th (units)
        tw[0] = cis(th / 2);
        tw[1] = cis(th / 4);
        tw[2] = cis(th / 8);
        tw[3] = cis(th / 8 + 4096.0L);   // + pi/4
    };
    // stage 1: theta = pi/2 = 8192 units -> residue j has angle theta/8 + delta_j pi/4
    auto th1 = [&](int j) { return 8192.0L / 8 + 4096.0L * wide_delta(j); };
    auto th2 = [&](int j, int l) { return th1(j) / 8 + 4096.0L * wide_delta(l); };
    for (int t = 0; t < kWT; t++) {
        WideConsts& K = tab[t];
        {   // stage-2 role: (j, q)
            const int j = t >> 4, q = t & 15;
            fill_tw(K.tw2, th1(j));
            K.sig = q < 8 ? 1.0 : -1.0;
            for (int l = 0; l < 8; l++) {
                if (q < 8) K.kap[l] = cplx{1.0, 0.0};
                else {
                    const cplx r = cis(th2(j, l) / 2);
                    K.kap[l] = cplx{r.x, -r.y};
                }
            }
        }
        {   // stage-3 role: (jl, g)
            const int jl = t >> 1, g = t & 1;
            const long double th = th2(jl >> 3, jl & 7);
            const cplx r = cis(th / 2);
            K.rg = g ? cplx{-r.x, -r.y} : r;
            fill_tw(K.tw3, th / 2 + (g ? 16384.0L : 0.0L));
        }
        K.pad_ = 0.0;
    }
}

}  // namespace fhestr
