/*
 * oracle/tfhe_oracle.c -- CPU oracle for the batched-PBS hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (fhestring_b200/) never links or calls it.
 *
 * What it restates: the arithmetic behind the 13 tfhe-rs call sites of the reference's
 * per-character primitives (/root/reference/src/ciphertext/fheasciichar.rs:23,28,32,36-37,41-42,
 * 46-47,51-52,56-57,61-62,70,79,84,89,99,102) i.e. one shortint PBS =
 * keyswitch -> modulus switch -> blind rotation -> sample extract, plus LUT (accumulator)
 * generation and the client-side encrypt/decrypt of /root/reference/src/client_key.rs:45-106.
 *
 * That arithmetic lives in the un-vendored crate tfhe 0.5.2 (/root/reference/Cargo.lock:416-417,
 * concrete-fft 0.4.0 at Cargo.lock:168-169), which is NOT present in /root/reference and cannot
 * be built here (no Rust toolchain, no network).  The algorithms below follow the published TFHE
 * construction with the tfhe-rs 0.5 conventions written down in SURVEY.md Appendix A
 * (A.2 encoding, A.4 signed decomposer, A.5 keyswitch, A.6 modulus switch, A.7 blind rotation,
 * A.8 sample extract, LUT generation in SURVEY.md section 2.5).
 *
 * PARITY STATUS: ciphertext-level parity with tfhe-rs is UNPINNED (the reference holds no golden
 * ciphertext, KAT or fixture: SURVEY.md section 8c).  What IS pinned: decrypted plaintext results,
 * against the literal strings of the reference's 43 unit tests (src/main.rs:138-1153) through
 * oracle/string_oracle.py.  This file is the exact-integer ground truth for the engine's
 * keyswitch / mod-switch / sample-extract / LUT kernels (bit-exact) and the centre of the
 * tolerance band for the FFT blind rotation.
 *
 * Two blind-rotation variants are provided:
 *   - exact:  negacyclic products over Z/2^64 by schoolbook multiplication (no rounding at all);
 *   - fft:    the f64 fold+twist negacyclic FFT route tfhe-rs itself uses (A.7).  This one is the
 *             "port" CPU baseline timed by bench.py and the calibration for the GPU tolerance.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t u64;
typedef int64_t i64;
typedef uint32_t u32;
typedef uint8_t u8;

typedef struct {
    int32_t n;            /* small LWE dimension (742) */
    int32_t N;            /* polynomial size (2048), power of two */
    int32_t k;            /* GLWE dimension; only 1 is supported */
    int32_t pbs_base_log; /* 23 */
    int32_t pbs_level;    /* 1 */
    int32_t ks_base_log;  /* 3 */
    int32_t ks_level;     /* 5 */
    int32_t delta_log;    /* 59: 2 msg + 2 carry bits + 1 padding bit */
    double lwe_std;       /* 7.069849454709433e-6 */
    double glwe_std;      /* 2.9403601535432533e-16 */
} orc_params;

#define CLONES __attribute__((target_clones("default", "avx2", "arch=x86-64-v4")))

/* ------------------------------------------------------------------ RNG (seeded, deterministic) */
typedef struct { u64 s[4]; int have_spare; double spare; } rng_t;

static u64 splitmix64(u64 *x) {
    u64 z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static void rng_seed(rng_t *r, u64 seed, u64 stream) {
    u64 x = seed ^ (stream * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull);
    for (int i = 0; i < 4; i++) r->s[i] = splitmix64(&x);
    r->have_spare = 0;
}
static inline u64 rotl64(u64 x, int k) { return (x << k) | (x >> (64 - k)); }
static u64 rng_u64(rng_t *r) { /* xoshiro256** */
    u64 *s = r->s;
    u64 result = rotl64(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl64(s[3], 45);
    return result;
}
static double rng_unit(rng_t *r) { return ((rng_u64(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static double rng_gauss(rng_t *r) { /* Box-Muller */
    if (r->have_spare) { r->have_spare = 0; return r->spare; }
    double u = rng_unit(r), v = rng_unit(r);
    double m = sqrt(-2.0 * log(u));
    r->spare = m * sin(6.283185307179586476925 * v);
    r->have_spare = 1;
    return m * cos(6.283185307179586476925 * v);
}
/* torus noise: round(std * 2^64 * g) as a wrapping u64 */
static u64 rng_torus_noise(rng_t *r, double std) {
    double e = rng_gauss(r) * std * 18446744073709551616.0;
    return (u64)(i64)llrint(e);
}

/* ------------------------------------------------------------------ A.4 signed decomposer */
/* digits[lvl-1] for lvl = 1..level, each in [-B/2, B/2], sum digits[lvl-1] * 2^(64-base_log*lvl)
 * == closest representable value of x (mod 2^64). */
static inline void decompose(u64 x, int base_log, int level, i64 *digits) {
    const int rep = base_log * level;       /* representable bits */
    const int r = 64 - rep;
    u64 state = ((x >> (r - 1)) + 1) >> 1;  /* closest representable, kept on `rep` bits */
    if (rep < 64) state &= (((u64)1) << rep) - 1;
    const u64 mask = (((u64)1) << base_log) - 1;
    for (int lvl = level; lvl >= 1; lvl--) {
        u64 d = state & mask;
        state >>= base_log;
        u64 carry = (((d - 1) | state) & d) >> (base_log - 1);
        state += carry;
        digits[lvl - 1] = (i64)d - (i64)(carry << base_log);
    }
}

void orc_decompose(u64 x, int base_log, int level, i64 *digits) { decompose(x, base_log, level, digits); }

/* ------------------------------------------------------------------ A.6 modulus switch to 2N */
static inline u32 modswitch(u64 x, int log2_2N) {
    u64 t = x >> (64 - log2_2N - 1);
    t += t & 1;
    t >>= 1;
    return (u32)t; /* in [0, 2N] ; 2N == 0 as a rotation */
}
u32 orc_modswitch(u64 x, int log2_2N) { return modswitch(x, log2_2N); }

static int ilog2(int x) { int l = 0; while ((1 << l) < x) l++; return l; }

/* ------------------------------------------------------------------ keys */
/* negacyclic out += a * s for a binary polynomial s */
CLONES static void negacyclic_mul_binary_acc(u64 *out, const u64 *a, const u8 *s, int N) {
    for (int j = 0; j < N; j++) {
        if (!s[j]) continue;
        for (int u = 0; u < N - j; u++) out[u + j] += a[u];
        for (int u = N - j; u < N; u++) out[u + j - N] -= a[u];
    }
}

/* Layouts (row-major):
 *   s_lwe[n], s_glwe[N]                        secret keys, one byte per bit
 *   bsk[n][pbs_level][2 rows][2 cols][N]       standard-domain GGSW per small-key bit; row (lvl,r) is a
 *                                              GLWE encryption of 0 with s_lwe[i] << (64 - base_log*lvl)
 *                                              added to coefficient 0 of polynomial r (A.7)
 *   ksk[N][ks_level][n+1]                      level 1 first; KSK[i][lvl] encrypts
 *                                              s_glwe[i] << (64 - ks_base_log*lvl) under s_lwe (A.5)
 */
void orc_keygen(const orc_params *p, u64 seed, u8 *s_lwe, u8 *s_glwe, u64 *bsk, u64 *ksk) {
    const int n = p->n, N = p->N;
    rng_t r;
    rng_seed(&r, seed, 1);
    for (int i = 0; i < n; i++) s_lwe[i] = (u8)(rng_u64(&r) >> 63);
    rng_seed(&r, seed, 2);
    for (int i = 0; i < N; i++) s_glwe[i] = (u8)(rng_u64(&r) >> 63);

    if (bsk) {
#pragma omp parallel for schedule(dynamic, 4)
        for (int i = 0; i < n; i++) {
            rng_t ri;
            rng_seed(&ri, seed, 0x1000000ull + (u64)i);
            for (int lvl = 1; lvl <= p->pbs_level; lvl++)
                for (int row = 0; row < 2; row++) {
                    u64 *A = bsk + ((((size_t)i * p->pbs_level + (lvl - 1)) * 2 + row) * 2 + 0) * N;
                    u64 *B = A + N;
                    for (int j = 0; j < N; j++) A[j] = rng_u64(&ri);
                    for (int j = 0; j < N; j++) B[j] = rng_torus_noise(&ri, p->glwe_std);
                    negacyclic_mul_binary_acc(B, A, s_glwe, N);
                    u64 g = ((u64)s_lwe[i]) << (64 - p->pbs_base_log * lvl);
                    if (row == 0) A[0] += g; else B[0] += g;
                }
        }
    }
    if (ksk) {
#pragma omp parallel for schedule(static)
        for (int i = 0; i < N; i++) {
            rng_t ri;
            rng_seed(&ri, seed, 0x2000000ull + (u64)i);
            for (int lvl = 1; lvl <= p->ks_level; lvl++) {
                u64 *ct = ksk + ((size_t)i * p->ks_level + (lvl - 1)) * (n + 1);
                u64 body = rng_torus_noise(&ri, p->lwe_std);
                for (int c = 0; c < n; c++) {
                    ct[c] = rng_u64(&ri);
                    if (s_lwe[c]) body += ct[c];
                }
                body += ((u64)s_glwe[i]) << (64 - p->ks_base_log * lvl);
                ct[n] = body;
            }
        }
    }
}

/* LWE encryption of a raw torus plaintext under a binary key of dimension dim. out[dim+1]. */
void orc_lwe_encrypt(const u8 *key, int dim, u64 plaintext, double std, u64 seed, u64 stream, u64 *out) {
    rng_t r;
    rng_seed(&r, seed, 0x3000000ull + stream);
    u64 body = plaintext + rng_torus_noise(&r, std);
    for (int i = 0; i < dim; i++) {
        out[i] = rng_u64(&r);
        if (key[i]) body += out[i];
    }
    out[dim] = body;
}
void orc_lwe_trivial(int dim, u64 plaintext, u64 *out) {
    memset(out, 0, sizeof(u64) * dim);
    out[dim] = plaintext;
}
u64 orc_lwe_phase(const u8 *key, int dim, const u64 *ct) {
    u64 ph = ct[dim];
    for (int i = 0; i < dim; i++) if (key[i]) ph -= ct[i];
    return ph;
}
/* A.2 decode: ((phase + delta/2) >> delta_log) mod 2^(64-delta_log-1) -- padding bit dropped */
u32 orc_decode(u64 phase, int delta_log) {
    u64 v = (phase + (((u64)1) << (delta_log - 1))) >> delta_log;
    return (u32)(v & ((((u64)1) << (63 - delta_log)) - 1));
}
void orc_lwe_phase_batch(const u8 *key, int dim, const u64 *cts, int count, u64 *phases) {
#pragma omp parallel for schedule(static)
    for (int b = 0; b < count; b++) phases[b] = orc_lwe_phase(key, dim, cts + (size_t)b * (dim + 1));
}

/* ------------------------------------------------------------------ A.5 keyswitch */
CLONES void orc_keyswitch(const orc_params *p, const u64 *ksk, const u64 *in, u64 *out) {
    const int n = p->n, Nb = p->N * p->k, L = p->ks_level;
    i64 digits[64];
    for (int c = 0; c < n; c++) out[c] = 0;
    out[n] = in[Nb];
    for (int i = 0; i < Nb; i++) {
        decompose(in[i], p->ks_base_log, L, digits);
        for (int lvl = 1; lvl <= L; lvl++) {
            const u64 d = (u64)digits[lvl - 1];
            if (!d) continue;
            const u64 *row = ksk + ((size_t)i * L + (lvl - 1)) * (n + 1);
            for (int c = 0; c <= n; c++) out[c] -= d * row[c];
        }
    }
}
void orc_keyswitch_batch(const orc_params *p, const u64 *ksk, const u64 *in, int count, u64 *out) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < count; b++)
        orc_keyswitch(p, ksk, in + (size_t)b * (p->N * p->k + 1), out + (size_t)b * (p->n + 1));
}

/* ------------------------------------------------------------------ section 2.5 LUT polynomial */
/* table[m] for m in [0, 2^(63-delta_log)) -> body polynomial of the trivial GLWE accumulator */
void orc_lut_poly(int N, const u8 *table, int delta_log, u64 *out) {
    const int entries = 1 << (63 - delta_log);
    const int box = N / entries;
    u64 *tmp = (u64 *)malloc(sizeof(u64) * N);
    for (int i = 0; i < entries; i++)
        for (int j = 0; j < box; j++) tmp[i * box + j] = ((u64)table[i]) << delta_log;
    for (int j = 0; j < box / 2; j++) tmp[j] = (u64)0 - tmp[j];
    for (int j = 0; j < N; j++) out[j] = tmp[(j + box / 2) % N]; /* rotate left by box/2 */
    free(tmp);
}

/* ------------------------------------------------------------------ negacyclic helpers */
/* out = in * X^e, e in [0, 2N) */
static void monomial_mul(u64 *out, const u64 *in, int e, int N) {
    int neg = 0;
    if (e >= N) { e -= N; neg = 1; }
    for (int j = 0; j < N - e; j++) out[j + e] = neg ? (u64)0 - in[j] : in[j];
    for (int j = N - e; j < N; j++) out[j + e - N] = neg ? in[j] : (u64)0 - in[j];
}
/* out += d * g (negacyclic, exact mod 2^64), d small signed digits */
CLONES static void negacyclic_mul_acc(u64 *out, const i64 *d, const u64 *g, int N) {
    for (int s = 0; s < N; s++) {
        const u64 ds = (u64)d[s];
        if (!ds) continue;
        for (int u = 0; u < N - s; u++) out[u + s] += ds * g[u];
        for (int u = N - s; u < N; u++) out[u + s - N] -= ds * g[u];
    }
}

/* exact external product: acc[2][N] += GGSW (rows [level][2][2][N]) (x) glwe[2][N] */
void orc_external_product_exact(const orc_params *p, const u64 *ggsw, const u64 *glwe, u64 *acc) {
    const int N = p->N, L = p->pbs_level;
    i64 *dig = (i64 *)malloc(sizeof(i64) * N * L);
    i64 tmp[64];
    for (int r = 0; r < 2; r++) {
        for (int j = 0; j < N; j++) {
            decompose(glwe[r * N + j], p->pbs_base_log, L, tmp);
            for (int l = 0; l < L; l++) dig[l * N + j] = tmp[l];
        }
        for (int l = 0; l < L; l++)
            for (int c = 0; c < 2; c++)
                negacyclic_mul_acc(acc + c * N, dig + l * N, ggsw + (((size_t)l * 2 + r) * 2 + c) * N, N);
    }
    free(dig);
}

/* ------------------------------------------------------------------ A.7 blind rotation, exact */
/* ks[n+1] (small-key LWE), lut[N] body polynomial; acc[2][N] receives the final accumulator */
void orc_blind_rotate_exact(const orc_params *p, const u64 *bsk, const u64 *ks, const u64 *lut, u64 *acc) {
    const int n = p->n, N = p->N, lg = ilog2(2 * N);
    u64 *rot = (u64 *)malloc(sizeof(u64) * 2 * N);
    u32 bt = modswitch(ks[n], lg) % (2 * N);
    memset(acc, 0, sizeof(u64) * N);
    monomial_mul(acc + N, lut, (2 * N - bt) % (2 * N), N);
    const size_t ggsw_sz = (size_t)p->pbs_level * 4 * N;
    for (int i = 0; i < n; i++) {
        u32 at = modswitch(ks[i], lg) % (2 * N);
        if (at == 0) continue;
        for (int r = 0; r < 2; r++) {
            monomial_mul(rot + r * N, acc + r * N, at, N);
            for (int j = 0; j < N; j++) rot[r * N + j] -= acc[r * N + j];
        }
        orc_external_product_exact(p, bsk + i * ggsw_sz, rot, acc);
    }
    free(rot);
}

/* ------------------------------------------------------------------ A.8 sample extract (coef 0) */
void orc_sample_extract(int N, const u64 *acc, u64 *out) {
    out[0] = acc[0];
    for (int j = 1; j < N; j++) out[j] = (u64)0 - acc[N - j];
    out[N] = acc[N + 0];
}

/* full exact PBS: in[N+1] (big key) -> out[N+1] (big key) */
void orc_pbs_exact(const orc_params *p, const u64 *bsk, const u64 *ksk, const u64 *lut, const u64 *in, u64 *out) {
    u64 *ks = (u64 *)malloc(sizeof(u64) * (p->n + 1));
    u64 *acc = (u64 *)malloc(sizeof(u64) * 2 * p->N);
    orc_keyswitch(p, ksk, in, ks);
    orc_blind_rotate_exact(p, bsk, ks, lut, acc);
    orc_sample_extract(p->N, acc, out);
    free(ks); free(acc);
}
void orc_pbs_exact_batch(const orc_params *p, const u64 *bsk, const u64 *ksk, const u64 *luts,
                         const int32_t *lut_ids, const u64 *in, int count, u64 *out) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < count; b++)
        orc_pbs_exact(p, bsk, ksk, luts + (size_t)lut_ids[b] * p->N, in + (size_t)b * (p->N + 1),
                      out + (size_t)b * (p->N + 1));
}

/* ------------------------------------------------------------------ f64 negacyclic FFT route */
/* X_k = P(y_k), y_k = exp(i*pi*(1-4k)/N), k < N/2, via fold (p_j + i p_{j+N/2}), twist exp(i*pi*j/N),
 * then an N/2-point forward DFT (DIF, output bit-reversed).  The inverse is a DIT taking the
 * bit-reversed order back to natural, so no permutation is ever materialised. */
typedef struct {
    int N, M;
    double *tw_re, *tw_im;       /* exp(-2 pi i j / M), j < M/2 */
    double *twist_re, *twist_im; /* exp(i pi j / N), j < M */
} fft_plan;

fft_plan *orc_fft_plan_new(int N) {
    fft_plan *pl = (fft_plan *)malloc(sizeof(fft_plan));
    const int M = N / 2;
    pl->N = N; pl->M = M;
    pl->tw_re = (double *)malloc(sizeof(double) * M / 2);
    pl->tw_im = (double *)malloc(sizeof(double) * M / 2);
    pl->twist_re = (double *)malloc(sizeof(double) * M);
    pl->twist_im = (double *)malloc(sizeof(double) * M);
    for (int j = 0; j < M / 2; j++) {
        long double a = -2.0L * 3.14159265358979323846264338327950288L * j / M;
        pl->tw_re[j] = (double)cosl(a); pl->tw_im[j] = (double)sinl(a);
    }
    for (int j = 0; j < M; j++) {
        long double a = 3.14159265358979323846264338327950288L * j / N;
        pl->twist_re[j] = (double)cosl(a); pl->twist_im[j] = (double)sinl(a);
    }
    return pl;
}
void orc_fft_plan_free(fft_plan *pl) {
    free(pl->tw_re); free(pl->tw_im); free(pl->twist_re); free(pl->twist_im); free(pl);
}

/* in-place DIF forward, natural in -> bit-reversed out (split re/im arrays) */
CLONES static void fft_dif(const fft_plan *pl, double *re, double *im) {
    const int M = pl->M;
    for (int half = M / 2, stride = 1; half >= 1; half >>= 1, stride <<= 1) {
        for (int base = 0; base < M; base += 2 * half) {
            double *ar = re + base, *ai = im + base, *br = re + base + half, *bi = im + base + half;
            for (int j = 0; j < half; j++) {
                const double wr = pl->tw_re[j * stride], wi = pl->tw_im[j * stride];
                const double xr = ar[j] - br[j], xi = ai[j] - bi[j];
                ar[j] += br[j]; ai[j] += bi[j];
                br[j] = xr * wr - xi * wi; bi[j] = xr * wi + xi * wr;
            }
        }
    }
}
/* in-place DIT inverse (conjugate twiddles), bit-reversed in -> natural out, unscaled */
CLONES static void fft_dit_inv(const fft_plan *pl, double *re, double *im) {
    const int M = pl->M;
    for (int half = 1, stride = M / 2; half < M; half <<= 1, stride >>= 1) {
        for (int base = 0; base < M; base += 2 * half) {
            double *ar = re + base, *ai = im + base, *br = re + base + half, *bi = im + base + half;
            for (int j = 0; j < half; j++) {
                const double wr = pl->tw_re[j * stride], wi = -pl->tw_im[j * stride];
                const double tr = br[j] * wr - bi[j] * wi, ti = br[j] * wi + bi[j] * wr;
                br[j] = ar[j] - tr; bi[j] = ai[j] - ti;
                ar[j] += tr; ai[j] += ti;
            }
        }
    }
}
/* forward transform of a signed-integer polynomial */
CLONES static void fwd_i64(const fft_plan *pl, const i64 *p, double *re, double *im) {
    const int M = pl->M;
    for (int j = 0; j < M; j++) {
        const double a = (double)p[j], b = (double)p[j + M];
        re[j] = a * pl->twist_re[j] - b * pl->twist_im[j];
        im[j] = a * pl->twist_im[j] + b * pl->twist_re[j];
    }
    fft_dif(pl, re, im);
}
/* Fourier image of a torus polynomial read as signed i64 scaled by 2^-64 (the BSK conversion) */
void orc_fft_forward_torus(const fft_plan *pl, const u64 *p, double *re, double *im) {
    const int M = pl->M;
    const double sc = 1.0 / 18446744073709551616.0;
    for (int j = 0; j < M; j++) {
        const double a = (double)(i64)p[j] * sc, b = (double)(i64)p[j + M] * sc;
        re[j] = a * pl->twist_re[j] - b * pl->twist_im[j];
        im[j] = a * pl->twist_im[j] + b * pl->twist_re[j];
    }
    fft_dif(pl, re, im);
}
/* inverse transform, result added to a torus polynomial: acc += round(frac(x) * 2^64) */
CLONES static void inv_add_torus(const fft_plan *pl, double *re, double *im, u64 *acc) {
    const int M = pl->M;
    const double inv = 1.0 / M;
    fft_dit_inv(pl, re, im);
    for (int j = 0; j < M; j++) {
        /* untwist by exp(-i pi j / N) */
        const double cr = pl->twist_re[j], ci = -pl->twist_im[j];
        double a = (re[j] * cr - im[j] * ci) * inv, b = (re[j] * ci + im[j] * cr) * inv;
        a -= rint(a); b -= rint(b);
        acc[j] += (u64)(i64)llrint(a * 18446744073709551616.0);
        acc[j + M] += (u64)(i64)llrint(b * 18446744073709551616.0);
    }
}

/* Fourier BSK: [n][level][2 rows][2 cols] x (re[M], im[M]) doubles */
size_t orc_fourier_bsk_doubles(const orc_params *p) { return (size_t)p->n * p->pbs_level * 4 * p->N; }
void orc_fourier_bsk(const orc_params *p, const u64 *bsk, double *fbsk) {
    const int N = p->N, M = N / 2;
    fft_plan *pl = orc_fft_plan_new(N);
    const size_t polys = (size_t)p->n * p->pbs_level * 4;
#pragma omp parallel for schedule(static)
    for (size_t q = 0; q < polys; q++) orc_fft_forward_torus(pl, bsk + q * N, fbsk + q * N, fbsk + q * N + M);
    orc_fft_plan_free(pl);
}

/* FFT external product on one GGSW: acc[2][N] += ggsw_f (x) glwe[2][N].  scratch: 6*M doubles + N i64 */
static void external_product_fft(const orc_params *p, const fft_plan *pl, const double *gf, const u64 *glwe,
                                 u64 *acc, double *scratch, i64 *dig) {
    const int N = p->N, M = N / 2, L = p->pbs_level;
    double *dre = scratch, *dim_ = scratch + M;
    double *o0r = scratch + 2 * M, *o0i = scratch + 3 * M, *o1r = scratch + 4 * M, *o1i = scratch + 5 * M;
    i64 tmp[64];
    memset(o0r, 0, sizeof(double) * 4 * M);
    for (int r = 0; r < 2; r++)
        for (int l = 0; l < L; l++) {
            for (int j = 0; j < N; j++) {
                decompose(glwe[r * N + j], p->pbs_base_log, L, tmp);
                dig[j] = tmp[l];
            }
            fwd_i64(pl, dig, dre, dim_);
            const double *g0 = gf + (((size_t)l * 2 + r) * 2 + 0) * N, *g1 = g0 + N;
            for (int j = 0; j < M; j++) {
                o0r[j] += dre[j] * g0[j] - dim_[j] * g0[j + M];
                o0i[j] += dre[j] * g0[j + M] + dim_[j] * g0[j];
                o1r[j] += dre[j] * g1[j] - dim_[j] * g1[j + M];
                o1i[j] += dre[j] * g1[j + M] + dim_[j] * g1[j];
            }
        }
    inv_add_torus(pl, o0r, o0i, acc);
    inv_add_torus(pl, o1r, o1i, acc + N);
}
void orc_external_product_fft(const orc_params *p, const double *ggsw_f, const u64 *glwe, u64 *acc) {
    fft_plan *pl = orc_fft_plan_new(p->N);
    double *scratch = (double *)malloc(sizeof(double) * 3 * p->N);
    i64 *dig = (i64 *)malloc(sizeof(i64) * p->N);
    external_product_fft(p, pl, ggsw_f, glwe, acc, scratch, dig);
    free(scratch); free(dig); orc_fft_plan_free(pl);
}

void orc_blind_rotate_fft(const orc_params *p, const fft_plan *pl, const double *fbsk, const u64 *ks,
                          const u64 *lut, u64 *acc) {
    const int n = p->n, N = p->N, lg = ilog2(2 * N);
    u64 *rot = (u64 *)malloc(sizeof(u64) * 2 * N);
    double *scratch = (double *)malloc(sizeof(double) * 3 * N);
    i64 *dig = (i64 *)malloc(sizeof(i64) * N);
    u32 bt = modswitch(ks[n], lg) % (2 * N);
    memset(acc, 0, sizeof(u64) * N);
    monomial_mul(acc + N, lut, (2 * N - bt) % (2 * N), N);
    const size_t ggsw_sz = (size_t)p->pbs_level * 4 * N;
    for (int i = 0; i < n; i++) {
        u32 at = modswitch(ks[i], lg) % (2 * N);
        if (at == 0) continue;
        for (int r = 0; r < 2; r++) {
            monomial_mul(rot + r * N, acc + r * N, at, N);
            for (int j = 0; j < N; j++) rot[r * N + j] -= acc[r * N + j];
        }
        external_product_fft(p, pl, fbsk + i * ggsw_sz, rot, acc, scratch, dig);
    }
    free(rot); free(scratch); free(dig);
}

/* FFT-route PBS over a batch, OpenMP across ciphertexts (the "port" CPU baseline).
 * Returns the number of threads used. */
int orc_pbs_fft_batch(const orc_params *p, const double *fbsk, const u64 *ksk, const u64 *luts,
                      const int32_t *lut_ids, const u64 *in, int count, u64 *out) {
    int threads = 1;
#pragma omp parallel
    {
#ifdef _OPENMP
#pragma omp single
        threads = omp_get_num_threads();
#endif
        fft_plan *pl = orc_fft_plan_new(p->N);
        u64 *ks = (u64 *)malloc(sizeof(u64) * (p->n + 1));
        u64 *acc = (u64 *)malloc(sizeof(u64) * 2 * p->N);
#pragma omp for schedule(dynamic, 1)
        for (int b = 0; b < count; b++) {
            orc_keyswitch(p, ksk, in + (size_t)b * (p->N + 1), ks);
            orc_blind_rotate_fft(p, pl, fbsk, ks, luts + (size_t)lut_ids[b] * p->N, acc);
            orc_sample_extract(p->N, acc, out + (size_t)b * (p->N + 1));
        }
        free(ks); free(acc); orc_fft_plan_free(pl);
    }
    return threads;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
