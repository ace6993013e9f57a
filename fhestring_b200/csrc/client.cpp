// client.cpp -- host-side client: key generation, encryption, decryption.
// Mirrors MyClientKey of /root/reference/src/client_key.rs:9-106 (from_params -> gen_keys_radix :31,
// encrypt :45-65, decrypt :96-106) and FheAsciiChar::encrypt/decrypt (fheasciichar.rs:27-33): an
// encrypted u8 is 4 blocks of 2 message bits (+2 carry bits, +1 padding bit), little endian.
// Runs on the CPU exactly like the reference's client does; none of this is on the PBS path.
#include <sys/random.h>

#include <cerrno>
#include <cmath>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fhestr_engine.h"

namespace {
typedef uint64_t u64;

// ChaCha20 (RFC 8439 block function) as the generator behind every secret key, mask and noise sample.  The 256-bit key
// comes from the operating system (getrandom) unless the caller asks for a deterministic TEST key; a stream is a
// (domain, index) pair in the 96-bit nonce, so secret keys, the bootstrapping key, the keyswitching key and every
// encrypted block draw from computationally independent streams -- the mask words published in a ciphertext say
// nothing about its noise sample's neighbours or about any key.  (The reference gets the same from tfhe-rs'
// seeder + AES-CTR generator behind gen_keys_radix, /root/reference/src/client_key.rs:31.)
struct ChaChaKey { uint32_t w[8]; };

inline uint32_t rotl32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }
#define FHESTR_QR(a, b, c, d) a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12); \
                              a += b; d ^= a; d = rotl32(d, 8); c += d; b ^= c; b = rotl32(b, 7)
inline void chacha20_block(const ChaChaKey& key, uint32_t counter, uint32_t domain, u64 index, uint32_t out[16]) {
    uint32_t st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                       key.w[0], key.w[1], key.w[2], key.w[3], key.w[4], key.w[5], key.w[6], key.w[7],
                       counter, domain, (uint32_t)index, (uint32_t)(index >> 32)};
    uint32_t x[16];
    for (int i = 0; i < 16; i++) x[i] = st[i];
    for (int r = 0; r < 10; r++) {
        FHESTR_QR(x[0], x[4], x[8], x[12]); FHESTR_QR(x[1], x[5], x[9], x[13]);
        FHESTR_QR(x[2], x[6], x[10], x[14]); FHESTR_QR(x[3], x[7], x[11], x[15]);
        FHESTR_QR(x[0], x[5], x[10], x[15]); FHESTR_QR(x[1], x[6], x[11], x[12]);
        FHESTR_QR(x[2], x[7], x[8], x[13]); FHESTR_QR(x[3], x[4], x[9], x[14]);
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + st[i];
}

enum : uint32_t { kDomLweSecret = 1, kDomGlweSecret = 2, kDomBsk = 3, kDomKsk = 4, kDomEncrypt = 5 };

struct Rng {
    const ChaChaKey& key;
    uint32_t domain, counter = 0;
    u64 index;
    uint32_t buf[16];
    int pos = 16;
    bool have_spare = false;
    double spare = 0;
    Rng(const ChaChaKey& k, uint32_t dom, u64 idx) : key(k), domain(dom), index(idx) {}
    u64 next() {
        if (pos >= 16) { chacha20_block(key, counter++, domain, index, buf); pos = 0; }
        const u64 v = (u64)buf[pos] | ((u64)buf[pos + 1] << 32);
        pos += 2;
        return v;
    }
    double unit() { return ((next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
    double gauss() {
        if (have_spare) { have_spare = false; return spare; }
        const double u = unit(), v = unit();
        const double m = std::sqrt(-2.0 * std::log(u));
        spare = m * std::sin(6.283185307179586476925 * v);
        have_spare = true;
        return m * std::cos(6.283185307179586476925 * v);
    }
    u64 torus_noise(double std_dev) { return (u64)(int64_t)std::llrint(gauss() * std_dev * 18446744073709551616.0); }
};

// key material: 32 bytes from the OS, or (tests, benchmarks) expanded from a 64-bit seed
bool make_key(u64 seed, ChaChaKey& out) {
    if (seed == 0) {
        unsigned char raw[32];
        size_t got = 0;
        while (got < sizeof raw) {
            const ssize_t r = getrandom(raw + got, sizeof raw - got, 0);
            if (r < 0) { if (errno == EINTR) continue; return false; }
            got += (size_t)r;
        }
        memcpy(out.w, raw, sizeof raw);
        return true;
    }
    ChaChaKey tmp{};
    tmp.w[0] = (uint32_t)seed; tmp.w[1] = (uint32_t)(seed >> 32);
    tmp.w[2] = 0x74736574u; tmp.w[3] = 0x79656b2du;   // "test-key": a seeded key is for tests only
    uint32_t blk[16];
    chacha20_block(tmp, 0, 0, 0, blk);
    memcpy(out.w, blk, sizeof out.w);
    return true;
}

template <class F>
void parallel_for(int n, F f) {
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)(hw ? hw : 4);
    if (nt > n) nt = n;
    if (nt <= 1) { for (int i = 0; i < n; i++) f(i); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([=] { for (int i = t; i < n; i += nt) f(i); });
    for (auto& x : th) x.join();
}
}  // namespace

struct fhestr_client {
    fhestr_params prm;
    double lwe_std, glwe_std;
    ChaChaKey key;
    u64 enc_counter = 0;
    std::vector<uint8_t> s_lwe, s_glwe;
};

#pragma GCC visibility push(default)
extern "C" {

int fhestr_client_create(const fhestr_params* p, double lwe_std, double glwe_std, uint64_t seed, fhestr_client** out) {
    if (!p || !out || p->k != 1 || p->n < 1 || p->N < 2 || (p->N & (p->N - 1))) return FHESTR_E_INVALID;
    fhestr_client* c = new (std::nothrow) fhestr_client();
    if (!c) return FHESTR_E_STATE;
    c->prm = *p; c->lwe_std = lwe_std; c->glwe_std = glwe_std;
    if (!make_key(seed, c->key)) { delete c; return FHESTR_E_STATE; }   // no entropy: refuse, never a guessable key
    c->s_lwe.resize(p->n); c->s_glwe.resize(p->N);
    Rng r1(c->key, kDomLweSecret, 0), r2(c->key, kDomGlweSecret, 0);
    for (auto& b : c->s_lwe) b = (uint8_t)(r1.next() >> 63);
    for (auto& b : c->s_glwe) b = (uint8_t)(r2.next() >> 63);
    *out = c;
    return FHESTR_OK;
}

void fhestr_client_destroy(fhestr_client* c) { delete c; }

int fhestr_client_secret_keys(const fhestr_client* c, uint8_t* s_lwe, uint8_t* s_glwe) {
    if (!c) return FHESTR_E_INVALID;
    if (s_lwe) memcpy(s_lwe, c->s_lwe.data(), c->s_lwe.size());
    if (s_glwe) memcpy(s_glwe, c->s_glwe.data(), c->s_glwe.size());
    return FHESTR_OK;
}

int fhestr_client_server_keys(fhestr_client* c, uint64_t* bsk, uint64_t* ksk) {
    if (!c || !bsk || !ksk) return FHESTR_E_INVALID;
    const fhestr_params& p = c->prm;
    const int n = p.n, N = p.N, L = p.pbs_level, KL = p.ks_level;
    // bootstrapping key: GGSW(s_lwe[i]) = rows (level, r) = GLWE encryption of zero + s_i * q/B^level on
    // coefficient 0 of polynomial r
    parallel_for(n, [&](int i) {
        Rng rng(c->key, kDomBsk, (u64)i);
        for (int lvl = 1; lvl <= L; lvl++)
            for (int row = 0; row < 2; row++) {
                u64* A = bsk + ((((size_t)i * L + (lvl - 1)) * 2 + row) * 2) * N;
                u64* B = A + N;
                for (int j = 0; j < N; j++) A[j] = rng.next();
                for (int j = 0; j < N; j++) B[j] = rng.torus_noise(c->glwe_std);
                for (int j = 0; j < N; j++) {  // B += A * S (negacyclic), S binary
                    if (!c->s_glwe[j]) continue;
                    for (int u = 0; u < N - j; u++) B[u + j] += A[u];
                    for (int u = N - j; u < N; u++) B[u + j - N] -= A[u];
                }
                const u64 g = ((u64)c->s_lwe[i]) << (64 - p.pbs_base_log * lvl);
                if (row == 0) A[0] += g; else B[0] += g;
            }
    });
    // keyswitching key big -> small
    parallel_for(N, [&](int i) {
        Rng rng(c->key, kDomKsk, (u64)i);
        for (int lvl = 1; lvl <= KL; lvl++) {
            u64* ct = ksk + ((size_t)i * KL + (lvl - 1)) * (n + 1);
            u64 body = rng.torus_noise(c->lwe_std);
            for (int k = 0; k < n; k++) {
                ct[k] = rng.next();
                if (c->s_lwe[k]) body += ct[k];
            }
            ct[n] = body + (((u64)c->s_glwe[i]) << (64 - p.ks_base_log * lvl));
        }
    });
    return FHESTR_OK;
}

int fhestr_client_encrypt_blocks(fhestr_client* c, const uint8_t* values, uint32_t count, uint64_t* cts) {
    if (!c || !values || !cts) return FHESTR_E_INVALID;
    const int N = c->prm.N;
    const u64 base = c->enc_counter;
    c->enc_counter += count;
    parallel_for((int)count, [&](int b) {
        Rng rng(c->key, kDomEncrypt, base + (u64)b);
        u64* ct = cts + (size_t)b * (N + 1);
        u64 body = (((u64)values[b]) << c->prm.delta_log) + rng.torus_noise(c->glwe_std);
        for (int i = 0; i < N; i++) {
            ct[i] = rng.next();
            if (c->s_glwe[i]) body += ct[i];
        }
        ct[N] = body;
    });
    return FHESTR_OK;
}

int fhestr_client_decrypt_blocks(const fhestr_client* c, const uint64_t* cts, uint32_t count, uint8_t* values,
                                 int64_t* phase_err) {
    if (!c || !cts || !values) return FHESTR_E_INVALID;
    const int N = c->prm.N, dl = c->prm.delta_log;
    parallel_for((int)count, [&](int b) {
        const u64* ct = cts + (size_t)b * (N + 1);
        u64 ph = ct[N];
        for (int i = 0; i < N; i++) if (c->s_glwe[i]) ph -= ct[i];
        const u64 v = (ph + (1ull << (dl - 1))) >> dl;
        values[b] = (uint8_t)(v & ((1ull << (63 - dl)) - 1));
        if (phase_err) phase_err[b] = (int64_t)(ph - (v << dl));
    });
    return FHESTR_OK;
}

int fhestr_client_encrypt_u8(fhestr_client* c, const uint8_t* bytes, uint32_t count, uint64_t* cts) {
    if (!c || !bytes || !cts) return FHESTR_E_INVALID;
    std::vector<uint8_t> blocks((size_t)count * 4);
    for (uint32_t i = 0; i < count; i++)
        for (int b = 0; b < 4; b++) blocks[(size_t)i * 4 + b] = (bytes[i] >> (2 * b)) & 3;
    return fhestr_client_encrypt_blocks(c, blocks.data(), count * 4, cts);
}

int fhestr_client_decrypt_u8(const fhestr_client* c, const uint64_t* cts, uint32_t count, uint8_t* bytes) {
    if (!c || !bytes || !cts) return FHESTR_E_INVALID;
    std::vector<uint8_t> blocks((size_t)count * 4);
    int rc = fhestr_client_decrypt_blocks(c, cts, count * 4, blocks.data(), nullptr);
    if (rc) return rc;
    for (uint32_t i = 0; i < count; i++) {
        unsigned v = 0;  // RadixClientKey::decrypt::<u8>: sum of (msg+carry) << 2i, wrapping to u8
        for (int b = 0; b < 4; b++) v += (unsigned)blocks[(size_t)i * 4 + b] << (2 * b);
        bytes[i] = (uint8_t)v;
    }
    return FHESTR_OK;
}

}  // extern "C"
#pragma GCC visibility pop
