// Host emulation of the latency form of the blind rotation (fhestring_b200/csrc/br_wide.cuh): 128 std::threads stand in
// for the four warps of one PBS, std::barrier for __syncthreads; the bulk-TMA key pipeline is a plain pointer.
// Built by tests/test_br_wide_emulation.py with g++ -std=c++20; checks the transform layout (three radix-8 stages,
// the half level, both exchanges), the Fourier-key layout, rotation, mod-switch and sample extract against the oracle.
#include <barrier>
#include <cstring>
#include <thread>
#include <vector>

#include "../../fhestring_b200/csrc/br_wide.cuh"

using namespace fhestr;

#ifndef EMU_PIPELINED
#define EMU_PIPELINED 0
#endif
struct HostWideCtx {
    static constexpr bool kPipelined = EMU_PIPELINED != 0;
    int tid_;
    acc_t* acc_;            // [2][kN]
    cplx* buf0_;
    cplx* buf1_;
    uint16_t* atilde_;
    const cplx* bsk_;       // [n][kWKeyTile]
    std::barrier<>* bar;
    std::barrier<>* pbar0 = nullptr;   // one barrier per polynomial (the pipelined step)
    std::barrier<>* pbar1 = nullptr;
    void sync_poly(int p) { (p ? pbar1 : pbar0)->arrive_and_wait(); }
    int tid() const { return tid_; }
    acc_t* acc(int p) { return acc_ + p * kN; }
    cplx* buf0() { return buf0_; }
    cplx* buf1() { return buf1_; }
    uint16_t* atilde() { return atilde_; }
    void sync() { bar->arrive_and_wait(); }
    void key_prefetch(int) {}
    void key_prefetch_current(int) {}
    void key_release(int) {}
    void pre_write_sync() {}
    const cplx* key_wait(int step) { return bsk_ + (size_t)step * kWKeyTile; }
    acc_t acc_ld_rot(int p, uint32_t x) const {
        const acc_t v = acc_[p * kN + ((x >> 2) & (kN - 1))];
        return (x & 0x2000u) ? (acc_t)0 - v : v;
    }
};

extern "C" {

// bsk_std: [n][2 rows][2 cols][N] u64.  out: [n][kWKeyTile] complex
void emu_wide_convert_bsk(int n, const u64* bsk_std, double* out) {
    std::vector<WideConsts> tab(kWT);
    make_wide_consts(tab.data());
    std::vector<cplx> b0(kWBuf0), b1(kWBuf1);
    std::barrier<> bar(kWT);
    std::vector<std::thread> th;
    for (int t = 0; t < kWT; t++)
        th.emplace_back([&, t] {
            HostWideCtx c{t, nullptr, b0.data(), b1.data(), nullptr, nullptr, &bar};
            for (int i = 0; i < n; i++)
                for (int row = 0; row < 2; row++)
                    for (int col = 0; col < 2; col++)
                        wide_bsk_poly_forward(c, bsk_std + (((size_t)i * 2 + row) * 2 + col) * kN,
                                              reinterpret_cast<cplx*>(out) + (size_t)i * kWKeyTile, row, col, tab[t]);
        });
    for (auto& t : th) t.join();
}

void emu_wide_blind_rotate(int n, const u64* ks, const u64* lut, const u64* init_acc, const double* bsk_w,
                           u64* out_lwe, u64* out_acc) {
    std::vector<WideConsts> tab(kWT);
    make_wide_consts(tab.data());
    std::vector<acc_t> acc(2 * kN);
    std::vector<cplx> b0(kWBuf0), b1(kWBuf1);
    std::vector<uint16_t> at(n + 256);
    std::barrier<> bar(kWT), pb0(kWT), pb1(kWT);
    BrJobView job{ks, lut, init_acc, out_lwe, out_acc, n};
    std::vector<std::thread> th;
    for (int t = 0; t < kWT; t++)
        th.emplace_back([&, t] {
            HostWideCtx c{t, acc.data(), b0.data(), b1.data(), at.data(), reinterpret_cast<const cplx*>(bsk_w), &bar, &pb0, &pb1};
            wide_thread_main(c, job, tab[t]);
        });
    for (auto& t : th) t.join();
}

// forward then inverse of one folded polynomial (no key): out = 1024 * in, checks the transform pair alone;
// spectrum (optional) receives the 1024 spectrum points in (u, t) order
void emu_wide_roundtrip(const double* in_re, const double* in_im, double* out_re, double* out_im, double* spec) {
    std::vector<WideConsts> tab(kWT);
    make_wide_consts(tab.data());
    std::vector<cplx> b0(kWBuf0), b1(kWBuf1);
    std::barrier<> bar(kWT);
    std::vector<std::thread> th;
    for (int t = 0; t < kWT; t++)
        th.emplace_back([&, t] {
            HostWideCtx c{t, nullptr, b0.data(), b1.data(), nullptr, nullptr, &bar};
            double re[1][8], im[1][8];
            for (int k = 0; k < 8; k++) { re[0][k] = in_re[128 * k + t]; im[0][k] = in_im[128 * k + t]; }
            wide_forward<1>(c, re, im, tab[t]);
            if (spec) for (int u = 0; u < 8; u++) { spec[2 * (u * kWT + t)] = re[0][u]; spec[2 * (u * kWT + t) + 1] = im[0][u]; }
            c.sync();
            wide_inverse<1>(c, re, im, tab[t]);
            for (int k = 0; k < 8; k++) { out_re[128 * k + t] = re[0][k]; out_im[128 * k + t] = im[0][k]; }
        });
    for (auto& t : th) t.join();
}
}
