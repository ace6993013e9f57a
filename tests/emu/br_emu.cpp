// Host emulation of the blind-rotation thread program (fhestring_b200/csrc/br_core.cuh):
// 64 std::threads stand in for the two warps of one PBS, std::barrier for __syncwarp / bar.sync.
// Built by tests/test_br_emulation.py with g++ -std=c++20; checks the kernel's index logic, FFT
// layout and Fourier-BSK layout against the oracle on a machine with no GPU.
#include <barrier>
#include <cstring>
#include <thread>
#include <vector>

#include "../../fhestring_b200/csrc/br_core.cuh"

using namespace fhestr;

struct HostCtx {
    int lane_, poly_;
    acc_t* acc_;
    double* xbuf_;
    double* xbuf_partner_;
    uint16_t* atilde_;
    std::barrier<>* warp_bar;
    std::barrier<>* pair_bar;
    int lane() const { return lane_; }
    int poly() const { return poly_; }
    acc_t* acc() { return acc_; }
    double* xbuf() { return xbuf_; }
    double* xbuf_partner() { return xbuf_partner_; }
    uint16_t* atilde() { return atilde_; }
    void syncwarp() { warp_bar->arrive_and_wait(); }
    void pair_sync() { pair_bar->arrive_and_wait(); }
    cplx ldg(const cplx* p) const { return *p; }
    // the device keeps the twiddles in tensor memory; the emulation reads the 8 twiddles of chunk ch from the table
    void tw_ld(int ch, uint32_t (&r)[32], const cplx* tf) const {
        for (int j = 0; j < 8; j++) std::memcpy(&r[4 * j], &tf[(ch * 8 + j) * 32 + lane_], sizeof(cplx));
    }
    void tw_wait(uint32_t (&)[32]) const {}
    static double tw_word(uint32_t lo, uint32_t hi) {
        const u64 b = ((u64)hi << 32) | lo;
        double d;
        std::memcpy(&d, &b, 8);
        return d;
    }
    // word ((x >> 2) mod N) of this polynomial's accumulator, negated when bit 13 of the byte offset is set
    acc_t acc_ld_rot(uint32_t x) const {
        const acc_t v = acc_[(x >> 2) & (kN - 1)];
        return (x & 0x2000u) ? (acc_t)0 - v : v;
    }
};

extern "C" {

// bsk_std: [n][2 rows][2 cols][N] u64 (pbs_level == 1).  out: [n][kBskStepElems] complex
void emu_convert_bsk(int n, const u64* bsk_std, double* out) {
    std::vector<cplx> tf(1024);
    make_twiddles(tf.data());
    std::vector<double> xbuf(kWarpXbufDoubles);
    std::barrier<> wb(32);
    std::vector<std::thread> th;
    for (int lane = 0; lane < 32; lane++)
        th.emplace_back([&, lane] {
            HostCtx c{lane, 0, nullptr, xbuf.data(), nullptr, nullptr, &wb, nullptr};
            for (int i = 0; i < n; i++)
                for (int row = 0; row < 2; row++)
                    for (int col = 0; col < 2; col++)
                        bsk_poly_forward(c, bsk_std + (((size_t)i * 2 + row) * 2 + col) * kN,
                                         reinterpret_cast<cplx*>(out) + (size_t)i * kBskStepElems, row, col,
                                         tf.data());
        });
    for (auto& t : th) t.join();
}

// one PBS blind rotation; bsk_f from emu_convert_bsk.  init_acc/out_lwe/out_acc may be null.
void emu_blind_rotate(int n, const u64* ks, const u64* lut, const u64* init_acc, const double* bsk_f,
                      u64* out_lwe, u64* out_acc) {
    std::vector<cplx> tf(1024);
    make_twiddles(tf.data());
    std::vector<acc_t> acc(2 * kN);
    std::vector<double> xbuf(2 * kWarpXbufDoubles);
    std::vector<uint16_t> at(n + 64);
    std::barrier<> wb0(32), wb1(32), pb(64);
    BrJobView job{ks, lut, init_acc, out_lwe, out_acc, n};
    std::vector<std::thread> th;
    for (int w = 0; w < 2; w++)
        for (int lane = 0; lane < 32; lane++)
            th.emplace_back([&, w, lane] {
                HostCtx c{lane, w, acc.data() + w * kN, xbuf.data() + w * kWarpXbufDoubles,
                          xbuf.data() + (1 - w) * kWarpXbufDoubles, at.data(), w ? &wb1 : &wb0, &pb};
                br_thread_main(c, job, reinterpret_cast<const cplx*>(bsk_f), tf.data());
            });
    for (auto& t : th) t.join();
}
}
