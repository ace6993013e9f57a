// Host emulation of the four-warp blind-rotation thread program (fhestring_b200/csrc/br_quad.cuh):
// 128 std::threads stand in for the four warps of one PBS, std::barrier for __syncwarp / bar.sync.
// Built by tests/test_br_quad_emulation.py with g++ -std=c++20.
#include <barrier>
#include <cstring>
#include <thread>
#include <vector>

#include "../../fhestring_b200/csrc/br_quad.cuh"

using namespace fhestr;

struct QuadHostCtx {
    int tau_, poly_;
    acc_t* acc_;
    cplx* exch_;
    cplx* exch_partner_;
    uint16_t* atilde_;
    std::barrier<>* warp_bar;
    std::barrier<>* poly_bar;
    std::barrier<>* cta_bar;
    int tau() const { return tau_; }
    int poly() const { return poly_; }
    acc_t* acc() { return acc_; }
    cplx* exch() { return exch_; }
    const cplx* exch_partner() { return exch_partner_; }
    uint16_t* atilde() { return atilde_; }
    void syncwarp() { warp_bar->arrive_and_wait(); }
    void poly_sync() { poly_bar->arrive_and_wait(); }
    void cta_sync() { cta_bar->arrive_and_wait(); }
    cplx ldg(const cplx* p) const { return *p; }
};

struct Tables {
    std::vector<cplx> t;
    QuadTables qt;
    Tables() : t(1024 + 1024 + 64) {
        make_quad_tables(t.data(), t.data() + 1024, t.data() + 2048);
        qt = QuadTables{t.data(), t.data() + 1024, t.data() + 2048};
    }
};

extern "C" {

// bsk_std: [n][2 rows][2 cols][N] u64.  out: [n][kQBskStepElems] complex
void quad_emu_convert_bsk(int n, const u64* bsk_std, double* out) {
    Tables T;
    std::vector<cplx> ex(kQExchCplx);
    std::barrier<> wb0(32), wb1(32), pb(64);
    std::vector<std::thread> th;
    for (int tau = 0; tau < 64; tau++)
        th.emplace_back([&, tau] {
            QuadHostCtx c{tau, 0, nullptr, ex.data(), nullptr, nullptr, tau < 32 ? &wb0 : &wb1, &pb, nullptr};
            for (int i = 0; i < n; i++)
                for (int row = 0; row < 2; row++)
                    for (int col = 0; col < 2; col++)
                        quad_bsk_poly_forward(c, bsk_std + (((size_t)i * 2 + row) * 2 + col) * kN,
                                              reinterpret_cast<cplx*>(out) + (size_t)i * kQBskStepElems, row, col, T.qt);
        });
    for (auto& t : th) t.join();
}

// one PBS blind rotation; bsk_f from quad_emu_convert_bsk.  init_acc/out_lwe/out_acc may be null.
void quad_emu_blind_rotate(int n, const u64* ks, const u64* lut, const u64* init_acc, const double* bsk_f,
                           u64* out_lwe, u64* out_acc) {
    Tables T;
    std::vector<acc_t> acc(2 * kN);
    std::vector<cplx> ex(2 * kQExchCplx);
    std::vector<uint16_t> at(n + 128);
    std::barrier<> wb[4] = {std::barrier<>(32), std::barrier<>(32), std::barrier<>(32), std::barrier<>(32)};
    std::barrier<> pb[2] = {std::barrier<>(64), std::barrier<>(64)};
    std::barrier<> cb(128);
    BrJobView job{ks, lut, init_acc, out_lwe, out_acc, n};
    std::vector<std::thread> th;
    for (int tid = 0; tid < 128; tid++)
        th.emplace_back([&, tid] {
            const int warp = tid >> 5, p = warp >> 1;
            QuadHostCtx c{tid & 63, p, acc.data() + p * kN, ex.data() + p * kQExchCplx,
                          ex.data() + (1 - p) * kQExchCplx, at.data(), &wb[warp], &pb[p], &cb};
            quad_thread_main(c, job, reinterpret_cast<const cplx*>(bsk_f), T.qt);
        });
    for (auto& t : th) t.join();
}
}
