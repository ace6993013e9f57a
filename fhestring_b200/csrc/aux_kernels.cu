// aux_kernels.cu -- K5 (LUT accumulator generation), K6 (key conversion), trivial ciphertexts and the
// DFMA peak microbenchmark, for sm_100a.
#include "kernels.cuh"

namespace fhestr {

// ---- K5: 16-entry function table -> body polynomial of the trivial GLWE accumulator:
// boxes of N/16 coefficients f(i) << delta_log, first half-box negated, rotated left by half a box.
// Replaces shortint generate_lookup_table (used by every op of fheasciichar.rs:35-104).
__global__ void lut_poly_kernel(const uint8_t* table, int entries, int delta_log, u64* out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= kN) return;
    const int box = kN / entries;
    const int m = (j + box / 2) & (kN - 1);
    const uint8_t e = table[m / box];
    // a half-step table (fhestr_lut_register: entries 0x80 | e) holds e - 1/2; the engine adds the 1/2 back after the
    // sample extract, so the negacyclic half reads 1 - e
    const u64 v = (e & 0x80) ? (((u64)(e & 0x7f)) << delta_log) - ((u64)1 << (delta_log - 1)) : ((u64)e) << delta_log;
    out[j] = (m < box / 2) ? (u64)0 - v : v;
}
int launch_lut_poly(const uint8_t* table_dev, int entries, int delta_log, u64* out, cudaStream_t s) {
    lut_poly_kernel<<<kN / 256, 256, 0, s>>>(table_dev, entries, delta_log, out);
    return 1;
}

// ---- K6: standard-domain GGSW polynomials -> Fourier BSK in the engine layout.  One warp per
// polynomial, same forward transform as the blind rotation (br_core.cuh: bsk_poly_forward).
struct ConvCtx {
    int lane_;
    double* xbuf_;
    __device__ __forceinline__ int lane() const { return lane_; }
    __device__ __forceinline__ double* xbuf() { return xbuf_; }
    __device__ __forceinline__ void syncwarp() { __syncwarp(); }
    __device__ __forceinline__ cplx ldg(const cplx* p) const {
        const double2 v = __ldg(reinterpret_cast<const double2*>(p));
        return cplx{v.x, v.y};
    }
    // the blind rotation keeps its twiddles in tensor memory; the conversion (once per key) reads the chunk from the table
    __device__ __forceinline__ void tw_ld(int ch, uint32_t (&r)[32], const cplx* tf) const {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double2 w = __ldg(reinterpret_cast<const double2*>(tf + (ch * 8 + j) * 32 + lane_));
            r[4 * j] = (uint32_t)__double2loint(w.x); r[4 * j + 1] = (uint32_t)__double2hiint(w.x);
            r[4 * j + 2] = (uint32_t)__double2loint(w.y); r[4 * j + 3] = (uint32_t)__double2hiint(w.y);
        }
    }
    __device__ __forceinline__ void tw_wait(uint32_t (&)[32]) const {}
    static __device__ __forceinline__ double tw_word(uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }
};
__global__ void __launch_bounds__(64) bsk_convert_kernel(const u64* bsk_std, int n_polys, const cplx* tf, cplx* out) {
    __shared__ __align__(16) double xb[2][kWarpXbufDoubles];
    const int warp = threadIdx.x >> 5;
    const int q = blockIdx.x * 2 + warp;  // polynomial index = (step*2 + row)*2 + col
    if (q >= n_polys) return;
    ConvCtx c{(int)(threadIdx.x & 31), xb[warp]};
    const int step = q >> 2, row = (q >> 1) & 1, col = q & 1;
    bsk_poly_forward(c, bsk_std + (size_t)q * kN, out + (size_t)step * kBskStepElems, row, col, tf);
}
int launch_bsk_convert(const u64* bsk_std, int n, const cplx* tf, cplx* out, cudaStream_t s) {
    const int polys = n * 4;
    bsk_convert_kernel<<<(polys + 1) / 2, 64, 0, s>>>(bsk_std, polys, tf, out);
    return 1;
}

// ---- trivial ciphertexts: mask 0, body value << delta_log (create_trivial_radix, fheasciichar.rs:23)
__global__ void trivial_kernel(u64* arena, uint32_t first, uint32_t count, const uint8_t* values, int delta_log) {
    const uint32_t b = blockIdx.x;
    if (b >= count) return;
    u64* ct = arena + (size_t)(first + b) * (kN + 1);
    for (int idx = threadIdx.x; idx <= kN; idx += blockDim.x)
        ct[idx] = (idx == kN) ? ((u64)values[b]) << delta_log : 0ull;
}
int launch_trivial(u64* arena, uint32_t first, uint32_t count, const uint8_t* values_dev, int delta_log, cudaStream_t s) {
    if (!count) return 0;
    trivial_kernel<<<count, 256, 0, s>>>(arena, first, count, values_dev, delta_log);
    return 1;
}

// ---- scattered arena blocks -> one contiguous staging buffer (fhestr_ct_download_slots: the result chars of a string
// method are wherever their last level left them).  16 KiB per block, coalesced 16-byte... 8-byte words.
__global__ void gather_blocks_kernel(const u64* arena, const uint32_t* slots, uint32_t count, u64* out) {
    const uint32_t b = blockIdx.x;
    if (b >= count) return;
    const u64* src = arena + (size_t)slots[b] * (kN + 1);
    u64* dst = out + (size_t)b * (kN + 1);
    for (int i = threadIdx.x; i <= kN; i += blockDim.x) dst[i] = src[i];
}
int launch_gather_blocks(const u64* arena, const uint32_t* slots_dev, uint32_t count, u64* out, cudaStream_t s) {
    if (!count) return 0;
    gather_blocks_kernel<<<count, 256, 0, s>>>(arena, slots_dev, count, out);
    return 1;
}

// ---- cross-GPU level barrier over NVLink peer memory (one process per GPU, arenas and flag arrays mapped with
// cudaIpc).  The blind rotation of a level has already stored its results into every peer's arena; this only
// orders those stores before the next level anywhere reads them.
struct PeerFlags { uint32_t* p[8]; };
__global__ void peer_signal_kernel(PeerFlags peers, int rank, int world, uint32_t epoch) {
    const int r = threadIdx.x;
    if (r >= world || r == rank) return;
    __threadfence_system();   // the previous kernel's remote stores are complete at kernel end; keep the flag behind them
    *reinterpret_cast<volatile uint32_t*>(peers.p[r] + rank) = epoch;
}
__global__ void peer_wait_kernel(uint32_t* my_flags, uint32_t* status, int rank, int world, uint32_t epoch) {
    const int r = threadIdx.x;
    if (r >= world || r == rank) return;
    const long long t0 = clock64();
    while ((int32_t)(*reinterpret_cast<volatile uint32_t*>(my_flags + r) - epoch) < 0) {
        if (clock64() - t0 > 20000000000LL) { atomicExch(status, 1u); return; }   // ~10 s: a peer died; report, do not hang
        __nanosleep(200);
    }
    __threadfence_system();
}
int launch_peer_barrier(uint32_t* const* peer_flags, uint32_t* my_flags, uint32_t* status, int rank, int world, uint32_t epoch, cudaStream_t s) {
    PeerFlags pf{};
    for (int r = 0; r < world && r < 8; r++) pf.p[r] = peer_flags[r];
    peer_signal_kernel<<<1, 32, 0, s>>>(pf, rank, world, epoch);
    peer_wait_kernel<<<1, 32, 0, s>>>(my_flags, status, rank, world, epoch);
    return 2;
}

// ---- DFMA peak: 16 independent FMA chains per thread, 8 warps x 4 CTAs per SM
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* sink, int iters) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const double m = 1.0000000001, c = 1e-12;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = fma(a[i], m, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    if (s == 123.456) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int launch_dfma_peak(double* sink, int iters, unsigned long long* fmas, cudaStream_t s) {
    const int grid = 148 * 8, block = 256;
    dfma_peak_kernel<<<grid, block, 0, s>>>(sink, iters);
    *fmas = (unsigned long long)grid * block * (unsigned long long)iters * 8ull * 16ull;
    return 1;
}

}  // namespace fhestr
