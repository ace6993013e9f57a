// probe: which warps of a 256-thread CTA share an SM sub-partition (FP64 pipe)?  Two chosen warps run a DFMA chain
// mix that saturates one sub-partition's FP64 pipe; the pair takes twice as long when both sit on the same one.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 1) probe(int wa, int wb, double* sink, long long* cyc, unsigned* wid) {
    const int w = threadIdx.x >> 5;
    unsigned warpid; asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    if ((threadIdx.x & 31) == 0) wid[w] = warpid;
    __syncthreads();
    if (w != wa && w != wb) return;
    double x[8];
    for (int i = 0; i < 8; i++) x[i] = 1.0 + threadIdx.x * 1e-9 + i;
    long long t0 = clock64();
    for (int it = 0; it < 4096; it++)
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = fma(x[i], 1.0000001, 1e-9);
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < 8; i++) s += x[i];
    sink[threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0 && w == wa) cyc[0] = t1 - t0;
}
int main() {
    double* sink; long long* cyc; unsigned* wid;
    cudaMalloc(&sink, 256 * 8); cudaMalloc(&cyc, 8); cudaMalloc(&wid, 32);
    unsigned hw[8];
    for (int a = 0; a < 8; a++) {
        printf("warp %d:", a);
        for (int b = 0; b < 8; b++) {
            probe<<<1, 256>>>(a, b, sink, cyc, wid);
            long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf(" %6lld", h);
        }
        printf("\n");
    }
    cudaMemcpy(hw, wid, 32, cudaMemcpyDeviceToHost);
    printf("%%warpid of warps 0..7:"); for (int i = 0; i < 8; i++) printf(" %u", hw[i]); printf("\n");
    return 0;
}
