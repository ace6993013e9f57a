#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
__global__ void __launch_bounds__(256) imma_probe(int* out, int iters) {
    uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    int c[8][4];
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) c[i][j] = 0;
    #pragma unroll 1
    for (int it = 0; it < iters; it++) {
        #pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    int s = 0;
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    int* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int ctas : {148, 148 * 2, 148 * 4}) {
        const int iters = 20000;
        imma_probe<<<ctas, 256>>>(out, iters); cudaDeviceSynchronize();
        cudaEventRecord(e0); imma_probe<<<ctas, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double macs = (double)ctas * 8 /*warps*/ * iters * 8 * (16.0 * 8 * 32);
        printf("ctas/SM=%d  %.1f TMAC/s (%.1f TOPS)  per SM per clk: %.0f MAC  err=%s\n", ctas / 148, macs / ms / 1e9, 2 * macs / ms / 1e9,
               macs / (ms * 1e-3) / 148 / 1.965e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
