// probe: how does the hardware place the warps of small CTAs on SM sub-partitions (warp slots)?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(unsigned* out) {
    unsigned smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    if ((threadIdx.x & 31) == 0) {
        int w = threadIdx.x >> 5;
        int idx = (blockIdx.x * (blockDim.x >> 5) + w) * 2;
        out[idx] = smid; out[idx + 1] = warpid;
    }
    // stay resident a little so that CTAs co-reside
    long long t0 = clock64(); while (clock64() - t0 < 200000) {}
}
int main() {
    for (int threads : {64, 128, 256}) {
        int grid = 148 * (256 / threads);
        unsigned* d; cudaMalloc(&d, grid * (threads / 32) * 2 * sizeof(unsigned));
        // same footprint as the blind-rotation kernel: 51712 B dynamic smem per 64 threads
        size_t smem = 51712 * (threads / 64);
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe<<<grid, threads, smem>>>(d);
        cudaDeviceSynchronize();
        int n = grid * (threads / 32);
        unsigned* h = new unsigned[n * 2];
        cudaMemcpy(h, d, n * 2 * sizeof(unsigned), cudaMemcpyDeviceToHost);
        printf("threads=%d:", threads);
        for (int b = 0; b < grid; b++) if (h[b * (threads / 32) * 2] == h[0]) {
            printf(" [cta %d:", b);
            for (int w = 0; w < threads / 32; w++) printf(" %u", h[(b * (threads / 32) + w) * 2 + 1]);
            printf("]");
        }
        printf("\n");
        cudaFree(d); delete[] h;
    }
    return 0;
}
