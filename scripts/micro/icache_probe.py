"""Microbenchmark generator: straight-line loop bodies of growing size (mix of independent DFMA chains and integer
ops), 64-thread CTAs, 4 per SM -- how does issue rate depend on the body size once it exceeds the instruction
caches?  usage: python icache_probe.py > /tmp/icache_probe.cu ; nvcc ... ; ./a.out"""
import sys
sizes = [512, 1024, 2048, 3072, 4096, 5120, 8192]   # instructions per loop body (approx)
print("#include <cstdio>\n#include <cuda_runtime.h>")
for kind in ("fp64", "int", "mix"):
    for n in sizes:
        print(f"__global__ void __launch_bounds__(64, 4) k_{kind}_{n}(double* out, int iters, int seed) {{")
        print("  double a[16]; unsigned u[16];")
        print("  for (int i = 0; i < 16; i++) { a[i] = 1.0 + 1e-9 * (threadIdx.x + i + seed); u[i] = threadIdx.x * 7 + i + seed; }")
        print("  const double m = 1.0000000001, c = 1e-12;")
        print("  #pragma unroll 1\n  for (int it = 0; it < iters; it++) {")
        for j in range(n // 16):
            for i in range(16):
                if kind == "fp64" or (kind == "mix" and (i % 2 == 0)):
                    # distinct constants defeat any CSE / rerolling
                    print(f"    a[{i}] = fma(a[{i}], m, c + {j * 16 + i}e-20);")
                else:
                    print(f"    u[{i}] = (u[{i}] ^ {j * 16 + i + 1}u) + (u[{(i + 1) % 16}] >> {1 + (j + i) % 7});")
        print("  }")
        print("  double s = 0; unsigned w = 0; for (int i = 0; i < 16; i++) { s += a[i]; w += u[i]; }")
        print("  if (s == 123.456 || w == 12345u) out[blockIdx.x * blockDim.x + threadIdx.x] = s + w;")
        print("}")
print("int main() {")
print("  double* out; cudaMalloc(&out, 1 << 24); cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);")
print("  int dev_clock; cudaDeviceGetAttribute(&dev_clock, cudaDevAttrClockRate, 0);")
for kind in ("fp64", "int", "mix"):
    for n in sizes:
        print(f"  for (int ctas : {{148 * 1, 148 * 4}}) {{")
        print(f"    int iters = 4000000 / {n}; k_{kind}_{n}<<<ctas, 64>>>(out, iters, 1); cudaDeviceSynchronize();")
        print(f"    cudaEventRecord(e0); k_{kind}_{n}<<<ctas, 64>>>(out, iters, 2); cudaEventRecord(e1); cudaEventSynchronize(e1);")
        print(f"    float ms; cudaEventElapsedTime(&ms, e0, e1);")
        print(f"    double instr = (double)iters * {n}; double cyc = ms * 1e-3 * 1.965e9;")
        print(f"    printf(\"{kind} body={n} ctas/SM=%d  cycles/instr/warp=%.3f  err=%s\\n\", ctas / 148, cyc / instr, cudaGetErrorString(cudaGetLastError()));")
        print("  }")
print("  return 0;\n}")
