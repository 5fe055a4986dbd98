// Micro-benchmark: issue rate of cvt.rn.f16x2.f32 (F2FP.PACK_AB), ex2.approx (MUFU) and FFMA per SM per clock.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int KIND>
__global__ void k(float* out, int iters) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0) {
        uint32_t h;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x[i]), "f"(x[(i + 1) & 7]));
        x[i] = __uint_as_float(h | 0x3f000000u);
      } else if (KIND == 1) {
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x[i]) : "f"(x[i]));
      } else {
        asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(x[i]) : "f"(x[i]), "f"(1.0001f), "f"(0.5f));
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const char* names[3] = {"cvt.f16x2 (+LOP)", "ex2.approx", "ffma"};
  for (int kind = 0; kind < 3; ++kind) {
    const int iters = 4000, warps = 32;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (kind == 0) k<0><<<148, warps * 32>>>(out, iters);
      else if (kind == 1) k<1><<<148, warps * 32>>>(out, iters);
      else k<2><<<148, warps * 32>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
    }
    double ops = 148.0 * warps * 32 * iters * 8;
    printf("%-18s %.3f ms  %.1f lane-ops/clk/SM\n", names[kind], ms, ops / 148 / (ms * 1e-3 * p.clockRate * 1e3));
  }
  return 0;
}
