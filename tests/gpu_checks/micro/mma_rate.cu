// Micro-benchmark: legacy mma.sync throughput on this GPU (TF32 m16n8k8, BF16 m16n8k16), per SM per clock.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
template <int KIND>
__global__ void k(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = 2, a2 = 3, a3 = 4, b0 = 5, b1 = 6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4 * 4);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  for (int kind = 0; kind < 2; ++kind) for (int warps : {4, 8, 16, 32}) {
    int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (kind == 0) k<0><<<148, warps * 32>>>(out, iters); else k<1><<<148, warps * 32>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mmas = 148.0 * warps * iters * 8;
    double flop = mmas * (kind == 0 ? 2.0 * 16 * 8 * 8 : 2.0 * 16 * 8 * 16);
    printf("%s warps/SM=%2d: %.3f ms  %.1f TFLOP/s  %.3f mma/clk/SM (at %.0f MHz)\n", kind == 0 ? "tf32 m16n8k8 " : "bf16 m16n8k16", warps, ms,
           flop / ms / 1e9, mmas / 148 / (ms * 1e-3 * p.clockRate * 1e3), p.clockRate / 1e3);
  }
  return 0;
}
