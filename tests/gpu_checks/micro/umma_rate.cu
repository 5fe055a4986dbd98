// Micro-benchmark: tcgen05.mma kind::tf32 issue/execution rate of one CTA per SM on operands resident in shared
// memory (no TMA traffic): SS mode N = 128 / 256, with the operand walk of gemm_tc_kernel (4 k8 slices of a 128 B
// swizzle span, a ring of stages), TS mode (A from TMEM), and the 3xTF32 triple.
#include "common.cuh"
#include <cstdio>
using namespace tpc;
unsigned long long g_tpc_launches = 0;

template <int MODE, int N>
__global__ void __launch_bounds__(320, 1) k(int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t bars[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (6 * 16384 + 2 * 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (MODE >= 6 && warp >= 1) {
    // the other roles of the GEMM kernels: 9 warps spinning on an mbarrier that completes when the MMA stream ends
    if (MODE == 6) mbar_wait(&bars[3], 0);
    if (MODE == 7) {   // same wait with a suspend-time hint: the hardware parks the thread instead of re-probing
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "SPIN_LOOP:\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
          "@p bra SPIN_DONE;\n\t"
          "bra SPIN_LOOP;\n\t"
          "SPIN_DONE:\n\t}" ::"r"(smem_u32(&bars[3])), "r"(0u), "r"(1000000u) : "memory");
    }
  }
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_tf32(128, N, 0, 0);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int st = it % 3;
      const uint32_t sa = smem_u32(smem + st * 16384);
      const uint32_t sb = smem_u32(smem + 3 * 16384 + st * 32768);
      const uint64_t da = umma_smem_desc_sw128(sa, 16, 1024), db = umma_smem_desc_sw128(sb, 16, 1024);
      const uint64_t dbl = umma_smem_desc_sw128(sb + 16384, 16, 1024);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (MODE == 0) umma_tf32_ss(tm, da + 2 * kk, db + 2 * kk, idesc, 1u);
        if (MODE == 1) umma_tf32_ts(tm, tm + 256 + st * 32 + 8 * kk, db + 2 * kk, idesc, 1u);
        if (MODE == 2) {
          umma_tf32_ss(tm, da + 2 * kk, db + 2 * kk, idesc, 1u);
          umma_tf32_ss(tm, da + 2 * kk, dbl + 2 * kk, idesc, 1u);
          umma_tf32_ts(tm, tm + 256 + st * 32 + 8 * kk, db + 2 * kk, idesc, 1u);
        }
        if (MODE >= 3) {   // all-TS triple as issued by the GEMM kernels
          umma_tf32_ts(tm, tm + 384 + st * 32 + 8 * kk, db + 2 * kk, idesc, 1u);
          umma_tf32_ts(tm, tm + 384 + st * 32 + 8 * kk, dbl + 2 * kk, idesc, 1u);
          umma_tf32_ts(tm, tm + 256 + st * 32 + 8 * kk, db + 2 * kk, idesc, 1u);
        }
      }
      if (MODE >= 4) umma_commit(&bars[st]);          // per-k-block commit to a stage barrier (never waited on)
      if (MODE == 5) { mbar_try_wait(&bars[3], 1); tc_fence_after(); }   // the wait + fence of the next k-block
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (MODE >= 6) mbar_arrive(&bars[3]);
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}

template <int MODE, int N>
void run(const char* name, long long* d) {
  const int iters = 4000, smem = 6 * 16384 + 2 * 32768 + 2048;
  cudaFuncSetAttribute(k<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int grid : {1, 148}) {
    k<MODE, N><<<grid, MODE >= 6 ? 320 : 128, smem>>>(iters, d);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double n_mma = iters * 4.0 * (MODE >= 2 ? 3 : 1);
    printf("%-28s grid %3d: %7.1f clk per UMMA (M=128 N=%d K=8) -> %.0f TFLOP/s at 148 SMs x 1.9 GHz  [%s]\n", name, grid,
           h / n_mma, N, 2.0 * 128 * N * 8 / (h / n_mma) * 148 * 1.9e9 / 1e12, cudaGetErrorString(cudaGetLastError()));
  }
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  run<0, 128>("SS", d);
  run<0, 256>("SS", d);
  run<1, 128>("TS (A in TMEM)", d);
  run<2, 128>("3xTF32 triple SS,SS,TS", d);
  run<3, 128>("3xTF32 triple TS,TS,TS", d);
  run<4, 128>("  + commit per k-block", d);
  run<5, 128>("  + try_wait + fence", d);
  run<6, 128>("  + 9 warps spinning (lean)", d);
  run<7, 128>("  + 9 warps, suspend hint", d);
  return 0;
}
