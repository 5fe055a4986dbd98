// Micro-benchmark: how far can one thread run ahead of the tensor pipe? Time to ISSUE n back-to-back tcgen05.mma
// (TF32 128x128x8, 64 clocks each) from an idle pipe, and the cost of an already-satisfied mbarrier wait.
#include "common.cuh"
#include <cstdio>
using namespace tpc;
unsigned long long g_tpc_launches = 0;

__global__ void __launch_bounds__(128, 1) k(long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, done;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 2 * 16384 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = i < 1024 ? __uint_as_float((i * 7 + 1) & 1023) : 1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&done, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_tf32(128, 128, 0, 0);
    const uint64_t da = umma_smem_desc_sw128(smem_u32(smem), 16, 1024), db = umma_smem_desc_sw128(smem_u32(smem + 16384), 16, 1024);
    uint32_t ph = 0;
    for (int n = 1; n <= 32; ++n) {
      const long long t0 = clock64();
      for (int i = 0; i < n; ++i) umma_tf32_ss(tm, da + 2 * (i & 3), db + 2 * (i & 3), idesc, 1u);
      const long long t1 = clock64();
      umma_commit(&done);
      mbar_wait(&done, ph);
      ph ^= 1;
      const long long t2 = clock64();
      out[2 * n] = t1 - t0;
      out[2 * n + 1] = t2 - t0;
    }
    // satisfied wait: the barrier completed long ago
    mbar_arrive(&bar);
    for (int i = 0; i < 2000; ++i) __nanosleep(20);
    long long t0 = clock64();
    for (int i = 0; i < 64; ++i) mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[0] = (t1 - t0) / 64;
    t0 = clock64();
    for (int i = 0; i < 64; ++i) tc_fence_after();
    t1 = clock64();
    out[1] = (t1 - t0) / 64;
    // variants of the satisfied wait, each followed by a branch on its result (as in a spin loop)
    uint32_t okv = 0;
#define WAIT_VARIANT(slot, INSTR)                                                                              \
    t0 = clock64();                                                                                            \
    for (int i = 0; i < 64; ++i) {                                                                             \
      uint32_t ok;                                                                                             \
      do {                                                                                                     \
        asm volatile("{\n\t.reg .pred p;\n\t" INSTR " p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"               \
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");                                    \
      } while (!ok);                                                                                           \
      okv += ok;                                                                                               \
    }                                                                                                          \
    t1 = clock64();                                                                                            \
    out[slot] = (t1 - t0) / 64;
    WAIT_VARIANT(66, "mbarrier.test_wait.parity.shared::cta.b64")
    WAIT_VARIANT(67, "mbarrier.try_wait.parity.relaxed.cta.shared::cta.b64")
    WAIT_VARIANT(68, "mbarrier.try_wait.parity.shared::cta.b64")
    t0 = clock64();
    for (int i = 0; i < 64; ++i) { while (!mbar_try_wait(&bar, 0)) {} }
    t1 = clock64();
    out[70] = (t1 - t0) / 64;
    t0 = clock64();
    for (int i = 0; i < 64; ++i) {
      uint32_t spins = 0;
      while (!mbar_try_wait(&bar, 0)) { if (++spins > (1u << 26)) { out[79] = 1; break; } }
    }
    t1 = clock64();
    out[71] = (t1 - t0) / 64;
    // dependent shared-memory load latency for scale
    volatile uint32_t* sp = reinterpret_cast<volatile uint32_t*>(smem);
    uint32_t idx = 0;
    t0 = clock64();
    for (int i = 0; i < 64; ++i) idx = sp[idx & 1023];
    t1 = clock64();
    out[69] = (t1 - t0) / 64 + (idx + okv) * 0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}
int main() {
  long long* d; cudaMalloc(&d, 80 * 8); cudaMemset(d, 0, 80 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 16384 + 2048);
  k<<<1, 128, 2 * 16384 + 2048>>>(d);
  cudaDeviceSynchronize();
  long long h[80]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("satisfied mbarrier wait: %lld clk; tcgen05.fence::after_thread_sync: %lld clk  [%s]\n", h[0], h[1], cudaGetErrorString(cudaGetLastError()));
  printf("test_wait %lld clk; try_wait.relaxed.cta %lld clk; try_wait (default acquire.cta) %lld clk; dependent LDS %lld clk\n", h[66], h[67], h[68], h[69]);
  printf("lean spin loop %lld clk; bounded loop that flags and breaks %lld clk\n", h[70], h[71]);
  for (int n = 1; n <= 32; n += 4) printf("n=%2d MMAs: issue returns after %5lld clk, all complete (commit + wait) after %5lld clk\n", n, h[2 * n], h[2 * n + 1]);
  return 0;
}
