// Phase timing of the attention kernels: clock64 stamps taken by thread 0 of every CTA (ATTN_PROFILE build of
// csrc/attention.cu), averaged over the grid. Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DATTN_PROFILE -I include \
//        -I transformer-pointer-critic_b200/csrc tests/gpu_checks/micro/attn_phase.cu -o tests/gpu_checks/micro/attn_phase
#include "attention.cu"
#include <cstdio>
#include <cstdlib>
#include <vector>
unsigned long long g_tpc_launches = 0;

int main(int argc, char** argv) {
  const int C = 2048, L = argc > 1 ? atoi(argv[1]) : 151;
  const float frac = argc > 2 ? atof(argv[2]) : 0.33f;
  std::vector<float> hq((size_t)C * L * 384), hm((size_t)C * L, 0.f), hg((size_t)C * L * 128);
  srand(1);
  auto rnd = []() { float s = 0; for (int i = 0; i < 12; ++i) s += rand() / (float)RAND_MAX; return s - 6.f; };
  for (auto& x : hq) x = rnd();
  for (auto& x : hg) x = rnd();
  const int nm = (int)(L * frac);
  for (int c = 0; c < C; ++c) for (int j = L - nm; j < L; ++j) hm[(size_t)c * L + j] = 1.f;
  float *qkv, *mask, *att, *lse, *datt, *dqkv;
  cudaMalloc(&qkv, hq.size() * 4); cudaMalloc(&mask, hm.size() * 4); cudaMalloc(&att, (size_t)C * L * 128 * 4);
  cudaMalloc(&lse, (size_t)C * L * 8 * 4); cudaMalloc(&datt, hg.size() * 4); cudaMalloc(&dqkv, hq.size() * 4);
  cudaMemcpy(qkv, hq.data(), hq.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(mask, hm.data(), hm.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(datt, hg.data(), hg.size() * 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<long long> st((size_t)8 * 32768);
  for (int which = 0; which < 2; ++which) {
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      int rc = which == 0 ? tpc::attention_fwd(qkv, mask, L, 0, C, L, att, lse, 0)
                          : tpc::attention_bwd(qkv, att, lse, datt, mask, L, 0, C, L, dqkv, 0);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      if (rc) { printf("rc %d\n", rc); return 1; }
      cudaEventElapsedTime(&ms, e0, e1);
    }
    cudaMemcpyFromSymbol(st.data(), tpc::attn_prof, st.size() * 8);
    const int n = 296;   // persistent grids: every launch has at least 2 CTAs per SM
    double acc[8] = {0};
    for (int b = 0; b < n; ++b) for (int i = 1; i < 8; ++i) acc[i] += (double)(st[b * 8 + i] - st[b * 8 + i - 1]);
    printf("%s L=%d masked=%.2f: %.1f us; mean clocks between stamps:", which ? "bwd" : "fwd", L, frac, ms * 1e3);
    for (int i = 1; i < 8; ++i) printf(" [%d-%d] %.0f", i - 1, i, acc[i] / n);
    printf("\n");
  }
  return 0;
}
