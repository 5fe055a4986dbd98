// Where the K = 512 LayerNorm GEMM (the roofline launch of bench.py) spends its time: clock64 accumulators in the
// producer / splitter / MMA-issuer roles of CTA 0 (GEMM_PROFILE build of csrc/gemm_tc.cu). Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DGEMM_PROFILE -I include \
//        -I transformer-pointer-critic_b200/csrc tests/gpu_checks/micro/gemm_phase.cu -lcuda -o tests/gpu_checks/micro/gemm_phase
#include "gemm_tc.cu"
#include <cstdio>
#include <cstdlib>
#include <vector>
unsigned long long g_tpc_launches = 0;

int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 10240 * 151, K = argc > 2 ? atoi(argv[2]) : 512;
  const int split = argc > 3 ? atoi(argv[3]) : 1, ln = argc > 4 ? atoi(argv[4]) : 1;
  const int N = argc > 5 ? atoi(argv[5]) : 128;   // N > 128 or K == 128 without LayerNorm: the weight-resident kernel
  if (tpc::gemm_tc_init()) { printf("init failed\n"); return 1; }
  tpc::TmapCache* tc = tpc::tmap_cache_create();
  float *A, *W, *Wlo, *res, *D, *g, *b, *rstd;
  cudaMalloc(&A, (size_t)M * K * 4); cudaMalloc(&W, (size_t)N * K * 4); cudaMalloc(&Wlo, (size_t)N * K * 4);
  cudaMalloc(&res, (size_t)M * 128 * 4); cudaMalloc(&D, (size_t)M * N * 4);
  cudaMalloc(&g, 4 * N); cudaMalloc(&b, 4 * N); cudaMalloc(&rstd, (size_t)M * 4);
  cudaMemset(A, 0, (size_t)M * K * 4); cudaMemset(W, 0, (size_t)N * K * 4); cudaMemset(Wlo, 0, (size_t)N * K * 4);
  cudaMemset(res, 0, (size_t)M * 128 * 4); cudaMemset(g, 0, 4 * N); cudaMemset(b, 0, 4 * N);
  if (argc > 7 && atoi(argv[7])) {   // random operands instead of zeros (data-dependent power -> SM clock)
    std::vector<float> h((size_t)M * K);
    unsigned x = 12345u;
    for (auto& v : h) { x = x * 1664525u + 1013904223u; v = ((int)(x >> 8) - (1 << 23)) * (1.0f / (1 << 23)); }
    cudaMemcpy(A, h.data(), (size_t)M * K * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(W, h.data(), (size_t)N * K * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(Wlo, h.data() + 7, (size_t)N * K * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(res, h.data() + 11, (size_t)M * 128 * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(g, h.data() + 3, 4 * N, cudaMemcpyHostToDevice);
  }
  tpc::GemmEpi e;
  e.bias = b;
  if (ln) { e.gamma = g; e.beta = b; e.rstd_out = rstd; e.ln = 1; }
  if (ln) { e.aux = res; e.ldaux = 128; e.aux_mode = 1; }
  e.round_out = 0;
  if (!ln) e.relu = 1;
  const int maskmode = argc > 6 ? atoi(argv[6]) : 0;   // 1: relu epilogue writes the bit image; 2: gate by the bit image; 3: gate by aux (in place)
  uint32_t* mask = nullptr;
  const int Mpad = (M + 127) / 128 * 128;
  if (maskmode == 1 || maskmode == 2) { cudaMalloc(&mask, (size_t)Mpad * (N / 32) * 4); cudaMemset(mask, 0xa5, (size_t)Mpad * (N / 32) * 4); }
  if (maskmode == 1) { e.mask_out = mask; e.mask_ld = Mpad; }
  if (maskmode == 2) { e.relu = 0; e.bias = nullptr; e.mask_in = mask; e.mask_ld = Mpad; }
  if (maskmode == 3) { e.relu = 0; e.bias = nullptr; e.aux = D; e.ldaux = N; e.aux_mode = 2; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < (argc > 8 ? atoi(argv[8]) : 4); ++rep) {
#ifdef GEMM_PROFILE
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(tpc::gemm_prof, z, sizeof(z));
#endif
    cudaEventRecord(e0);
    int rc = tpc::launch_gemm(tc, A, K, W, K, D, N, M, N, K, e, 1, 148, 0, split ? Wlo : nullptr);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    if (rc) { printf("rc %d\n", rc); return 1; }
    cudaEventElapsedTime(&ms, e0, e1);
  }
  unsigned long long pr[16] = {0};
#ifdef GEMM_PROFILE
  cudaMemcpyFromSymbol(pr, tpc::gemm_prof, sizeof(pr));
#endif
  const double kbs = (double)((M + 127) / 128) * (N / 128) / 148 * (K / 32);   // k-blocks of CTA 0
  printf("M=%d K=%d split=%d ln=%d: %.1f us, %.0f k-blocks per CTA -> %.0f clk per k-block at 1.9 GHz\n", M, K, split, ln, ms * 1e3, kbs,
         ms * 1e-3 * 1.9e9 / kbs);
#ifdef GEMM_CLOCKS
  long long ck[2] = {0, 0};
  cudaMemcpyFromSymbol(ck, tpc::gemm_clk, sizeof(ck));
  if (ck[1] > 0) printf("  CTA 0 MMA issuer: %lld clk in %lld ns -> SM clock %.3f GHz, %.0f clk per k-block (measured clock)\n", ck[0], ck[1],
                        (double)ck[0] / ck[1], ck[0] / kbs);
#endif
  const char* nm[13] = {"producer: wait empty", "splitter: wait full", "splitter: lds+alu", "splitter: tmem st+wait",
                       "splitter: fence+arrive", "mma: wait full", "mma: wait lofull", "mma: issue+commit", "mma: wait tempty (per tile)",
                       "epilogue: wait tfull", "epilogue: chunk ld+math+sts", "epilogue: wait store read", "epilogue: barrier+store"};
  if (pr[14]) printf("  TMA issue -> data seen by a waiting splitter: %.0f clk average over %llu k-blocks\n", (double)pr[13] / pr[14], pr[14]);
  if (pr[14]) printf("  commit issued -> producer sees the stage empty: %.0f clk; TMA age when the MMA issuer reaches the stage: %.0f clk (bres kernels; epilogue slot 12 unused there)\n",
                     (double)pr[15] / pr[14], (double)pr[12] / pr[14]);
  for (int i = 0; i < 13; ++i) printf("  %-30s %8.0f clk per k-block\n", nm[i], pr[i] / kbs);
  return 0;
}
