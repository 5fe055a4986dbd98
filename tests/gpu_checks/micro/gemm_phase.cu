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
  if (tpc::gemm_tc_init()) { printf("init failed\n"); return 1; }
  tpc::TmapCache* tc = tpc::tmap_cache_create();
  float *A, *W, *Wlo, *res, *D, *g, *b, *rstd;
  cudaMalloc(&A, (size_t)M * K * 4); cudaMalloc(&W, 128 * K * 4); cudaMalloc(&Wlo, 128 * K * 4);
  cudaMalloc(&res, (size_t)M * 128 * 4); cudaMalloc(&D, (size_t)M * 128 * 4);
  cudaMalloc(&g, 512); cudaMalloc(&b, 512); cudaMalloc(&rstd, (size_t)M * 4);
  cudaMemset(A, 0, (size_t)M * K * 4); cudaMemset(W, 0, 128 * K * 4); cudaMemset(Wlo, 0, 128 * K * 4);
  cudaMemset(res, 0, (size_t)M * 128 * 4); cudaMemset(g, 0, 512); cudaMemset(b, 0, 512);
  tpc::GemmEpi e;
  e.bias = b;
  if (ln) { e.gamma = g; e.beta = b; e.rstd_out = rstd; e.ln = 1; }
  e.aux = res; e.ldaux = 128; e.aux_mode = 1; e.round_out = 0;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 4; ++rep) {
    long long z[16] = {0};
    cudaMemcpyToSymbol(tpc::gemm_prof, z, sizeof(z));
    cudaEventRecord(e0);
    int rc = tpc::launch_gemm(tc, A, K, W, K, D, 128, M, 128, K, e, 1, 148, 0, split ? Wlo : nullptr);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    if (rc) { printf("rc %d\n", rc); return 1; }
    cudaEventElapsedTime(&ms, e0, e1);
  }
  long long pr[16];
  cudaMemcpyFromSymbol(pr, tpc::gemm_prof, sizeof(pr));
  const double kbs = (double)((M + 127) / 128 + 147) / 148 * (K / 32);   // k-blocks of CTA 0
  printf("M=%d K=%d split=%d ln=%d: %.1f us, %.0f k-blocks per CTA -> %.0f clk per k-block at 1.9 GHz\n", M, K, split, ln, ms * 1e3, kbs,
         ms * 1e-3 * 1.9e9 / kbs);
  const char* nm[9] = {"producer: wait empty", "splitter: wait full", "splitter: lds+alu", "splitter: tmem st+wait",
                       "splitter: fence+arrive", "mma: wait full", "mma: wait lofull", "mma: issue+commit", "mma: wait tempty (per tile)"};
  for (int i = 0; i < 9; ++i) printf("  %-30s %8.0f clk per k-block\n", nm[i], pr[i] / kbs);
  return 0;
}
