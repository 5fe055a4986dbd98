// C-ABI entry points for the stand-alone operators (unit-test granularity): tensor-core GEMMs.
// Declared in include/tpc.h.
#include "handle.cuh"
#include "kernels.cuh"

using namespace tpc;

unsigned long long g_tpc_launches = 0;

namespace {
// per-episode return = sum over the T step rewards; out = {min, max, mean} over the batch (agents/trainer.py:89-100).
// One CTA: the batch is a few thousand episodes at most.
__global__ void __launch_bounds__(256) reward_stats_kernel(const float* __restrict__ rewards, int B, int T,
                                                           float* __restrict__ out) {
  __shared__ float smn[8], smx[8], ssum[8];
  float mn = INFINITY, mx = -INFINITY, sum = 0.f;
  for (int b = threadIdx.x; b < B; b += 256) {
    float r = 0.f;
    for (int t = 0; t < T; ++t) r += rewards[(size_t)b * T + t];
    mn = fminf(mn, r);
    mx = fmaxf(mx, r);
    sum += r;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
  }
  if ((threadIdx.x & 31) == 0) smn[threadIdx.x >> 5] = mn, smx[threadIdx.x >> 5] = mx, ssum[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) mn = fminf(mn, smn[w]), mx = fmaxf(mx, smx[w]), sum += ssum[w];
    out[0] = mn;
    out[1] = mx;
    out[2] = sum / (float)B;
  }
}

__global__ void scale_kernel(const float* __restrict__ in, int n, float scale, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * scale;
}

// tcgen05.mma kind::tf32 issue rate on operands resident in shared memory: the ceiling of every projection GEMM
// (one CTA per SM, thread 0 issues, 3 operand stages walked like the GEMM kernels do, accumulator in TMEM).
__global__ void __launch_bounds__(128, 1) tf32_peak_kernel(int iters) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (3 * 16384 + 3 * 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_tf32(128, 256, 0, 0);
    for (int it = 0; it < iters; ++it) {
      const int st = it % 3;
      const uint64_t da = umma_smem_desc_sw128(smem_u32(smem + st * 16384), 16, 1024);
      const uint64_t db = umma_smem_desc_sw128(smem_u32(smem + 3 * 16384 + st * 32768), 16, 1024);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_tf32_ss(tm, da + 2 * kk, db + 2 * kk, idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}
}  // namespace

extern "C" {

unsigned long long tpc_launch_count(void) { return g_tpc_launches; }

int tpc_zero(tpc_handle* h, void* ptr, size_t nbytes, void* stream) {
  if (!h || (!ptr && nbytes)) return TPC_ERR_ARG;
  if (nbytes) TPC_CHECK_CUDA(cudaMemsetAsync(ptr, 0, nbytes, (cudaStream_t)stream));
  return TPC_OK;
}

int tpc_scale(tpc_handle* h, const float* in, int n, float scale, float* out, void* stream) {
  if (!h || !in || !out || n < 0) return TPC_ERR_ARG;
  if (n == 0) return TPC_OK;
  scale_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(in, n, scale, out);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int tpc_tf32_peak_probe(tpc_handle* h, int iters, float* tflops_host, void* stream) {
  if (!h || !tflops_host || iters < 1) return TPC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int smem = 3 * 16384 + 3 * 32768 + 2048;
  TPC_CHECK_CUDA(cudaFuncSetAttribute(tf32_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  TPC_CHECK_CUDA(cudaEventCreate(&e0));
  TPC_CHECK_CUDA(cudaEventCreate(&e1));
  tf32_peak_kernel<<<h->num_sms, 128, smem, st>>>(iters / 8 + 1);   // warm-up
  TPC_CHECK_CUDA(cudaEventRecord(e0, st));
  tf32_peak_kernel<<<h->num_sms, 128, smem, st>>>(iters);
  TPC_CHECK_CUDA(cudaEventRecord(e1, st));
  TPC_CHECK_CUDA(cudaEventSynchronize(e1));
  g_tpc_launches += 2;
  float ms = 0.f;
  TPC_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  TPC_CHECK_CUDA(cudaGetLastError());
  *tflops_host = (float)(2.0 * 128 * 256 * 8 * 4.0 * iters * h->num_sms / (ms * 1e-3) / 1e12);
  return TPC_OK;
}

int tpc_create(tpc_handle** out) {
  if (!out) return TPC_ERR_ARG;
  int dev = 0;
  TPC_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  TPC_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    fprintf(stderr, "[tpc] this library is built for sm_100a only (device is sm_%d%d)\n", prop.major, prop.minor);
    return TPC_ERR_DRIVER;
  }
  TPC_TRY(gemm_tc_init());
  tpc_handle* h = new tpc_handle();
  h->device = dev;
  h->num_sms = prop.multiProcessorCount;
  h->tmaps = tmap_cache_create();
  *out = h;
  return TPC_OK;
}

int tpc_destroy(tpc_handle* h) {
  if (!h) return TPC_OK;
  tmap_cache_destroy(h->tmaps);
  delete h;
  return TPC_OK;
}

int tpc_num_sms(const tpc_handle* h) { return h ? h->num_sms : 0; }

// D[M,N] = epi(A[M,K] @ Bt[N,K]^T + bias); flags: bit0 relu, bit1 layernorm(+rstd), bit2 round_out,
// bit3 atomic split-K; aux_mode 0/1/2 (none / add / relu-gate). Bt_corr != NULL: fp32-accurate split product (Bt raw
// fp32, Bt_corr = the bf16 correction image of tpc_gemm_split_image).
int tpc_gemm(tpc_handle* h, const float* A, int lda, const float* Bt, int ldb, float* D, int ldd, int M, int N, int K,
             const float* bias, const float* aux, int ldaux, int aux_mode, const float* gamma, const float* beta,
             float* rstd_out, int flags, int k_splits, const void* Bt_corr, void* stream) {
  if (!h) return TPC_ERR_ARG;
  GemmEpi e;
  e.bias = bias;
  e.aux = aux;
  e.ldaux = ldaux;
  e.aux_mode = aux_mode;
  e.gamma = gamma;
  e.beta = beta;
  e.rstd_out = rstd_out;
  e.relu = (flags & 1) ? 1 : 0;
  e.ln = (flags & 2) ? 1 : 0;
  e.round_out = (flags & 4) ? 1 : 0;
  e.atomic = (flags & 8) ? 1 : 0;
  return launch_gemm(h->tmaps, A, lda, Bt, ldb, D, ldd, M, N, K, e, k_splits, h->num_sms, (cudaStream_t)stream, Bt_corr);
}

int tpc_gemm_ln_bwd(tpc_handle* h, const float* A, int lda, const float* Bt, int ldb, float* D, int ldd, int M, int K,
                    const float* aux, int ldaux, const float* y, const float* rstd, const float* gamma,
                    const float* beta, float* dgamma, float* dbeta, void* stream) {
  if (!h || !A || !Bt || !D || !y || !rstd || !gamma || !beta || !dgamma || !dbeta) return TPC_ERR_ARG;
  GemmEpi e;
  if (aux) { e.aux = aux; e.ldaux = ldaux; e.aux_mode = 1; }
  e.gamma = gamma;
  e.beta = beta;
  e.lnb_y = y;
  e.lnb_rstd = rstd;
  e.lnb_dgamma = dgamma;
  e.lnb_dbeta = dbeta;
  return launch_gemm(h->tmaps, A, lda, Bt, ldb, D, ldd, M, 128, K, e, 1, h->num_sms, (cudaStream_t)stream, nullptr);
}

int tpc_gemm_split_image(tpc_handle* h, const float* Bt, int ldb, int N, int K, void* image, void* stream) {
  if (!h) return TPC_ERR_ARG;
  return build_split_image(Bt, ldb, N, K, image, (cudaStream_t)stream);
}

// dW[Kin,Nout] += X[M,Kin]^T @ dY[M,Nout]   (dW row-major with leading dimension ldw, atomically accumulated);
// db[Nout] += colsum(dY) when db != NULL (fused into the same pass)
int tpc_wgrad(tpc_handle* h, const float* X, int ldx, const float* dY, int ldy, float* dW, int ldw, float* db, int M,
              int Kin, int Nout, void* stream) {
  if (!h) return TPC_ERR_ARG;
  WgradDst d;
  const int nj = ceil_div(Nout, 128);
  if (nj > 4) return TPC_ERR_SHAPE;
  for (int j = 0; j < nj; ++j) {
    d.ptr[j] = dW + j * 128;
    d.ld[j] = ldw;
    d.bias[j] = db ? db + j * 128 : nullptr;
  }
  return launch_wgrad(h->tmaps, X, ldx, dY, ldy, M, Kin, Nout, d, h->num_sms, (cudaStream_t)stream);
}

// Self-attention of one encoder stack on packed qkv[C*L][384] (q | k | v); mask[c*mS + m0 + j] != 0 hides key j.
// att[C*L][128], lse[C*L][8]. Replaces MultiHeadAttention.call's scaled_dot_product_attention + head merge
// (agents/models/transformer/common/attention.py:37-48, common/utils.py:13-47).
int tpc_attention_fwd(tpc_handle* h, const float* qkv, const float* mask, int mS, int m0, int C, int L, float* att,
                      float* lse, void* stream) {
  if (!h || !qkv || !mask || !att) return TPC_ERR_ARG;
  return attention_fwd(qkv, mask, mS, m0, C, L, att, lse, (cudaStream_t)stream);
}
// dqkv[C*L][384] from datt[C*L][128] (backward of the same ops, agents/trainer.py:115-141)
int tpc_attention_bwd(tpc_handle* h, const float* qkv, const float* att, const float* lse, const float* datt,
                      const float* mask, int mS, int m0, int C, int L, float* dqkv, void* stream) {
  if (!h || !qkv || !att || !lse || !datt || !mask || !dqkv) return TPC_ERR_ARG;
  return attention_bwd(qkv, att, lse, datt, mask, mS, m0, C, L, dqkv, (cudaStream_t)stream);
}


// ---- step-wise loss arithmetic (the fused update does the same inside tpc_a2c_grads) --------------------------------
// replaces Agent.compute_discounted_rewards (agents/agent.py:110-124): returns (B,T), bootstrap (B) may be NULL (= 0)
int tpc_discounted_returns(tpc_handle* h, const float* rewards, const float* bootstrap, int B, int T, float gamma,
                           float* returns, void* stream) {
  if (!h || !rewards || !returns) return TPC_ERR_ARG;
  if (B < 0 || T < 0) return TPC_ERR_SHAPE;
  return discounted_returns(rewards, B, T, gamma, returns, (cudaStream_t)stream, bootstrap, 1);
}
// replaces the arithmetic of Agent.compute_value_loss (agents/agent.py:155-164): advantages = R - V,
// loss[0] = c_v * mean((R - V)^2)
int tpc_value_loss(tpc_handle* h, const float* returns_v, const float* values, int n, float c_v, float* advantages,
                   float* loss, void* stream) {
  if (!h || !returns_v || !values || !advantages || !loss) return TPC_ERR_ARG;
  if (n < 0) return TPC_ERR_SHAPE;
  TPC_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream));
  return critic_loss_bwd(returns_v, values, n, c_v, 0.f, advantages, nullptr, loss, (cudaStream_t)stream,
                         n > 0 ? 1.f / (float)n : 0.f);
}
// replaces the arithmetic of Agent.compute_actor_loss (agents/agent.py:189-240): per-state sparse cross-entropy and
// entropy, sums[0..2] = mean policy loss, mean(policy_loss * advantage - c_e * entropy), mean entropy
int tpc_actor_loss(tpc_handle* h, const float* logits, const int32_t* actions, const float* advantages, int n, int S,
                   float c_e, float* policy_loss, float* entropy, float* sums, void* stream) {
  if (!h || !logits || !actions || !advantages || !sums) return TPC_ERR_ARG;
  if (n < 0 || S < 1) return TPC_ERR_SHAPE;
  TPC_CHECK_CUDA(cudaMemsetAsync(sums, 0, 3 * sizeof(float), (cudaStream_t)stream));
  return actor_loss_bwd(logits, actions, advantages, n, S, c_e, 0.f, nullptr, sums - 1, (cudaStream_t)stream,
                        n > 0 ? 1.f / (float)n : 0.f, policy_loss, entropy);
}

// replaces the batch statistics of the training loop (agents/trainer.py:89-100): out[3] = min, max, mean over the
// batch of the per-episode returns (sum of the T step rewards)
int tpc_episode_reward_stats(tpc_handle* h, const float* rewards, int B, int T, float* out, void* stream) {
  if (!h || !rewards || !out) return TPC_ERR_ARG;
  if (B < 1 || T < 1) return TPC_ERR_SHAPE;
  reward_stats_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(rewards, B, T, out);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}
}  // extern "C"
