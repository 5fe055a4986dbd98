// C-ABI entry points for the stand-alone operators (unit-test granularity): tensor-core GEMMs.
// Declared in include/tpc.h.
#include "handle.cuh"
#include "kernels.cuh"

using namespace tpc;

unsigned long long g_tpc_launches = 0;

extern "C" {

unsigned long long tpc_launch_count(void) { return g_tpc_launches; }

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
// bit3 atomic split-K; aux_mode 0/1/2 (none / add / relu-gate). Bt_lo != NULL: 3xTF32 (Bt raw fp32, Bt_lo = Bt -
// trunc_tf32(Bt)), fp32-accurate result.
int tpc_gemm(tpc_handle* h, const float* A, int lda, const float* Bt, int ldb, float* D, int ldd, int M, int N, int K,
             const float* bias, const float* aux, int ldaux, int aux_mode, const float* gamma, const float* beta,
             float* rstd_out, int flags, int k_splits, const float* Bt_lo, void* stream) {
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
  return launch_gemm(h->tmaps, A, lda, Bt, ldb, D, ldd, M, N, K, e, k_splits, h->num_sms, (cudaStream_t)stream, Bt_lo);
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

}  // extern "C"
