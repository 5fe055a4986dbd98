// Host-side interface of the tcgen05 GEMM kernels (implemented in gemm_tc.cu).
#pragma once
#include "common.cuh"

namespace tpc {

// Epilogue of   D[M,N] = epi( A[M,K] @ Bt[N,K]^T )   (both operands K-major, fp32 storage, TF32 MMA)
struct GemmEpi {
  const float* bias = nullptr;    // [N]
  const float* gamma = nullptr;   // LayerNorm scale [N]   (ln != 0, requires N == 128)
  const float* beta = nullptr;    // LayerNorm shift [N]
  float* rstd_out = nullptr;      // [M] 1/sqrt(var+eps) per row (saved for backward), optional
  const float* aux = nullptr;     // [M,N] (ld = ldaux) second input of the epilogue
  int ldaux = 0;
  int aux_mode = 0;               // 0 none | 1 v += aux (residual / grad accumulation) | 2 v = aux > 0 ? v : 0 (ReLU bwd)
  int relu = 0;                   // v = max(v, 0) after bias
  int ln = 0;                     // LayerNorm over the row after bias + aux
  int round_out = 1;              // store RN-rounded TF32 values (the output feeds a tensor-core operand)
  int atomic = 0;                 // split-K: atomically accumulate into D (D pre-initialised), no other epilogue op
  float ln_eps = 1e-6f;
  // ReLU mask as a bit image (weight-resident kernel only, see gemm_weight_resident()), laid out
  // [N / 128 column tiles][mask_ld rows][4 words], mask_ld = M rounded up to a multiple of 128 (so the words one CTA
  // needs for a row tile are 2 KB contiguous): bit j of word c = output column 128 n_t + 32 c + j. mask_out is written
  // by a split-mode relu epilogue (bit = output positive), mask_in gates the output of a single-pass GEMM
  // (v = bit ? v : 0): the ReLU backward then reads 16 bytes per row tile instead of 512 bytes of saved activation.
  uint32_t* mask_out = nullptr;
  const uint32_t* mask_in = nullptr;
  int mask_ld = 0;
  // LayerNorm BACKWARD fused into the epilogue (generic kernel, N == 128, single pass): the GEMM result (+ aux) is the
  // gradient w.r.t. a LayerNorm output y = LN(pre); instead of storing it, the epilogue stores the gradient w.r.t. pre,
  //   g = dy * gamma, xhat = (y - beta) / gamma, dpre = rstd * (g - mean(g) - xhat * mean(g * xhat))
  // and accumulates dgamma += sum_rows(dy * xhat), dbeta += sum_rows(dy) (kernels.cuh: ln_bwd, which this replaces
  // where the producer of dy is such a GEMM: dy is then never written to or re-read from HBM).
  // lnb_y [M,128] = the saved LayerNorm output, lnb_rstd [M]; gamma / beta above are the LayerNorm parameters.
  const float* lnb_y = nullptr;
  const float* lnb_rstd = nullptr;
  float* lnb_dgamma = nullptr;
  float* lnb_dbeta = nullptr;
};

// true when launch_gemm routes this shape to the weight-resident kernel (the one that supports mask_out / mask_in)
bool gemm_weight_resident(int M, int N, int K, const GemmEpi& epi, bool split, int num_sms);

// Caches CUtensorMap objects per (pointer, shape, box); owned by the library handle.
struct TmapCache;
TmapCache* tmap_cache_create();
void tmap_cache_destroy(TmapCache*);

int gemm_tc_init();  // resolves cuTensorMapEncodeTiled, sets kernel attributes. idempotent.

// D[M,N] = epi(A[M,K] @ Bt[N,K]^T). K % 32 == 0, N % 128 == 0. k_splits > 1 requires epi.atomic.
// Bt_corr != nullptr selects the fp32-accurate split product:
//   D = A_h.B_h + A_h.B_r + A_r.B_h on kind::f16 MMAs with fp32 accumulation, x_h = fp16(x), x_r = fp16(x - x_h).
// Bt_corr is the fp16 split image [2][N][ldb] = {B_h, B_r} made by build_split_image (same byte size as an fp32
// [N][ldb] matrix; Bt itself is then only used for its shape); the activation halves are split on chip. Error ~2^-21
// relative; values beyond +-65504 saturate (fp16 range), parts below 6e-5 keep an absolute error <= 3e-8.
int launch_gemm(TmapCache* tc, const float* A, int lda, const float* Bt, int ldb, float* D, int ldd, int M, int N,
                int K, const GemmEpi& epi, int k_splits, int num_sms, cudaStream_t stream,
                const void* Bt_corr = nullptr);
int build_split_image(const float* Bt, int ldb, int N, int K, void* image, cudaStream_t stream);

// dW[nj][Kin, 128] += X[M,Kin]^T @ dY[M, nj*128 .. nj*128+127]  for every 128-wide column tile nj of dY.
// dptr[nj] is the destination of tile nj with leading dimension ldd[nj]; accumulation is atomic.
struct WgradDst {
  float* ptr[4] = {nullptr, nullptr, nullptr, nullptr};
  int ld[4] = {0, 0, 0, 0};
  float* bias[4] = {nullptr, nullptr, nullptr, nullptr};  // optional: db[128] += colsum(dY tile nj) (fused, same pass)
};
int launch_wgrad(TmapCache* tc, const float* X, int ldx, const float* dY, int ldy, int M, int Kin, int Nout,
                 const WgradDst& dst, int num_sms, cudaStream_t stream);

}  // namespace tpc
