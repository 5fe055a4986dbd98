// Launchers of the CUDA-core kernels around the tensor-core GEMMs (implemented in kernels_simt.cu,
// attention.cu, env.cu). All row-major fp32; "rows" of 128 features unless noted.
#pragma once
#include "common.cuh"

namespace tpc {

// Row map used where a compact per-stack layout meets the concatenated layout (common/encoder.py:91-93):
// compact row r of a stack with L tokens per state  <->  token (r / L) * S + s0 + (r % L) of the (.., S, ..) tensor.
struct RowMap { int L, S, s0; };

// Token window of the ACTOR's compact buffers. The first `gap` rule tokens of every state in the batch (rules placed
// at earlier steps) are masked as attention keys in every stack and as pointer targets, and their own encoder rows are
// only read by pointer logits that are masked anyway, so the actor leaves them out: exact. Compact token
// j of a state <-> token (j < N ? j : j + gap) of the full (.., Sfull, ..) tensors (state, masks, logits).
struct TokWin {
  int N = 1 << 30, gap = 0, Sfull = 0;
  __host__ __device__ int full(int j) const { return j < N ? j : j + gap; }
};

// out[r][0:128] = x[tok(r)][0:F] @ W[F][128] + b ; x has ldx floats per token; feature f == ldx is read from
// `extra[tok]` when extra != nullptr (is_empty column of the "cost" reward state, env.py:424-426).
// Restates the Dense / TimeDistributed(Dense) embeddings (common/encoder.py:101-121, actor/decoder.py:74).
int small_dense_fwd(const float* x, int ldx, const float* extra, int rows, RowMap map, const float* W, const float* b,
                    int F, float* out, cudaStream_t st);
// dW[F][128] += x^T dOut, db[128] += colsum(dOut)
int small_dense_bwd(const float* x, int ldx, const float* extra, int rows, RowMap map, const float* dOut, int F,
                    float* dW, float* db, cudaStream_t st);

// LayerNorm backward (tf.keras LayerNormalization, eps inside rstd): dpre = rstd * (g - mean(g) - xhat * mean(g*xhat)),
// g = dout * gamma, xhat = (out - beta) / gamma. dout rows are addressed through `dmap` (identity when L == S).
// dgamma/dbeta are accumulated atomically.
int ln_bwd(const float* dout, RowMap dmap, const float* out, const float* rstd, const float* gamma, const float* beta,
           int rows, float* dpre, float* dgamma, float* dbeta, cudaStream_t st);

// db[N] += colsum(dY[M][N])
int colsum(const float* dY, int ld, int M, int N, float* db, cudaStream_t st);

// dst[tok(r)][0:128] = src[r][0:128]   (concatenate_layer, common/encoder.py:91-93)
int scatter_rows(const float* src, int rows, RowMap map, float* dst, cudaStream_t st);

// Self-attention of one encoder stack (common/attention.py:37-48 + common/utils.py:31-45) on packed
// qkv[rows][384] (q | k | v), rows = C * L; mask[(c * mS) + m0 + j] != 0 hides key j.
// att[rows][128] (heads merged), lse[rows][8] = log-sum-exp of the scaled masked logits (saved for backward).
// mask_split / mask_gap: key j reads mask[c * mS + m0 + (j < mask_split ? j : j + mask_gap)] (TokWin of the joint stack)
int attention_fwd(const float* qkv, const float* mask, int mS, int m0, int C, int L, float* att, float* lse,
                  cudaStream_t st, int mask_split = 1 << 30, int mask_gap = 0);
// dqkv[rows][384] from datt[rows][128]
int attention_bwd(const float* qkv, const float* att, const float* lse, const float* datt, const float* mask, int mS,
                  int m0, int C, int L, float* dqkv, cudaStream_t st, int mask_split = 1 << 30, int mask_gap = 0);

// Backward of the decoder cross attention with ONE query per state (actor/decoder_layer.py:26-27; the forward is
// part of decoder_chain_fwd): q2[C][128], kvw[C*S][384] (K2 | V2 | W2e), probs pc[C][8][S] ->
// dq2[C][128], dkvw[:, 0:256]
int cross_attn_bwd(const float* dctx, const float* q2, const float* kvw, const float* pc, int C, int S, float* dq2,
                   float* dkvw, cudaStream_t st);

// One decoder layer on the single decoder token of every state (actor/decoder.py:74-81, actor/decoder_layer.py:18-32) as
// ONE kernel, 2 (rollout) or 8 (update) states per CTA, fp32 FMA on CUDA cores (the products are GEMV-sized: a
// 148-CTA tensor-core launch each would cost more in launch latency than in arithmetic):
//   [y0 = Dense_emb(dec_input)]                          first layer only (y_in == nullptr)
//   v1 = Dense_v1(y); o1 = LN1(Dense_o1(v1) + y)          mha1 over one key: softmax == 1
//   q2 = Dense_q2(o1); ctx = softmax(q2 K2^T / 4 + mask * -1e9) V2 per head over the S compact tokens of kvw
//   o2 = LN2(Dense_o2(ctx) + o1); f = relu(Dense_f1(o2)); o3 = LN3(Dense_f2(f) + o2)
//   [w1d = Dense_W1(o3)]                                 last layer only (Wp1 != nullptr): pointer_attention.py:43
// Every intermediate is also written to global memory: the backward pass reads them (agents/trainer.py:125-141).
// Weights are the canonical Keras tensors W[in][out], b[out]; dff a multiple of 128, at most 512.
struct DecChainArgs {
  const float* dec_input; int F;                 // [C][F]
  const float* y_in;                             // [C][128] input of a layer > 0, nullptr for the first one
  const float *Wemb, *bemb;
  const float *Wv1, *bv1, *Wo1, *bo1, *g1, *be1;
  const float *Wq2, *bq2, *Wo2, *bo2, *g2, *be2;
  const float *Wf1, *bf1, *Wf2, *bf2, *g3, *be3; int dff;
  const float *Wp1, *bp1;
  const float* kvw;                              // [C * S][384]: K2 | V2 | W2e of the compact tokens
  const float* mask; int S; TokWin win;          // mha mask (C, win.Sfull or S)
  float *y0, *v1, *o1, *rstd1, *q2, *pc, *ctx, *o2, *rstd2, *f, *o3, *rstd3, *w1d;
  int C; float ln_eps;
};
int decoder_chain_fwd(const DecChainArgs& a, cudaStream_t st);

// Pointer head (actor/pointer_attention.py:40-66): w1d[C][128] = W1 dec + b, W2e = kvw[:, 256:384];
// score (pre-clip), logits (= C*tanh(score) - mask*1e9), probs = softmax(logits), argmax(probs).
int pointer_fwd(const float* w1d, const float* kvw, const float* V, const float* bV, float clipC, int has_clip,
                const float* ptr_mask, int C, int S, float* score, float* logits, float* probs, int32_t* argmax,
                cudaStream_t st, TokWin w = TokWin());
// dlogits[C][S] -> dw1d[C][128], dkvw[:, 256:384], dV[128] += , dbV +=
int pointer_bwd(const float* dlogits, const float* score, const float* w1d, const float* kvw, const float* V,
                float clipC, int has_clip, int C, int S, float* dw1d, float* dkvw, float* dV, float* dbV,
                cudaStream_t st, TokWin w = TokWin());

// value head tail (critic/model.py:81-83): v[c] = h1[c] . Wv + bv
int value_head_fwd(const float* h1, const float* Wv, const float* bv, int C, float* v, cudaStream_t st);
int value_head_bwd(const float* dv, const float* h1, const float* Wv, int C, float* dh1, float* dWv, float* dbv,
                   cudaStream_t st);

// out[r][0:128] = bias[0:128]
int fill_rows_bias(float* out, const float* bias, int rows, cudaStream_t st);
int round_tf32_inplace(float* x, size_t n, cudaStream_t st);

// Returns (agents/agent.py:110-124): rewards[B][T] -> R[t*B + b] (horizontal: R[b*T + t]); bootstrap[B] = value after
// the last step (NULL = 0, agents/trainer.py:103-105)
int discounted_returns(const float* rewards, int B, int T, float gamma, float* R, cudaStream_t st,
                       const float* bootstrap = nullptr, int horizontal = 0);
// critic loss pieces (agents/agent.py:155-164): adv = R - V, dv = -2 c_v (R - V) * inv_total (optional),
// losses[0] += out_scale * sum c_v (R-V)^2
int critic_loss_bwd(const float* R, const float* V, int n, float c_v, float inv_total, float* adv, float* dv,
                    float* losses, cudaStream_t st, float out_scale = 1.f);
// actor loss pieces (agents/agent.py:189-240): dlogits = adv (softmax - onehot) * inv_total (optional); per-state
// sparse cross-entropy / entropy (optional); losses[1..3] += out_scale * sums of {policy loss, total, entropy}
int actor_loss_bwd(const float* logits, const int32_t* actions, const float* adv, int n, int S, float c_e,
                   float inv_total, float* dlogits, float* losses, cudaStream_t st, float out_scale = 1.f,
                   float* policy_out = nullptr, float* entropy_out = nullptr);

}  // namespace tpc
