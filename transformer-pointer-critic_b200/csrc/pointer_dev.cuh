// Pointer head of one state by one CTA of 256 threads (actor/pointer_attention.py:40-66), shared by
// pointer_fwd_kernel (kernels_simt.cu) and the fused decision kernel of the rollout (decide.cu).
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace tpc {

constexpr int PTR_MAX_S = 512;            // == XA_MAX_S
constexpr float PTR_BIG_NUMBER = 1e9f;    // actor/pointer_attention.py:19

// score = V . tanh(W1 dec + W2 enc) + bV ; logits = C tanh(score) - mask 1e9 ; probs = softmax(logits), left in
// sl[0..S); returns argmax(probs) (lowest index wins, tf.argmax) on every thread. sl: PTR_MAX_S floats, red / redi: 8.
// S counts the COMPACT tokens of a state (kvw, score, sl); ptr_mask / logits / probs are the full (C, w.Sfull) tensors
// and the returned index is a full one. Tokens left out of the window get logit -1e9 and probability 0: what the
// reference computes for them (|C tanh| <= C << ulp(1e9) = 64, and exp(-1e9 - max) == 0 in fp32).
__device__ __forceinline__ int pointer_head(int c, const float* __restrict__ w1d, const float* __restrict__ kvw,
                                            const float* __restrict__ V, const float* __restrict__ bV, float clipC,
                                            int has_clip, const float* __restrict__ ptr_mask, int S,
                                            float* __restrict__ score, float* __restrict__ logits,
                                            float* __restrict__ probs, float* sl, float* red, int* redi,
                                            TokWin w = TokWin()) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Sf = w.Sfull > 0 ? w.Sfull : S;
  for (int j = w.N + threadIdx.x; j < w.N + w.gap && w.Sfull > 0; j += 256) {
    if (logits) logits[(size_t)c * Sf + j] = -PTR_BIG_NUMBER;
    if (probs) probs[(size_t)c * Sf + j] = 0.f;
  }
  const float4 wd = reinterpret_cast<const float4*>(w1d + (size_t)c * 128)[lane];
  const float4 v = __ldg(reinterpret_cast<const float4*>(V) + lane);
  const float bv = __ldg(bV);
  for (int j = warp; j < S; j += 8) {
    const float4 we = reinterpret_cast<const float4*>(kvw + ((size_t)c * S + j) * 384 + 256)[lane];
    float s = v.x * tanhf(wd.x + we.x);
    s = fmaf(v.y, tanhf(wd.y + we.y), s);
    s = fmaf(v.z, tanhf(wd.z + we.z), s);
    s = fmaf(v.w, tanhf(wd.w + we.w), s);
    s = warp_sum(s) + bv;
    if (lane == 0) {
      float l = has_clip ? clipC * tanhf(s) : s;
      l -= ptr_mask[(size_t)c * Sf + w.full(j)] * PTR_BIG_NUMBER;
      sl[j] = l;
      if (score) score[(size_t)c * S + j] = s;
      if (logits) logits[(size_t)c * Sf + w.full(j)] = l;
    }
  }
  __syncthreads();
  // softmax over S + argmax(probs) (lowest index wins, tf.argmax)
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < S; j += 256) mx = fmaxf(mx, sl[j]);
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int j = threadIdx.x; j < S; j += 256) {
    const float e = expf(sl[j] - mx);
    sl[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += red[w];
  // NaN logits (diverged weights) compare false everywhere: threads that own a column start from it, the others
  // from column 0 with a probability no real one can lose against, so the result is always inside [0, S) like
  // tf.argmax's (which returns an in-range index for NaN rows too).
  float best = -1.f;
  int besti = threadIdx.x < S ? threadIdx.x : 0;
  for (int j = threadIdx.x; j < S; j += 256) {
    const float p = sl[j] / sum;
    sl[j] = p;
    if (probs) probs[(size_t)c * Sf + w.full(j)] = p;
    if (p > best) { best = p; besti = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
  }
  __syncthreads();
  if (lane == 0) { red[warp] = best; redi[warp] = besti; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (red[w] > best || (red[w] == best && redi[w] < besti)) { best = red[w]; besti = redi[w]; }
    redi[0] = w.full(min(max(besti, 0), S - 1));
  }
  __syncthreads();
  return redi[0];
}

}  // namespace tpc
