// CUDA-core kernels around the tensor-core GEMMs: embeddings, LayerNorm backward, bias gradients,
// decoder cross-attention (one query per state), pointer head, value head, losses.
// Memory-bound row kernels: one warp per 128-float row (float4 per lane, coalesced 512 B), grid-stride.
#include "kernels.cuh"
#include "pointer_dev.cuh"

namespace tpc {

namespace {

constexpr float BIG_NUMBER = 1e9f;  // actor/pointer_attention.py:19, common/utils.py:38

__device__ __forceinline__ int map_row(const RowMap& m, int r) { return (r / m.L) * m.S + m.s0 + (r % m.L); }

__device__ __forceinline__ float4 round4(float4 v) {
  return make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
}

// ------------------------------------------------------------------------------------------------------
__global__ void small_dense_fwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ extra, int rows,
                                       RowMap map, const float* __restrict__ W, const float* __restrict__ b, int F,
                                       float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = (int)(idx >> 5), c4 = (int)(idx & 31);
  if (r >= rows) return;
  const int tok = map_row(map, r);
  float4 acc = __ldg(reinterpret_cast<const float4*>(b) + c4);
  for (int f = 0; f < F; ++f) {
    const float xf = (f < ldx) ? x[(size_t)tok * ldx + f] : extra[tok];
    const float4 w = __ldg(reinterpret_cast<const float4*>(W + (size_t)f * 128) + c4);
    // Dense = matmul then bias add (Keras); accumulate products first, add bias last would differ in rounding
    // order only; fp32 fma chain keeps |err| ~1e-7
    acc.x = fmaf(xf, w.x, acc.x); acc.y = fmaf(xf, w.y, acc.y); acc.z = fmaf(xf, w.z, acc.z); acc.w = fmaf(xf, w.w, acc.w);
  }
  reinterpret_cast<float4*>(out + (size_t)r * 128)[c4] = acc;   // forward activations stay fp32 (fp32-accurate consumers)
}

__global__ void small_dense_bwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ extra, int rows,
                                       RowMap map, const float* __restrict__ dOut, int F, float* __restrict__ dW,
                                       float* __restrict__ db, int rows_per_block) {
  const int j = threadIdx.x;  // 128 threads: output column
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(r0 + rows_per_block, rows);
  float aw[4] = {0.f, 0.f, 0.f, 0.f};
  float ab = 0.f;
  for (int r = r0; r < r1; ++r) {
    const int tok = map_row(map, r);
    const float d = dOut[(size_t)r * 128 + j];
    ab += d;
#pragma unroll
    for (int f = 0; f < 4; ++f)
      if (f < F) {
        const float xf = (f < ldx) ? x[(size_t)tok * ldx + f] : extra[tok];
        aw[f] = fmaf(xf, d, aw[f]);
      }
  }
  atomicAdd(db + j, ab);
  for (int f = 0; f < F; ++f) atomicAdd(dW + (size_t)f * 128 + j, aw[f]);
}

// ------------------------------------------------------------------------------------------------------
__global__ void ln_bwd_kernel(const float* __restrict__ dout, RowMap dmap, const float* __restrict__ out,
                              const float* __restrict__ rstd, const float* __restrict__ gamma,
                              const float* __restrict__ beta, int rows, float* __restrict__ dpre,
                              float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + lane);
  const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + lane);
  float4 ig;
  ig.x = gm.x != 0.f ? 1.f / gm.x : 0.f; ig.y = gm.y != 0.f ? 1.f / gm.y : 0.f;
  ig.z = gm.z != 0.f ? 1.f / gm.z : 0.f; ig.w = gm.w != 0.f ? 1.f / gm.w : 0.f;
  float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = blockIdx.x * nwarps + warp; r < rows; r += gridDim.x * nwarps) {
    const float4 d = reinterpret_cast<const float4*>(dout + (size_t)map_row(dmap, r) * 128)[lane];
    const float4 o = reinterpret_cast<const float4*>(out + (size_t)r * 128)[lane];
    const float rs = rstd[r];
    float4 xh, g;
    xh.x = (o.x - bt.x) * ig.x; xh.y = (o.y - bt.y) * ig.y; xh.z = (o.z - bt.z) * ig.z; xh.w = (o.w - bt.w) * ig.w;
    g.x = d.x * gm.x; g.y = d.y * gm.y; g.z = d.z * gm.z; g.w = d.w * gm.w;
    const float mg = warp_sum((g.x + g.y) + (g.z + g.w)) * (1.f / 128.f);
    const float mgx = warp_sum((g.x * xh.x + g.y * xh.y) + (g.z * xh.z + g.w * xh.w)) * (1.f / 128.f);
    float4 dp;
    dp.x = rs * (g.x - mg - xh.x * mgx); dp.y = rs * (g.y - mg - xh.y * mgx);
    dp.z = rs * (g.z - mg - xh.z * mgx); dp.w = rs * (g.w - mg - xh.w * mgx);
    reinterpret_cast<float4*>(dpre + (size_t)r * 128)[lane] = round4(dp);
    ag.x = fmaf(d.x, xh.x, ag.x); ag.y = fmaf(d.y, xh.y, ag.y); ag.z = fmaf(d.z, xh.z, ag.z); ag.w = fmaf(d.w, xh.w, ag.w);
    ab.x += d.x; ab.y += d.y; ab.z += d.z; ab.w += d.w;
  }
  __shared__ float4 sg[8][32], sb[8][32];
  sg[warp][lane] = ag;
  sb[warp][lane] = ab;
  __syncthreads();
  if (warp == 0) {
    for (int w = 1; w < nwarps; ++w) {
      ag.x += sg[w][lane].x; ag.y += sg[w][lane].y; ag.z += sg[w][lane].z; ag.w += sg[w][lane].w;
      ab.x += sb[w][lane].x; ab.y += sb[w][lane].y; ab.z += sb[w][lane].z; ab.w += sb[w][lane].w;
    }
    atomicAdd(dgamma + lane * 4 + 0, ag.x); atomicAdd(dgamma + lane * 4 + 1, ag.y);
    atomicAdd(dgamma + lane * 4 + 2, ag.z); atomicAdd(dgamma + lane * 4 + 3, ag.w);
    atomicAdd(dbeta + lane * 4 + 0, ab.x); atomicAdd(dbeta + lane * 4 + 1, ab.y);
    atomicAdd(dbeta + lane * 4 + 2, ab.z); atomicAdd(dbeta + lane * 4 + 3, ab.w);
  }
}

__global__ void colsum_kernel(const float* __restrict__ dY, int ld, int M, int N, float* __restrict__ db,
                              int rows_per_block) {
  const int j = blockIdx.y * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(r0 + rows_per_block, M);
  float a0 = 0.f, a1 = 0.f;
  int r = r0;
  for (; r + 1 < r1; r += 2) {
    a0 += dY[(size_t)r * ld + j];
    a1 += dY[(size_t)(r + 1) * ld + j];
  }
  if (r < r1) a0 += dY[(size_t)r * ld + j];
  atomicAdd(db + j, a0 + a1);
}

__global__ void scatter_rows_kernel(const float* __restrict__ src, int rows, RowMap map, float* __restrict__ dst) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = (int)(idx >> 5), c4 = (int)(idx & 31);
  if (r >= rows) return;
  reinterpret_cast<float4*>(dst + (size_t)map_row(map, r) * 128)[c4] =
      reinterpret_cast<const float4*>(src + (size_t)r * 128)[c4];
}

// ------------------------------------------------------------------------------------------------------
// decoder cross attention: one CTA per state, warp == head
// ------------------------------------------------------------------------------------------------------
constexpr int XA_MAX_S = 512;

__global__ void __launch_bounds__(256) cross_attn_bwd_kernel(const float* __restrict__ dctx, const float* __restrict__ q2,
                                                             const float* __restrict__ kvw, const float* __restrict__ pc,
                                                             int S, float* __restrict__ dq2, float* __restrict__ dkvw) {
  __shared__ float ds[8][XA_MAX_S];
  const int c = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* kv = kvw + (size_t)c * S * 384;
  float* dkv = dkvw + (size_t)c * S * 384;
  float q[16], dc[16];
#pragma unroll
  for (int d4 = 0; d4 < 4; ++d4) {
    const float4 t = reinterpret_cast<const float4*>(q2 + (size_t)c * 128 + h * 16)[d4];
    q[d4 * 4 + 0] = t.x; q[d4 * 4 + 1] = t.y; q[d4 * 4 + 2] = t.z; q[d4 * 4 + 3] = t.w;
    const float4 u = reinterpret_cast<const float4*>(dctx + (size_t)c * 128 + h * 16)[d4];
    dc[d4 * 4 + 0] = u.x; dc[d4 * 4 + 1] = u.y; dc[d4 * 4 + 2] = u.z; dc[d4 * 4 + 3] = u.w;
  }
  // pass 1: dp_j = dctx . V_j ; pdp = sum_j p_j dp_j
  float pdp = 0.f;
  for (int j = lane; j < S; j += 32) {
    const float4* vr = reinterpret_cast<const float4*>(kv + (size_t)j * 384 + 128 + h * 16);
    float dp = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < 4; ++d4) {
      const float4 v = vr[d4];
      dp = fmaf(dc[d4 * 4 + 0], v.x, dp); dp = fmaf(dc[d4 * 4 + 1], v.y, dp);
      dp = fmaf(dc[d4 * 4 + 2], v.z, dp); dp = fmaf(dc[d4 * 4 + 3], v.w, dp);
    }
    ds[h][j] = dp;
    pdp = fmaf(pc[((size_t)c * 8 + h) * S + j], dp, pdp);
  }
  pdp = warp_sum(pdp);
  // pass 2: ds_j = p_j (dp_j - pdp); dK2_j = ds_j q / 4; dV2_j = p_j dctx
  for (int j = lane; j < S; j += 32) {
    const float p = pc[((size_t)c * 8 + h) * S + j];
    const float dsj = p * (ds[h][j] - pdp);
    ds[h][j] = dsj;
    float4* dk = reinterpret_cast<float4*>(dkv + (size_t)j * 384 + h * 16);
    float4* dv = reinterpret_cast<float4*>(dkv + (size_t)j * 384 + 128 + h * 16);
    const float sk = dsj * 0.25f;
#pragma unroll
    for (int d4 = 0; d4 < 4; ++d4) {
      dk[d4] = round4(make_float4(sk * q[d4 * 4 + 0], sk * q[d4 * 4 + 1], sk * q[d4 * 4 + 2], sk * q[d4 * 4 + 3]));
      dv[d4] = round4(make_float4(p * dc[d4 * 4 + 0], p * dc[d4 * 4 + 1], p * dc[d4 * 4 + 2], p * dc[d4 * 4 + 3]));
    }
  }
  __syncwarp();
  // pass 3: dq = sum_j ds_j K2_j / 4
  const int d = lane & 15, half = lane >> 4;
  float acc = 0.f;
  for (int j = half; j < S; j += 2) acc = fmaf(ds[h][j], kv[(size_t)j * 384 + h * 16 + d], acc);
  acc += __shfl_xor_sync(0xffffffffu, acc, 16);
  if (half == 0) dq2[(size_t)c * 128 + h * 16 + d] = tf32_rn(acc * 0.25f);
}


// ------------------------------------------------------------------------------------------------------
// fused one-token decoder layer (see DecChainArgs in kernels.cuh): 2 * EPL states per CTA, 256 threads.
// Thread (grp = tid >> 7, col = tid & 127) owns output column `col` of the EPL states grp * EPL .. of the CTA; the
// weight matrices stream through the CTA once per layer (EPL = 4 for the update's large batches: a quarter of the L2
// traffic; EPL = 1 for the rollout's small ones: more CTAs in flight, the cross-attention of a CTA's states is serial).
// ------------------------------------------------------------------------------------------------------
constexpr int DC_MAX_DFF = 512;

// acc[j] += sum_k x[j][k] * W[k][col]  (W row-major with leading dimension ldw; x rows in shared memory, stride xs_ld)
template <int EPL>
__device__ __forceinline__ void dc_gemv(const float* __restrict__ W, int ldw, int K, const float* x, int xs_ld,
                                        float (&acc)[EPL]) {
#pragma unroll 8
  for (int k = 0; k < K; ++k) {
    const float w = __ldg(W + (size_t)k * ldw);
#pragma unroll
    for (int j = 0; j < EPL; ++j) acc[j] = fmaf(x[j * xs_ld + k], w, acc[j]);
  }
}

// LayerNorm of the EPL rows a thread group holds one column of each: the arithmetic of the GEMM epilogue (two-pass
// mean / variance, biased variance, eps inside the square root). red: [2 groups][4 warps][EPL rows].
template <int EPL>
__device__ __forceinline__ void dc_layernorm(const float (&t)[EPL], float gamma, float beta, float eps,
                                             float (*red)[4][4], int grp, int wg, int lane, float (&o)[EPL],
                                             float (&r)[EPL]) {
  float s[EPL], d[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) s[j] = warp_sum(t[j]);
  __syncthreads();                       // red is free (its previous readers are past this point)
  if (lane == 0)
#pragma unroll
    for (int j = 0; j < EPL; ++j) red[grp][wg][j] = s[j];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    const float m = (red[grp][0][j] + red[grp][1][j] + red[grp][2][j] + red[grp][3][j]) * (1.0f / 128.0f);
    d[j] = t[j] - m;
    s[j] = warp_sum(d[j] * d[j]);
  }
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int j = 0; j < EPL; ++j) red[grp][wg][j] = s[j];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    r[j] = 1.0f / sqrtf((red[grp][0][j] + red[grp][1][j] + red[grp][2][j] + red[grp][3][j]) * (1.0f / 128.0f) + eps);
    o[j] = d[j] * r[j] * gamma + beta;
  }
}

template <int EPL>
__global__ void __launch_bounds__(256) decoder_chain_kernel(const DecChainArgs a) {
  constexpr int EPC = 2 * EPL;
  __shared__ float xs[EPC][128];               // the layer's running vector (y, o1, o2, o3)
  __shared__ float ys[EPC][128];               // v1, q2
  __shared__ float hs[EPC][DC_MAX_DFF];        // ctx, then the FFN hidden
  __shared__ float ps[8][XA_MAX_S];            // cross-attention scores / probabilities of one state
  __shared__ float red[2][4][4];
  const int tid = threadIdx.x, col = tid & 127, grp = tid >> 7, warp = tid >> 5, lane = tid & 31, wg = warp & 3;
  const int el = grp * EPL;                    // first local state of this thread
  const int cb = blockIdx.x * EPC + el;        // its global index
  bool ok[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) ok[j] = cb + j < a.C;
  auto put = [&](float* dst, int ld, int off, const float (&v)[EPL]) {
#pragma unroll
    for (int j = 0; j < EPL; ++j)
      if (ok[j]) dst[(size_t)(cb + j) * ld + off] = v[j];
  };
  auto put_row = [&](float* dst, const float (&v)[EPL]) {   // one value per state (rstd)
    if (col == 0)
#pragma unroll
      for (int j = 0; j < EPL; ++j)
        if (ok[j]) dst[cb + j] = v[j];
  };

  // ---- input of the layer
  float y[EPL];
  if (a.y_in) {
#pragma unroll
    for (int j = 0; j < EPL; ++j) y[j] = ok[j] ? a.y_in[(size_t)(cb + j) * 128 + col] : 0.f;
  } else {   // decoder embedding: products accumulated onto the bias in feature order (small_dense_fwd_kernel)
#pragma unroll
    for (int j = 0; j < EPL; ++j) y[j] = __ldg(a.bemb + col);
    for (int f = 0; f < a.F; ++f) {
      const float w = __ldg(a.Wemb + (size_t)f * 128 + col);
#pragma unroll
      for (int j = 0; j < EPL; ++j) y[j] = fmaf(ok[j] ? a.dec_input[(size_t)(cb + j) * a.F + f] : 0.f, w, y[j]);
    }
    put(a.y0, 128, col, y);
  }
#pragma unroll
  for (int j = 0; j < EPL; ++j) xs[el + j][col] = y[j];
  __syncthreads();

  // ---- mha1 with a single key: attn1 = Dense_o1(Dense_v1(y))
  float t[EPL], o[EPL], r[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) t[j] = __ldg(a.bv1 + col);
  dc_gemv<EPL>(a.Wv1 + col, 128, 128, &xs[el][0], 128, t);
  put(a.v1, 128, col, t);
#pragma unroll
  for (int j = 0; j < EPL; ++j) ys[el + j][col] = t[j];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < EPL; ++j) t[j] = __ldg(a.bo1 + col);
  dc_gemv<EPL>(a.Wo1 + col, 128, 128, &ys[el][0], 128, t);
#pragma unroll
  for (int j = 0; j < EPL; ++j) t[j] += y[j];
  dc_layernorm<EPL>(t, __ldg(a.g1 + col), __ldg(a.be1 + col), a.ln_eps, red, grp, wg, lane, o, r);
  put(a.o1, 128, col, o);
  put_row(a.rstd1, r);
  float o1v[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) o1v[j] = o[j];
  __syncthreads();                       // every read of xs (y) is done
#pragma unroll
  for (int j = 0; j < EPL; ++j) xs[el + j][col] = o1v[j];
  __syncthreads();

  // ---- mha2: the query
#pragma unroll
  for (int j = 0; j < EPL; ++j) t[j] = __ldg(a.bq2 + col);
  dc_gemv<EPL>(a.Wq2 + col, 128, 128, &xs[el][0], 128, t);
  put(a.q2, 128, col, t);
  __syncthreads();                       // reads of ys (v1) are done
#pragma unroll
  for (int j = 0; j < EPL; ++j) ys[el + j][col] = t[j];
  __syncthreads();

  // ---- cross attention, one state at a time: warp == head (one query per state, actor/decoder_layer.py:26-27)
  const int S = a.S, Sf = a.win.Sfull > 0 ? a.win.Sfull : S;
  for (int e = 0; e < EPC; ++e) {
    const int c = blockIdx.x * EPC + e;
    if (c >= a.C) break;                 // uniform over the CTA
    const int h = warp;
    const float* kv = a.kvw + (size_t)c * S * 384;
    float q[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) q[d] = ys[e][h * 16 + d];
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) {
      const float4* kr = reinterpret_cast<const float4*>(kv + (size_t)j * 384 + h * 16);
      float sc = 0.f;
#pragma unroll
      for (int d4 = 0; d4 < 4; ++d4) {
        const float4 k = kr[d4];
        sc = fmaf(q[d4 * 4 + 0], k.x, sc); sc = fmaf(q[d4 * 4 + 1], k.y, sc);
        sc = fmaf(q[d4 * 4 + 2], k.z, sc); sc = fmaf(q[d4 * 4 + 3], k.w, sc);
      }
      sc = sc * 0.25f + a.mask[(size_t)c * Sf + a.win.full(j)] * -BIG_NUMBER;   // / sqrt(16), + mask * -1e9 (common/utils.py:34-38)
      ps[h][j] = sc;
      mx = fmaxf(mx, sc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) {
      const float ex = expf(ps[h][j] - mx);
      ps[h][j] = ex;
      sum += ex;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < S; j += 32) {
      const float p = ps[h][j] * inv;
      ps[h][j] = p;
      a.pc[((size_t)c * 8 + h) * S + j] = p;
    }
    __syncwarp();
    const int d = lane & 15, half = lane >> 4;
    float acc = 0.f;
    for (int j = half; j < S; j += 2) acc = fmaf(ps[h][j], kv[(size_t)j * 384 + 128 + h * 16 + d], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    if (half == 0) {
      a.ctx[(size_t)c * 128 + h * 16 + d] = acc;
      hs[e][h * 16 + d] = acc;           // hs doubles as the ctx buffer until the FFN needs it
    }
    __syncwarp();
  }
  __syncthreads();

  // ---- mha2 output projection + LayerNorm 2
#pragma unroll
  for (int j = 0; j < EPL; ++j) t[j] = __ldg(a.bo2 + col);
  dc_gemv<EPL>(a.Wo2 + col, 128, 128, &hs[el][0], DC_MAX_DFF, t);
#pragma unroll
  for (int j = 0; j < EPL; ++j) t[j] += o1v[j];
  dc_layernorm<EPL>(t, __ldg(a.g2 + col), __ldg(a.be2 + col), a.ln_eps, red, grp, wg, lane, o, r);
  put(a.o2, 128, col, o);
  put_row(a.rstd2, r);
  float o2v[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) o2v[j] = o[j];
  __syncthreads();                       // reads of xs (o1) and hs (ctx) are done
#pragma unroll
  for (int j = 0; j < EPL; ++j) xs[el + j][col] = o2v[j];
  __syncthreads();

  // ---- FFN
  for (int jt = 0; jt < a.dff; jt += 128) {
#pragma unroll
    for (int j = 0; j < EPL; ++j) t[j] = __ldg(a.bf1 + jt + col);
    dc_gemv<EPL>(a.Wf1 + jt + col, a.dff, 128, &xs[el][0], 128, t);
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
      t[j] = fmaxf(t[j], 0.f);
      hs[el + j][jt + col] = t[j];
    }
    put(a.f, a.dff, jt + col, t);
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < EPL; ++j) t[j] = __ldg(a.bf2 + col);
  dc_gemv<EPL>(a.Wf2 + col, 128, a.dff, &hs[el][0], DC_MAX_DFF, t);
#pragma unroll
  for (int j = 0; j < EPL; ++j) t[j] += o2v[j];
  dc_layernorm<EPL>(t, __ldg(a.g3 + col), __ldg(a.be3 + col), a.ln_eps, red, grp, wg, lane, o, r);
  put(a.o3, 128, col, o);
  put_row(a.rstd3, r);

  // ---- pointer: W1 dec (last layer)
  if (a.Wp1) {
    __syncthreads();                     // reads of xs (o2) are done
#pragma unroll
    for (int j = 0; j < EPL; ++j) xs[el + j][col] = o[j];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < EPL; ++j) t[j] = __ldg(a.bp1 + col);
    dc_gemv<EPL>(a.Wp1 + col, 128, 128, &xs[el][0], 128, t);
    put(a.w1d, 128, col, t);
  }
}

// ------------------------------------------------------------------------------------------------------
// pointer head: one CTA (8 warps) per state; warp per key, lane holds 4 of the 128 units
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pointer_fwd_kernel(const float* __restrict__ w1d, const float* __restrict__ kvw,
                                                          const float* __restrict__ V, const float* __restrict__ bV,
                                                          float clipC, int has_clip, const float* __restrict__ ptr_mask,
                                                          int S, float* __restrict__ score, float* __restrict__ logits,
                                                          float* __restrict__ probs, int32_t* __restrict__ argmax, TokWin w) {
  __shared__ float sl[XA_MAX_S];
  __shared__ float red[8];
  __shared__ int redi[8];
  const int c = blockIdx.x;
  const int best = pointer_head(c, w1d, kvw, V, bV, clipC, has_clip, ptr_mask, S, score, logits, probs, sl, red, redi, w);
  if (threadIdx.x == 0 && argmax) argmax[c] = best;
}

__global__ void __launch_bounds__(256) pointer_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ score,
                                                          const float* __restrict__ w1d, const float* __restrict__ kvw,
                                                          const float* __restrict__ V, float clipC, int has_clip, int S,
                                                          float* __restrict__ dw1d, float* __restrict__ dkvw,
                                                          float* __restrict__ dV, float* __restrict__ dbV, TokWin w) {
  __shared__ float4 sw[8][32], sv[8][32];
  __shared__ float sb[8];
  const int c = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4 wd = reinterpret_cast<const float4*>(w1d + (size_t)c * 128)[lane];
  const float4 v = __ldg(reinterpret_cast<const float4*>(V) + lane);
  float4 aw = make_float4(0.f, 0.f, 0.f, 0.f), av = make_float4(0.f, 0.f, 0.f, 0.f);
  float ab = 0.f;
  for (int j = warp; j < S; j += 8) {
    const size_t row = (size_t)c * S + j;
    float dsc = dlogits[(size_t)c * (w.Sfull > 0 ? w.Sfull : S) + w.full(j)];
    if (has_clip) {
      const float t = tanhf(score[row]);
      dsc *= clipC * (1.f - t * t);
    }
    const float4 we = reinterpret_cast<const float4*>(kvw + row * 384 + 256)[lane];
    float4 t, dp;
    t.x = tanhf(wd.x + we.x); t.y = tanhf(wd.y + we.y); t.z = tanhf(wd.z + we.z); t.w = tanhf(wd.w + we.w);
    dp.x = dsc * v.x * (1.f - t.x * t.x); dp.y = dsc * v.y * (1.f - t.y * t.y);
    dp.z = dsc * v.z * (1.f - t.z * t.z); dp.w = dsc * v.w * (1.f - t.w * t.w);
    reinterpret_cast<float4*>(dkvw + row * 384 + 256)[lane] = round4(dp);
    aw.x += dp.x; aw.y += dp.y; aw.z += dp.z; aw.w += dp.w;
    av.x = fmaf(dsc, t.x, av.x); av.y = fmaf(dsc, t.y, av.y); av.z = fmaf(dsc, t.z, av.z); av.w = fmaf(dsc, t.w, av.w);
    ab += dsc;
  }
  sw[warp][lane] = aw;
  sv[warp][lane] = av;
  if (lane == 0) sb[warp] = ab;
  __syncthreads();
  if (warp == 0) {
    for (int w = 1; w < 8; ++w) {
      aw.x += sw[w][lane].x; aw.y += sw[w][lane].y; aw.z += sw[w][lane].z; aw.w += sw[w][lane].w;
      av.x += sv[w][lane].x; av.y += sv[w][lane].y; av.z += sv[w][lane].z; av.w += sv[w][lane].w;
    }
    reinterpret_cast<float4*>(dw1d + (size_t)c * 128)[lane] = round4(aw);
    atomicAdd(dV + lane * 4 + 0, av.x); atomicAdd(dV + lane * 4 + 1, av.y);
    atomicAdd(dV + lane * 4 + 2, av.z); atomicAdd(dV + lane * 4 + 3, av.w);
    if (lane == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += sb[w];
      atomicAdd(dbV, t);
    }
  }
}

// ------------------------------------------------------------------------------------------------------
__global__ void value_head_fwd_kernel(const float* __restrict__ h1, const float* __restrict__ Wv,
                                      const float* __restrict__ bv, int C, float* __restrict__ v) {
  const int lane = threadIdx.x & 31;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= C) return;
  const float4 a = reinterpret_cast<const float4*>(h1 + (size_t)c * 128)[lane];
  const float4 w = __ldg(reinterpret_cast<const float4*>(Wv) + lane);
  float s = a.x * w.x;
  s = fmaf(a.y, w.y, s); s = fmaf(a.z, w.z, s); s = fmaf(a.w, w.w, s);
  s = warp_sum(s);
  if (lane == 0) v[c] = s + __ldg(bv);
}

__global__ void value_head_bwd_kernel(const float* __restrict__ dv, const float* __restrict__ h1,
                                      const float* __restrict__ Wv, int C, float* __restrict__ dh1,
                                      float* __restrict__ dWv, float* __restrict__ dbv) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float4 w = __ldg(reinterpret_cast<const float4*>(Wv) + lane);
  float4 aw = make_float4(0.f, 0.f, 0.f, 0.f);
  float ab = 0.f;
  for (int c = blockIdx.x * nwarps + warp; c < C; c += gridDim.x * nwarps) {
    const float d = dv[c];
    const float4 a = reinterpret_cast<const float4*>(h1 + (size_t)c * 128)[lane];
    reinterpret_cast<float4*>(dh1 + (size_t)c * 128)[lane] = round4(make_float4(d * w.x, d * w.y, d * w.z, d * w.w));
    aw.x = fmaf(d, a.x, aw.x); aw.y = fmaf(d, a.y, aw.y); aw.z = fmaf(d, a.z, aw.z); aw.w = fmaf(d, a.w, aw.w);
    ab += d;
  }
  atomicAdd(dWv + lane * 4 + 0, aw.x); atomicAdd(dWv + lane * 4 + 1, aw.y);
  atomicAdd(dWv + lane * 4 + 2, aw.z); atomicAdd(dWv + lane * 4 + 3, aw.w);
  if (lane == 0) atomicAdd(dbv, ab);
}

__global__ void fill_rows_bias_kernel(float* __restrict__ out, const float* __restrict__ bias, int rows) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = (int)(idx >> 5), c4 = (int)(idx & 31);
  if (r >= rows) return;
  reinterpret_cast<float4*>(out + (size_t)r * 128)[c4] = __ldg(reinterpret_cast<const float4*>(bias) + c4);
}

__global__ void round_tf32_kernel(float* __restrict__ x, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    x[i] = tf32_rn(x[i]);
}

// ------------------------------------------------------------------------------------------------------
// bootstrap: value after the last step per episode (NULL = 0, agents/trainer.py:103-105); horizontal: R laid out (B,T)
// like the reference's return value instead of the step-major vector the critic consumes
__global__ void returns_kernel(const float* __restrict__ rewards, const float* __restrict__ bootstrap, int B, int T,
                               float gamma, int horizontal, float* __restrict__ R) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float boot = bootstrap ? bootstrap[b] : 0.f;
  for (int t = T - 1; t >= 0; --t) {
    // estimate = r + gamma * bootstrap (agents/agent.py:117-118), fp32, no fma contraction
    const float est = __fadd_rn(rewards[(size_t)b * T + t], __fmul_rn(gamma, boot));
    R[horizontal ? (size_t)b * T + t : (size_t)t * B + b] = est;
    boot = est;
  }
}

__global__ void critic_loss_bwd_kernel(const float* __restrict__ R, const float* __restrict__ V, int n, float c_v,
                                       float inv_total, float out_scale, float* __restrict__ adv, float* __restrict__ dv,
                                       float* __restrict__ losses) {
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float a = R[i] - V[i];
    adv[i] = a;
    if (dv) dv[i] = -2.f * c_v * a * inv_total;
    acc = fmaf(c_v * a, a, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) atomicAdd(losses + 0, acc * out_scale);
}

// warp per state
// dlogits / policy_out / entropy_out are optional; the three sums are scaled by out_scale (1 in the fused update, which
// divides by the job-wide state count afterwards; 1 / n for the step-wise API's means)
__global__ void actor_loss_bwd_kernel(const float* __restrict__ logits, const int32_t* __restrict__ actions,
                                      const float* __restrict__ adv, int n, int S, float c_e, float inv_total,
                                      float out_scale, float* __restrict__ dlogits, float* __restrict__ policy_out,
                                      float* __restrict__ entropy_out, float* __restrict__ losses) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const float* l = logits + (size_t)i * S;
  float mx = -INFINITY;
  for (int j = lane; j < S; j += 32) mx = fmaxf(mx, l[j]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < S; j += 32) sum += expf(l[j] - mx);
  sum = warp_sum(sum);
  const float lse = mx + logf(sum);
  const int a = min(max(actions[i], 0), S - 1);   // host-supplied index: never read outside the row
  const float ad = adv[i];
  float ent = 0.f;
  for (int j = lane; j < S; j += 32) {
    const float logp = l[j] - lse;
    const float p = expf(logp);
    ent -= p * logp;  // masked entries: p == 0 exactly (logp ~ -1e9), contributes 0
    if (dlogits) dlogits[(size_t)i * S + j] = ad * (p - (j == a ? 1.f : 0.f)) * inv_total;
  }
  ent = warp_sum(ent);
  if (lane == 0) {
    const float pl = -(l[a] - lse);
    if (policy_out) policy_out[i] = pl;
    if (entropy_out) entropy_out[i] = ent;
    atomicAdd(losses + 1, pl * out_scale);
    atomicAdd(losses + 2, (pl * ad - c_e * ent) * out_scale);
    atomicAdd(losses + 3, ent * out_scale);
  }
}

inline int blocks_for(long long threads, int bs) { return (int)((threads + bs - 1) / bs); }

}  // namespace

int small_dense_fwd(const float* x, int ldx, const float* extra, int rows, RowMap map, const float* W, const float* b,
                    int F, float* out, cudaStream_t st) {
  if (rows <= 0) return TPC_OK;
  if (F > ldx + (extra ? 1 : 0) || F > 4) return TPC_ERR_SHAPE;
  small_dense_fwd_kernel<<<blocks_for((long long)rows * 32, 256), 256, 0, st>>>(x, ldx, extra, rows, map, W, b, F, out);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int small_dense_bwd(const float* x, int ldx, const float* extra, int rows, RowMap map, const float* dOut, int F,
                    float* dW, float* db, cudaStream_t st) {
  if (rows <= 0) return TPC_OK;
  if (F > 4) return TPC_ERR_SHAPE;
  const int rpb = 256;
  small_dense_bwd_kernel<<<ceil_div(rows, rpb), 128, 0, st>>>(x, ldx, extra, rows, map, dOut, F, dW, db, rpb);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int ln_bwd(const float* dout, RowMap dmap, const float* out, const float* rstd, const float* gamma, const float* beta,
           int rows, float* dpre, float* dgamma, float* dbeta, cudaStream_t st) {
  if (rows <= 0) return TPC_OK;
  int blocks = ceil_div(rows, 8 * 16);
  if (blocks > 148 * 8) blocks = 148 * 8;
  ln_bwd_kernel<<<blocks, 256, 0, st>>>(dout, dmap, out, rstd, gamma, beta, rows, dpre, dgamma, dbeta);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int colsum(const float* dY, int ld, int M, int N, float* db, cudaStream_t st) {
  if (M <= 0) return TPC_OK;
  const int rpb = 512;
  dim3 grid(ceil_div(M, rpb), ceil_div(N, 128));
  colsum_kernel<<<grid, 128, 0, st>>>(dY, ld, M, N, db, rpb);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int scatter_rows(const float* src, int rows, RowMap map, float* dst, cudaStream_t st) {
  if (rows <= 0) return TPC_OK;
  scatter_rows_kernel<<<blocks_for((long long)rows * 32, 256), 256, 0, st>>>(src, rows, map, dst);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int cross_attn_bwd(const float* dctx, const float* q2, const float* kvw, const float* pc, int C, int S, float* dq2,
                   float* dkvw, cudaStream_t st) {
  if (C <= 0) return TPC_OK;
  if (S > XA_MAX_S) return TPC_ERR_SHAPE;
  cross_attn_bwd_kernel<<<C, 256, 0, st>>>(dctx, q2, kvw, pc, S, dq2, dkvw);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int decoder_chain_fwd(const DecChainArgs& a, cudaStream_t st) {
  if (a.C <= 0) return TPC_OK;
  if (a.S > XA_MAX_S || a.dff % 128 != 0 || a.dff < 128 || a.dff > DC_MAX_DFF) return TPC_ERR_SHAPE;
  if (a.C >= 4096) decoder_chain_kernel<4><<<ceil_div(a.C, 8), 256, 0, st>>>(a);
  else decoder_chain_kernel<1><<<ceil_div(a.C, 2), 256, 0, st>>>(a);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int pointer_fwd(const float* w1d, const float* kvw, const float* V, const float* bV, float clipC, int has_clip,
                const float* ptr_mask, int C, int S, float* score, float* logits, float* probs, int32_t* argmax,
                cudaStream_t st, TokWin w) {
  if (C <= 0) return TPC_OK;
  if (S > XA_MAX_S) return TPC_ERR_SHAPE;
  pointer_fwd_kernel<<<C, 256, 0, st>>>(w1d, kvw, V, bV, clipC, has_clip, ptr_mask, S, score, logits, probs, argmax, w);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int pointer_bwd(const float* dlogits, const float* score, const float* w1d, const float* kvw, const float* V,
                float clipC, int has_clip, int C, int S, float* dw1d, float* dkvw, float* dV, float* dbV,
                cudaStream_t st, TokWin w) {
  if (C <= 0) return TPC_OK;
  pointer_bwd_kernel<<<C, 256, 0, st>>>(dlogits, score, w1d, kvw, V, clipC, has_clip, S, dw1d, dkvw, dV, dbV, w);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int value_head_fwd(const float* h1, const float* Wv, const float* bv, int C, float* v, cudaStream_t st) {
  if (C <= 0) return TPC_OK;
  value_head_fwd_kernel<<<blocks_for((long long)C * 32, 256), 256, 0, st>>>(h1, Wv, bv, C, v);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int value_head_bwd(const float* dv, const float* h1, const float* Wv, int C, float* dh1, float* dWv, float* dbv,
                   cudaStream_t st) {
  if (C <= 0) return TPC_OK;
  int blocks = ceil_div(C, 8 * 8);
  if (blocks > 296) blocks = 296;
  value_head_bwd_kernel<<<blocks, 256, 0, st>>>(dv, h1, Wv, C, dh1, dWv, dbv);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int fill_rows_bias(float* out, const float* bias, int rows, cudaStream_t st) {
  if (rows <= 0) return TPC_OK;
  fill_rows_bias_kernel<<<blocks_for((long long)rows * 32, 256), 256, 0, st>>>(out, bias, rows);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int round_tf32_inplace(float* x, size_t n, cudaStream_t st) {
  if (n == 0) return TPC_OK;
  int blocks = (int)((n + 1023) / 1024);
  if (blocks > 1184) blocks = 1184;
  round_tf32_kernel<<<blocks, 256, 0, st>>>(x, n);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int discounted_returns(const float* rewards, int B, int T, float gamma, float* R, cudaStream_t st,
                       const float* bootstrap, int horizontal) {
  if (B <= 0 || T <= 0) return TPC_OK;
  returns_kernel<<<ceil_div(B, 128), 128, 0, st>>>(rewards, bootstrap, B, T, gamma, horizontal, R);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int critic_loss_bwd(const float* R, const float* V, int n, float c_v, float inv_total, float* adv, float* dv,
                    float* losses, cudaStream_t st, float out_scale) {
  if (n <= 0) return TPC_OK;
  int blocks = ceil_div(n, 256);
  if (blocks > 296) blocks = 296;
  critic_loss_bwd_kernel<<<blocks, 256, 0, st>>>(R, V, n, c_v, inv_total, out_scale, adv, dv, losses);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int actor_loss_bwd(const float* logits, const int32_t* actions, const float* adv, int n, int S, float c_e,
                   float inv_total, float* dlogits, float* losses, cudaStream_t st, float out_scale, float* policy_out,
                   float* entropy_out) {
  if (n <= 0) return TPC_OK;
  actor_loss_bwd_kernel<<<blocks_for((long long)n * 32, 256), 256, 0, st>>>(logits, actions, adv, n, S, c_e, inv_total,
                                                                             out_scale, dlogits, policy_out, entropy_out,
                                                                             losses);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

}  // namespace tpc
