// Multi-head self-attention of one encoder stack for short sequences (the reference tops out at
// 151 = 51 nodes + 100 rules): K/V (or Q/dO) of one (state, head) staged in shared memory,
// QK^T / PV on the legacy warp-level tensor path (mma.sync m16n8k8 TF32, fp32 accumulate), softmax with
// warp shuffles in registers.  Restates common/attention.py:37-48 + common/utils.py:31-45 (forward) and the
// autodiff of the same ops (agents/trainer.py:115-141).
//
// head_dim = 16 makes each head a K=16 contraction: far too thin for a 128-row tcgen05 tile, so
// this stays a warp-level kernel; the 128-wide projections around it are the tcgen05 GEMMs in gemm_tc.cu.
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace tpc {

namespace {

// phase stamps for tests/gpu_checks/micro/attn_phase.cu (compiled out of the library)
#ifdef ATTN_PROFILE
__device__ long long attn_prof[8 * 32768];
#define ATTN_STAMP(i)                                                                    \
  do {                                                                                    \
    if (threadIdx.x == 0 && blockIdx.x < 32768) attn_prof[blockIdx.x * 8 + (i)] = clock64(); \
  } while (0)
#else
#define ATTN_STAMP(i)
#endif

constexpr float BIG_NUMBER = 1e9f;

__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t fbits(float x) { return __float_as_uint(x); }
// 2^x on the SFU without the denormal fix-up sequence of exp2f/__expf (inputs <= 0 here, tiny results flush to 0)
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float QSCALE = 0.25f * LOG2E;   // 1/sqrt(depth=16), scores kept in log2 units so that exp == one ex2
__device__ __forceinline__ uint32_t tfbits(float x) { return tf32_rn_bits(x); }   // MMA truncates: add == round

// ---- operand images in shared memory ----
// The legacy tensor path issues one MMA per ~8.7 clocks per SM sub-partition whatever its shape, and that pipe is
// what bounds these kernels. Contractions over the 16-wide head dimension therefore use FP16 m16n8k16 (one MMA for
// all 16 dims; FP16 carries the same 11 significand bits as TF32): rows are first scaled by a per-CTA power of two
// so every row norm is <= 2^14 (no overflow; elements 2^-28 below the largest row norm start to lose precision), and split hi + lo for the 3-MMA "fp32-level" product q.k = ql.kh + qh.kl + qh.kh.
// Row image (16 words per row): for t = 0..3: [hi(2t,2t+1) hi(2t+8,2t+9) lo(2t,2t+1) lo(2t+8,2t+9)], i.e. lane
// (g, t) gets its whole A or B fragment pair, hi and lo, with one conflict-free LDS.128.
// Vt[d][VS]: V transposed and TF32-rounded, so the P.V fragment pair (keys 2t, 2t + 1) is one LDS.64 (the
// contraction over keys stays TF32 m16n8k8: P can span the whole fp32 exponent range).
__host__ __device__ inline int vt_stride(int Lp) { return (Lp % 32 == 0 || Lp % 32 == 16) ? Lp + 8 : Lp; }
// scores are kept in log2 units; when |q| * max|k| stays below this bound the softmax needs no row-max pass
// (2^s can neither overflow nor lose a term to underflow, and p / sum is scale invariant in floating point)
constexpr float EASY_BOUND = 96.f;

__device__ __forceinline__ void mma_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                        uint32_t b1, float c0, float c1, float c2, float c3) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(c0), "f"(c1), "f"(c2), "f"(c3));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }
// Power-of-two scale `down` with sqrt(n2) * down <= 2^target (n2 = largest squared row norm, so every element is
// covered), and its inverse. A target well above 1 keeps the lo parts / small elements out of FP16's subnormal range
// (below 2^-14 only 2^-24 absolute precision is left); FP16 tops out at 65504 ~ 2^16.
__device__ __forceinline__ void pow2_scale(float n2, int target, float& down, float& up) {
  int ex = 0;
  frexpf(n2, &ex);                     // n2 = m 2^ex, m in [0.5, 1)  ->  sqrt(n2) < 2^ceil(ex / 2)
  int e = ((ex + 1) >> 1) - target;
  e = max(-100, min(100, e));          // (inf / NaN / denormal norms: keep the scale finite)
  down = __int_as_float((127 - e) << 23);
  up = __int_as_float((127 + e) << 23);
}
// hi / lo FP16 row image of 16 scaled values
__device__ __forceinline__ void store_split_row(uint32_t* row, const float4 (&v)[4], float scale) {
  const float x[16] = {v[0].x * scale, v[0].y * scale, v[0].z * scale, v[0].w * scale, v[1].x * scale, v[1].y * scale,
                       v[1].z * scale, v[1].w * scale, v[2].x * scale, v[2].y * scale, v[2].z * scale, v[2].w * scale,
                       v[3].x * scale, v[3].y * scale, v[3].z * scale, v[3].w * scale};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const uint32_t h0 = pack_h2(x[2 * t], x[2 * t + 1]), h1 = pack_h2(x[2 * t + 8], x[2 * t + 9]);
    const float2 f0 = unpack_h2(h0), f1 = unpack_h2(h1);
    const uint32_t l0 = pack_h2(x[2 * t] - f0.x, x[2 * t + 1] - f0.y);
    const uint32_t l1 = pack_h2(x[2 * t + 8] - f1.x, x[2 * t + 9] - f1.y);
    *reinterpret_cast<uint4*>(row + 4 * t) = make_uint4(h0, h1, l0, l1);
  }
}
// Q rows go to a fragment-major image: per 16-row block, hi[lane][4] then lo[lane][4] (lane = g * 4 + t of the MMA
// A fragment: a0 = (row g, dims 2t, 2t+1), a1 = (row g + 8, same), a2 = (row g, dims 2t + 8, 2t + 9), a3 = (row g + 8, ...)),
// so a lane's quad is one LDS.128 landing in consecutive registers (no per-MMA register shuffling).
__device__ __forceinline__ void store_split_row_frag(uint32_t* img, int r, const float4 (&v)[4], float scale) {
  const float x[16] = {v[0].x * scale, v[0].y * scale, v[0].z * scale, v[0].w * scale, v[1].x * scale, v[1].y * scale,
                       v[1].z * scale, v[1].w * scale, v[2].x * scale, v[2].y * scale, v[2].z * scale, v[2].w * scale,
                       v[3].x * scale, v[3].y * scale, v[3].z * scale, v[3].w * scale};
  const int g = r & 7, half = (r >> 3) & 1;
  uint32_t* blk = img + (r >> 4) * 256 + g * 16 + half;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const uint32_t h0 = pack_h2(x[2 * t], x[2 * t + 1]), h1 = pack_h2(x[2 * t + 8], x[2 * t + 9]);
    const float2 f0 = unpack_h2(h0), f1 = unpack_h2(h1);
    blk[t * 4] = h0;
    blk[t * 4 + 2] = h1;
    blk[128 + t * 4] = pack_h2(x[2 * t] - f0.x, x[2 * t + 1] - f0.y);
    blk[128 + t * 4 + 2] = pack_h2(x[2 * t + 8] - f1.x, x[2 * t + 9] - f1.y);
  }
}
__device__ __forceinline__ float norm2(const float4 (&v)[4]) {
  float n = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) n += v[q].x * v[q].x + v[q].y * v[q].y + v[q].z * v[q].z + v[q].w * v[q].w;
  return n;
}
// ordered list of the live 8-key tiles (a tile is live if it holds an unmasked key), built by one warp with ballots.
// A masked key contributes exp(-1e9 - max) == 0 exactly in fp32, so dead tiles are skipped: exact, and
// during an episode a third to a half of the keys (placed rules, full nodes) are masked. If nothing is unmasked
// (cannot happen on the reference's trajectories) every tile stays live and the softmax is uniform, as in the reference.
__device__ __forceinline__ void build_live_list(float* Ms, int* lidx, int* n_out, int ntiles, int lane, float fill) {
  int n = 0;
  for (int nt0 = 0; nt0 < ntiles; nt0 += 32) {
    const int nt = nt0 + lane;
    bool l = false;
    if (nt < ntiles) {
      const float4 m0 = *reinterpret_cast<const float4*>(Ms + nt * 8), m1 = *reinterpret_cast<const float4*>(Ms + nt * 8 + 4);
      l = m0.x == 0.f || m0.y == 0.f || m0.z == 0.f || m0.w == 0.f || m1.x == 0.f || m1.y == 0.f || m1.z == 0.f || m1.w == 0.f;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, l);
    if (l) lidx[n + __popc(bal & ((1u << lane) - 1u))] = nt;
    n += __popc(bal);
  }
  if (n == 0) {
    // nothing unmasked: in the reference's fp32 the scores vanish in the rounding of score + (-1e9) and the softmax is
    // uniform over all keys. Same here: every tile live, the mask term dropped; the caller zeroes the score scale
    // (n_out < 0 flags the case).
    for (int nt = lane; nt < ntiles; nt += 32) lidx[nt] = nt;
    for (int j = lane; j < ntiles * 8; j += 32)
      if (Ms[j] != -INFINITY) Ms[j] = fill;
    n = -ntiles;
  }
  if (lane == 0) *n_out = n;
}

// RB 16-row query blocks (rows r0 + rstride b) of one warp against all live key tiles. The blocks share every K / V
// fragment load and give the scheduler RB independent MMA chains.
template <int RB, bool PV3X, bool EASY>
__device__ __forceinline__ void fwd_row_blocks(int r0, int rstride, int L, int hg, int h, int g, int t, const uint32_t* Qi,
                                               const uint32_t* Ki, const float* Vt, const float* Vl, int VS,
                                               const float* Ms, const int* lidx, int nlive, float sc_mul,
                                               float* __restrict__ att_c, float* __restrict__ lse_c) {
  // A fragments: the Q image is fragment-major, one LDS.128 is the whole register quad of a lane
  uint32_t qh[RB][4], ql[RB][4];
#pragma unroll
  for (int b = 0; b < RB; ++b) {
    const uint32_t* qf = Qi + (r0 + rstride * b) * 16 + (g * 4 + t) * 4;
    const uint4 hq = *reinterpret_cast<const uint4*>(qf), lq = *reinterpret_cast<const uint4*>(qf + 128);
    qh[b][0] = hq.x, qh[b][1] = hq.y, qh[b][2] = hq.z, qh[b][3] = hq.w;
    ql[b][0] = lq.x, ql[b][1] = lq.y, ql[b][2] = lq.z, ql[b][3] = lq.w;
  }
  const uint32_t* kbase = Ki + g * 16 + 4 * t;
  const float* vbase = Vt + g * VS + 2 * t;
  const float* mbase = Ms + 2 * t;
  // ---- pass 1 (rare): row maxima from the hi product
  float mx[RB][2];
#pragma unroll
  for (int b = 0; b < RB; ++b) mx[b][0] = mx[b][1] = 0.f;
  if (!EASY) {
#pragma unroll
    for (int b = 0; b < RB; ++b) mx[b][0] = mx[b][1] = -INFINITY;
    for (int li = 0; li < nlive; ++li) {
      const int nt = lidx[li];
      const uint2 kh = *reinterpret_cast<const uint2*>(kbase + nt * 128);
      const float2 mk = *reinterpret_cast<const float2*>(mbase + nt * 8);
#pragma unroll
      for (int b = 0; b < RB; ++b) {
        float s[4];
        mma_f16(s, qh[b][0], qh[b][1], qh[b][2], qh[b][3], kh.x, kh.y, 0.f, 0.f, 0.f, 0.f);
        mx[b][0] = fmaxf(mx[b][0], fmaxf(fmaf(s[0], sc_mul, mk.x), fmaf(s[1], sc_mul, mk.y)));
        mx[b][1] = fmaxf(mx[b][1], fmaxf(fmaf(s[2], sc_mul, mk.x), fmaf(s[3], sc_mul, mk.y)));
      }
    }
#pragma unroll
    for (int b = 0; b < RB; ++b)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[b][r] = fmaxf(mx[b][r], __shfl_xor_sync(0xffffffffu, mx[b][r], 1));
        mx[b][r] = fmaxf(mx[b][r], __shfl_xor_sync(0xffffffffu, mx[b][r], 2));
      }
  }
  // ---- pass 2: exact scores, softmax numerators, P.V
  float sum[RB][2], o[RB][2][4];
#pragma unroll
  for (int b = 0; b < RB; ++b) {
    sum[b][0] = sum[b][1] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) o[b][0][i] = o[b][1][i] = 0.f;
  }
#pragma unroll 2
  for (int li = 0; li < nlive; ++li) {
    const int nt = lidx[li];
    const uint4 kf = *reinterpret_cast<const uint4*>(kbase + nt * 128);   // kh(b0, b1), kl(b0, b1)
    const float2 mk = *reinterpret_cast<const float2*>(mbase + nt * 8);
    // P as A operand with its k index relabelled (k=t <-> key 2t, k=t+4 <-> key 2t+1): no shuffles needed.
    const float* vr = vbase + nt * 8;
    const float2 v0 = *reinterpret_cast<const float2*>(vr), v1 = *reinterpret_cast<const float2*>(vr + 8 * VS);
    float2 w0 = v0, w1 = v1;
    if (PV3X) {
      w0 = *reinterpret_cast<const float2*>(vr + (Vl - Vt));
      w1 = *reinterpret_cast<const float2*>(vr + (Vl - Vt) + 8 * VS);
    }
    float sc[RB][4];
    // small terms first; the RB chains are independent and interleave
#pragma unroll
    for (int b = 0; b < RB; ++b) mma_f16(sc[b], ql[b][0], ql[b][1], ql[b][2], ql[b][3], kf.x, kf.y, 0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int b = 0; b < RB; ++b)
      mma_f16(sc[b], qh[b][0], qh[b][1], qh[b][2], qh[b][3], kf.z, kf.w, sc[b][0], sc[b][1], sc[b][2], sc[b][3]);
#pragma unroll
    for (int b = 0; b < RB; ++b)
      mma_f16(sc[b], qh[b][0], qh[b][1], qh[b][2], qh[b][3], kf.x, kf.y, sc[b][0], sc[b][1], sc[b][2], sc[b][3]);
#pragma unroll
    for (int b = 0; b < RB; ++b) {
      // (rare two-pass path: the row maximum is folded into the additive mask term)
      const float ma0 = EASY ? mk.x : mk.x - mx[b][0], ma1 = EASY ? mk.y : mk.y - mx[b][0];
      const float mb0 = EASY ? mk.x : mk.x - mx[b][1], mb1 = EASY ? mk.y : mk.y - mx[b][1];
      const float p0 = ex2(fmaf(sc[b][0], sc_mul, ma0)), p1 = ex2(fmaf(sc[b][1], sc_mul, ma1));
      const float p2 = ex2(fmaf(sc[b][2], sc_mul, mb0)), p3 = ex2(fmaf(sc[b][3], sc_mul, mb1));
      sum[b][0] += p0 + p1;
      sum[b][1] += p2 + p3;
      if (PV3X) {
        // short sequences: 3xTF32 (with few keys there is no averaging of the operand rounding)
        const float h0 = tf32_rn(p0), h1 = tf32_rn(p2), h2 = tf32_rn(p1), h3 = tf32_rn(p3);
        const uint32_t a0 = fbits(h0), a1 = fbits(h1), a2 = fbits(h2), a3 = fbits(h3);
        const uint32_t l0 = fbits(p0 - h0), l1 = fbits(p2 - h1), l2 = fbits(p1 - h2), l3 = fbits(p3 - h3);
        mma_tf32(o[b][0], l0, l1, l2, l3, fbits(v0.x), fbits(v0.y));
        mma_tf32(o[b][1], l0, l1, l2, l3, fbits(v1.x), fbits(v1.y));
        mma_tf32(o[b][0], a0, a1, a2, a3, fbits(w0.x), fbits(w0.y));
        mma_tf32(o[b][1], a0, a1, a2, a3, fbits(w1.x), fbits(w1.y));
        mma_tf32(o[b][0], a0, a1, a2, a3, fbits(v0.x), fbits(v0.y));
        mma_tf32(o[b][1], a0, a1, a2, a3, fbits(v1.x), fbits(v1.y));
      } else {
        const uint32_t a0 = tfbits(p0), a1 = tfbits(p2), a2 = tfbits(p1), a3 = tfbits(p3);
        mma_tf32(o[b][0], a0, a1, a2, a3, fbits(v0.x), fbits(v0.y));
        mma_tf32(o[b][1], a0, a1, a2, a3, fbits(v1.x), fbits(v1.y));
      }
    }
  }
#pragma unroll
  for (int b = 0; b < RB; ++b) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float sm_ = sum[b][r];
      sm_ += __shfl_xor_sync(0xffffffffu, sm_, 1);
      sm_ += __shfl_xor_sync(0xffffffffu, sm_, 2);
      const float inv = 1.f / sm_;
      const int row = r0 + rstride * b + g + 8 * r;
      if (row < L) {
        float* dst = att_c + (size_t)row * 128 + hg + 2 * t;
        *reinterpret_cast<float2*>(dst) = make_float2(o[b][0][2 * r] * inv, o[b][0][2 * r + 1] * inv);
        *reinterpret_cast<float2*>(dst + 8) = make_float2(o[b][1][2 * r] * inv, o[b][1][2 * r + 1] * inv);
        if (t == 0 && lse_c) lse_c[(size_t)row * 8 + h] = mx[b][r] * LN2 + logf(sm_);   // natural-log units
      }
    }
  }
}

constexpr int RS = 20;   // raw (cp.async) row stride in floats: 16 of one head + 4, conflict-free float4 row reads

// cp.async of the head's 16 columns of rows [0, L) of matrix m (Q, K, V = column blocks 0, 1, 2 of qkv) into raw
__device__ __forceinline__ void prefetch_rows(const float* __restrict__ src, int ld, int L, float* raw) {
  for (int idx = threadIdx.x; idx < 4 * L; idx += blockDim.x) {
    const int r = idx >> 2, c4 = idx & 3;
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(raw + r * RS + c4 * 4));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + (size_t)r * ld + c4 * 4) : "memory");
  }
}
__device__ __forceinline__ void prefetch_mask(const float* __restrict__ src, int L, float* raw, int split, int gap) {
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(raw + j));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src + (j < split ? j : j + gap)) : "memory");
  }
}
__device__ __forceinline__ void raw_row(const float* raw, int r, int L, float4 (&x)[4]) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    x[q] = r < L ? *reinterpret_cast<const float4*>(raw + r * RS + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// Forward: persistent CTAs of 4 warps, each looping over (state, head) items (heads fastest, so the CTAs in flight
// share the 128-byte lines / DRAM pages of a state's rows: a head only touches 64 B of each 1536 B qkv row).
// Per item: the raw Q / K / V rows and the mask arrive by cp.async (issued one item ahead, so the global latency
// hides behind the previous item's math), are converted into the operand images, and the rest only touches shared
// memory. The warps take alternating 16-row query blocks, two at a time. Scores are the 3-MMA FP16 split product
// (fp32-level logits: the softmax amplifies score errors); P.V is plain TF32 with P, V rounded to nearest (3xTF32
// below 64 tokens, where there is no averaging of the rounding). The row maximum is only searched (pass 1) when
// the norm bound says the exponentials could leave fp32 range.
// NW warps per CTA: 2 up to 64 tokens (each warp two row blocks; twice the CTAs in flight), else 4.
template <bool PV3X, int NW>
__global__ void __launch_bounds__(NW * 32, NW == 5 ? 3 : 12 / NW) attn_fwd_kernel(const float* __restrict__ qkv, const float* __restrict__ mask,
                                                          int mS, int m0, int L, int nitems, float* __restrict__ att,
                                                          float* __restrict__ lse, int msplit, int mgap) {
  constexpr int FWD_WARPS = NW, FWD_THREADS = NW * 32;
  extern __shared__ float sm[];
  __shared__ unsigned nrm_bits[2][2];   // [item parity][Q, K]: largest squared row norm (non-negative floats order as uints)
  __shared__ int n_live_s;
  const int ntiles = (L + 7) >> 3;
  const int Lp = ntiles * 8;
  const int Lr = (Lp + 15) & ~15;    // query rows staged (A fragments read 16-row blocks)
  const int VS = vt_stride(Lp);
  uint32_t* Qi = reinterpret_cast<uint32_t*>(sm);   // [Lr][16]
  uint32_t* Ki = Qi + Lr * 16;                      // [Lp][16]
  float* Vt = reinterpret_cast<float*>(Ki + Lp * 16);
  float* Vl = Vt + 16 * VS;                       // lo part of V, only when PV3X
  float* Ms = Vl + (PV3X ? 16 * VS : 0);          // additive key mask: mask * -1e9, -inf beyond L
  int* lidx = reinterpret_cast<int*>(Ms + Lp);   // [ntiles] ordered list of live tiles
  float* raw = reinterpret_cast<float*>(lidx + ((ntiles + 3) & ~3));   // [3][Lp][RS] rows as they are in qkv (16 B aligned)
  float* Mraw = raw + 3 * Lp * RS;               // [Lp] the 0/1 mask row
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;

  auto prefetch = [&](int item) {
    const int c = item >> 3, hg = (item & 7) * 16;
    const float* base = qkv + (size_t)c * L * 384 + hg;
#pragma unroll
    for (int m = 0; m < 3; ++m) prefetch_rows(base + m * 128, 384, L, raw + m * Lp * RS);
    prefetch_mask(mask + (size_t)c * mS + m0, L, Mraw, msplit, mgap);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (tid < 4) (&nrm_bits[0][0])[tid] = 0u;
  int item = blockIdx.x;
  if (item < nitems) prefetch(item);
  for (int it = 0; item < nitems; item += gridDim.x, ++it) {
    const int c = item >> 3, h = item & 7, hg = h * 16, par = it & 1;
    ATTN_STAMP(0);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();   // raw rows of this item visible; every warp is done with the previous item's images
    ATTN_STAMP(1);
    for (int j = tid; j < Lp; j += FWD_THREADS) Ms[j] = j < L ? Mraw[j] * -BIG_NUMBER : -INFINITY;
    float qn = 0.f, kn = 0.f;
    for (int u = tid; u < 2 * Lp; u += FWD_THREADS) {
      float4 x[4];
      raw_row(raw + (u < Lp ? 0 : Lp * RS), u < Lp ? u : u - Lp, L, x);
      const float n2 = norm2(x);
      if (u < Lp) qn = fmaxf(qn, n2); else kn = fmaxf(kn, n2);
    }
    qn = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(qn)));
    kn = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(kn)));
    if (lane == 0) {
      atomicMax(&nrm_bits[par][0], __float_as_uint(qn));
      atomicMax(&nrm_bits[par][1], __float_as_uint(kn));
    }
    ATTN_STAMP(2);
    __syncthreads();
    ATTN_STAMP(3);
    if (warp == 0) build_live_list(Ms, lidx, &n_live_s, ntiles, lane, 0.f);
    if (tid < 2) nrm_bits[par ^ 1][tid] = 0u;   // for the next item (its atomics come after the next barrier)
    const float q2max = __uint_as_float(nrm_bits[par][0]), k2max = __uint_as_float(nrm_bits[par][1]);
    float q_down, q_up, k_down, k_up;
    pow2_scale(q2max, 14, q_down, q_up);
    pow2_scale(k2max, 14, k_down, k_up);
    for (int u = tid; u < Lr + 2 * Lp; u += FWD_THREADS) {
      float4 x[4];
      if (u < Lr) {
        raw_row(raw, u, L, x);
        store_split_row_frag(Qi, u, x, q_down);
      } else if (u < Lr + Lp) {
        raw_row(raw + Lp * RS, u - Lr, L, x);
        store_split_row(Ki + (u - Lr) * 16, x, k_down);
      } else {
        const int r = u - Lr - Lp;
        raw_row(raw + 2 * Lp * RS, r, L, x);
        const float v[16] = {x[0].x, x[0].y, x[0].z, x[0].w, x[1].x, x[1].y, x[1].z, x[1].w,
                             x[2].x, x[2].y, x[2].z, x[2].w, x[3].x, x[3].y, x[3].z, x[3].w};
#pragma unroll
        for (int dd = 0; dd < 16; ++dd) {
          const float hi = tf32_rn(v[dd]);
          Vt[dd * VS + r] = hi;
          if (PV3X) Vl[dd * VS + r] = v[dd] - hi;
        }
      }
    }
    ATTN_STAMP(4);
    __syncthreads();   // images ready; raw is free again
    if (item + (int)gridDim.x < nitems) prefetch(item + gridDim.x);
    ATTN_STAMP(5);
    const bool degenerate = n_live_s < 0;   // no unmasked key: uniform attention (see build_live_list)
    const int nlive = degenerate ? -n_live_s : n_live_s;
    const float sc_mul = degenerate ? 0.f : QSCALE * q_up * k_up;               // accumulator -> log2-unit score
    // NaN / inf norms take the two-pass path
    const bool easy = degenerate || q2max * k2max * (QSCALE * QSCALE) < EASY_BOUND * EASY_BOUND;
    float* att_c = att + (size_t)c * L * 128;
    float* lse_c = lse ? lse + (size_t)c * L * 8 : nullptr;
    // Warp w always runs on SM sub-partition w % 4, and the first warps of a CTA get the extra row block when the
    // block count is not a multiple of 4: rotate the assignment per item so the sub-partitions stay evenly loaded.
    const int rot = ((warp + it) % FWD_WARPS) * 16, stride = FWD_WARPS * 16;
    if (easy) {
      int r0 = rot;
      for (; r0 + stride < L; r0 += 2 * stride)
        fwd_row_blocks<2, PV3X, true>(r0, stride, L, hg, h, g, t, Qi, Ki, Vt, Vl, VS, Ms, lidx, nlive, sc_mul, att_c, lse_c);
      ATTN_STAMP(6);
      if (r0 < L) fwd_row_blocks<1, PV3X, true>(r0, stride, L, hg, h, g, t, Qi, Ki, Vt, Vl, VS, Ms, lidx, nlive, sc_mul, att_c, lse_c);
    } else {
      for (int r0 = rot; r0 < L; r0 += stride)
        fwd_row_blocks<1, PV3X, false>(r0, stride, L, hg, h, g, t, Qi, Ki, Vt, Vl, VS, Ms, lidx, nlive, sc_mul, att_c, lse_c);
    }
    ATTN_STAMP(7);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Backward. All five products run as FP16 m16n8k16 MMAs (half the MMA count of TF32 m16n8k8 for the same 11-bit
// operand precision): Q, K, V, dO are scaled by per-CTA powers of two (largest row norm -> <= 1) and rounded to
// FP16; P = exp(s - lse) <= 1 and dS' = P (dP' - D') = O(1) fit FP16 as they are. The power-of-two scales are
// undone exactly by the output constants.
//   phase A (16-row query blocks x 16-key steps): S = Q K^T, dP = dO V^T, dS, dQ += dS K
//   phase B (16-row key blocks x 16-query steps): S^T = K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q
// P is recomputed from the saved log-sum-exp; the row constants (-lse, -D) enter as the accumulator start values.
// Shared-memory operand images:
//   row images  Xh[row][8 words]   word 2t = dims (2t, 2t+1), word 2t+1 = dims (2t+8, 2t+9): the B fragment pair of
//                                  an 8-row tile is one LDS.64; A fragments take four scalar loads per row block
//   transposed  Xt[dim][TS words]  16-row group m at words 8m.., word 8m + 2t (+1) = rows (2t, 2t+1) (+8) packed:
//                                  the B fragment pair for a contraction over rows is one LDS.64
// The kernels are templated on the warp count NW = ceil(16-row blocks / 2), at most 5 (160 threads at the
// reference's 151 tokens): every warp then owns two row blocks and the staging needs one chunk.
__host__ __device__ inline int ts_stride(int Lr) {   // words per dim row of a transposed image, = 8 or 24 mod 32
  int n = Lr / 2;
  while (n % 32 != 8 && n % 32 != 24) n += 8;
  return n;
}
__device__ __forceinline__ uint32_t vld(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ float vldf(const float* p) { return *reinterpret_cast<const volatile float*>(p); }
// A fragment quad of rows (ra, ra + 8) from a row image. Scalar volatile loads: each value is its own register, so
// the quad is allocated once as four consecutive registers (vector loads make ptxas re-shuffle them at every MMA).
__device__ __forceinline__ void a_quad(const uint32_t* img, int ra, int t, uint32_t (&a)[4]) {
  a[0] = vld(img + ra * 8 + 2 * t);
  a[1] = vld(img + (ra + 8) * 8 + 2 * t);
  a[2] = vld(img + ra * 8 + 2 * t + 1);
  a[3] = vld(img + (ra + 8) * 8 + 2 * t + 1);
}
__device__ __forceinline__ void mma_f16_acc(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct BwdSmem {
  uint32_t *Qh, *Kh, *Vh, *Gh;   // row images
  uint32_t *Qt, *Kt, *Gt;        // transposed images
  float *Ms, *Lm, *Dm;           // [Lr] additive key mask; -lse2 / cS; -D'
  float4 *Lq, *Dq;               // [Lr / 8][4] the same two row constants as ready-made accumulator quads per (tile, t)
  int* lidx;                     // live 16-key groups
  int TS;
};
struct BwdScales {
  float cS;   // score accumulator -> log2 units
  float cQ, cK, cV;
};

// phase A for RB query blocks (rows r0 + rstride b) of one warp
template <int RB>
__device__ __forceinline__ void bwd_phase_a(int r0, int rstride, int L, int hg, int g, int t, const BwdSmem& S,
                                            int nlive, const BwdScales& sc, float* __restrict__ dbase) {
  uint32_t qa[RB][4], ga[RB][4];
  float cs[RB][4], cd[RB][4];
#pragma unroll
  for (int b = 0; b < RB; ++b) {
    const int ra = r0 + rstride * b + g;
    a_quad(S.Qh, ra, t, qa[b]);
    a_quad(S.Gh, ra, t, ga[b]);
    cs[b][0] = vldf(S.Lm + ra), cs[b][1] = vldf(S.Lm + ra), cs[b][2] = vldf(S.Lm + ra + 8), cs[b][3] = vldf(S.Lm + ra + 8);
    cd[b][0] = vldf(S.Dm + ra), cd[b][1] = vldf(S.Dm + ra), cd[b][2] = vldf(S.Dm + ra + 8), cd[b][3] = vldf(S.Dm + ra + 8);
  }
  float dq[RB][2][4];
#pragma unroll
  for (int b = 0; b < RB; ++b)
#pragma unroll
    for (int i = 0; i < 4; ++i) dq[b][0][i] = dq[b][1][i] = 0.f;
  const uint32_t* kh = S.Kh + g * 8 + 2 * t;
  const uint32_t* vh = S.Vh + g * 8 + 2 * t;
  const uint32_t* kt = S.Kt + g * S.TS + 2 * t;
  const float* ms = S.Ms + 2 * t;
  for (int li = 0; li < nlive; ++li) {   // live key groups only (masked keys: P == 0 exactly)
    const int m = S.lidx[li];
    const uint2 kb0 = *reinterpret_cast<const uint2*>(kh + m * 128), kb1 = *reinterpret_cast<const uint2*>(kh + m * 128 + 64);
    const uint2 vb0 = *reinterpret_cast<const uint2*>(vh + m * 128), vb1 = *reinterpret_cast<const uint2*>(vh + m * 128 + 64);
    const float2 mk0 = *reinterpret_cast<const float2*>(ms + m * 16), mk1 = *reinterpret_cast<const float2*>(ms + m * 16 + 8);
    const uint2 kt0 = *reinterpret_cast<const uint2*>(kt + m * 8), kt1 = *reinterpret_cast<const uint2*>(kt + 8 * S.TS + m * 8);
#pragma unroll
    for (int b = 0; b < RB; ++b) {
      float s0[4], s1[4], p0[4], p1[4];
      mma_f16(s0, qa[b][0], qa[b][1], qa[b][2], qa[b][3], kb0.x, kb0.y, cs[b][0], cs[b][1], cs[b][2], cs[b][3]);
      mma_f16(s1, qa[b][0], qa[b][1], qa[b][2], qa[b][3], kb1.x, kb1.y, cs[b][0], cs[b][1], cs[b][2], cs[b][3]);
      mma_f16(p0, ga[b][0], ga[b][1], ga[b][2], ga[b][3], vb0.x, vb0.y, cd[b][0], cd[b][1], cd[b][2], cd[b][3]);
      mma_f16(p1, ga[b][0], ga[b][1], ga[b][2], ga[b][3], vb1.x, vb1.y, cd[b][0], cd[b][1], cd[b][2], cd[b][3]);
      // dS' as the A operand of the contraction over the 16 keys: C layout of the two 8-key tiles == A layout
      uint32_t a[4];
      a[0] = pack_h2(ex2(fmaf(s0[0], sc.cS, mk0.x)) * p0[0], ex2(fmaf(s0[1], sc.cS, mk0.y)) * p0[1]);
      a[1] = pack_h2(ex2(fmaf(s0[2], sc.cS, mk0.x)) * p0[2], ex2(fmaf(s0[3], sc.cS, mk0.y)) * p0[3]);
      a[2] = pack_h2(ex2(fmaf(s1[0], sc.cS, mk1.x)) * p1[0], ex2(fmaf(s1[1], sc.cS, mk1.y)) * p1[1]);
      a[3] = pack_h2(ex2(fmaf(s1[2], sc.cS, mk1.x)) * p1[2], ex2(fmaf(s1[3], sc.cS, mk1.y)) * p1[3]);
      mma_f16_acc(dq[b][0], a, kt0.x, kt0.y);
      mma_f16_acc(dq[b][1], a, kt1.x, kt1.y);
    }
  }
#pragma unroll
  for (int b = 0; b < RB; ++b)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = r0 + rstride * b + g + 8 * r;
      if (row < L) {
        float* dst = dbase + (size_t)row * 384 + hg + 2 * t;
        *reinterpret_cast<float2*>(dst) = make_float2(tf32_rn(dq[b][0][2 * r] * sc.cQ), tf32_rn(dq[b][0][2 * r + 1] * sc.cQ));
        *reinterpret_cast<float2*>(dst + 8) = make_float2(tf32_rn(dq[b][1][2 * r] * sc.cQ), tf32_rn(dq[b][1][2 * r + 1] * sc.cQ));
      }
    }
}

// phase B for RB key blocks (rows j0 + rstride b) of one warp; tile rows are KEYS, columns are QUERIES
template <int RB>
__device__ __forceinline__ void bwd_phase_b(int j0, int rstride, int L, int Lr, int hg, int g, int t, const BwdSmem& S,
                                            unsigned long long live_bits, const BwdScales& sc, float* __restrict__ dbase) {
  uint32_t ka[RB][4], va[RB][4];
  float ma[RB], mb[RB];
  bool any_live = false;
#pragma unroll
  for (int b = 0; b < RB; ++b) {
    const int ja = j0 + rstride * b + g;
    a_quad(S.Kh, ja, t, ka[b]);
    a_quad(S.Vh, ja, t, va[b]);
    ma[b] = S.Ms[ja], mb[b] = S.Ms[ja + 8];
    const int grp = (j0 + rstride * b) >> 4;
    any_live |= grp >= 64 || ((live_bits >> grp) & 1ull);
  }
  float dk[RB][2][4], dv[RB][2][4];
#pragma unroll
  for (int b = 0; b < RB; ++b)
#pragma unroll
    for (int i = 0; i < 4; ++i) dk[b][0][i] = dk[b][1][i] = dv[b][0][i] = dv[b][1][i] = 0.f;
  const uint32_t* qh = S.Qh + g * 8 + 2 * t;
  const uint32_t* gh = S.Gh + g * 8 + 2 * t;
  const uint32_t* qt = S.Qt + g * S.TS + 2 * t;
  const uint32_t* gt = S.Gt + g * S.TS + 2 * t;
  // fully masked key blocks: dK = dV = 0 (still written below)
  for (int m = 0; m < (any_live ? (Lr >> 4) : 0); ++m) {
    const uint2 qb0 = *reinterpret_cast<const uint2*>(qh + m * 128), qb1 = *reinterpret_cast<const uint2*>(qh + m * 128 + 64);
    const uint2 gb0 = *reinterpret_cast<const uint2*>(gh + m * 128), gb1 = *reinterpret_cast<const uint2*>(gh + m * 128 + 64);
    const float4 l0 = S.Lq[(2 * m) * 4 + t], l1 = S.Lq[(2 * m + 1) * 4 + t];
    const float4 d0 = S.Dq[(2 * m) * 4 + t], d1 = S.Dq[(2 * m + 1) * 4 + t];
    const uint2 qt0 = *reinterpret_cast<const uint2*>(qt + m * 8), qt1 = *reinterpret_cast<const uint2*>(qt + 8 * S.TS + m * 8);
    const uint2 gt0 = *reinterpret_cast<const uint2*>(gt + m * 8), gt1 = *reinterpret_cast<const uint2*>(gt + 8 * S.TS + m * 8);
#pragma unroll
    for (int b = 0; b < RB; ++b) {
      float s0[4], s1[4], p0[4], p1[4];
      mma_f16(s0, ka[b][0], ka[b][1], ka[b][2], ka[b][3], qb0.x, qb0.y, l0.x, l0.y, l0.z, l0.w);
      mma_f16(s1, ka[b][0], ka[b][1], ka[b][2], ka[b][3], qb1.x, qb1.y, l1.x, l1.y, l1.z, l1.w);
      mma_f16(p0, va[b][0], va[b][1], va[b][2], va[b][3], gb0.x, gb0.y, d0.x, d0.y, d0.z, d0.w);
      mma_f16(p1, va[b][0], va[b][1], va[b][2], va[b][3], gb1.x, gb1.y, d1.x, d1.y, d1.z, d1.w);
      const float e00 = ex2(fmaf(s0[0], sc.cS, ma[b])), e01 = ex2(fmaf(s0[1], sc.cS, ma[b]));
      const float e02 = ex2(fmaf(s0[2], sc.cS, mb[b])), e03 = ex2(fmaf(s0[3], sc.cS, mb[b]));
      const float e10 = ex2(fmaf(s1[0], sc.cS, ma[b])), e11 = ex2(fmaf(s1[1], sc.cS, ma[b]));
      const float e12 = ex2(fmaf(s1[2], sc.cS, mb[b])), e13 = ex2(fmaf(s1[3], sc.cS, mb[b]));
      uint32_t pa[4], sa[4];
      pa[0] = pack_h2(e00, e01), pa[1] = pack_h2(e02, e03), pa[2] = pack_h2(e10, e11), pa[3] = pack_h2(e12, e13);
      sa[0] = pack_h2(e00 * p0[0], e01 * p0[1]), sa[1] = pack_h2(e02 * p0[2], e03 * p0[3]);
      sa[2] = pack_h2(e10 * p1[0], e11 * p1[1]), sa[3] = pack_h2(e12 * p1[2], e13 * p1[3]);
      mma_f16_acc(dv[b][0], pa, gt0.x, gt0.y);
      mma_f16_acc(dv[b][1], pa, gt1.x, gt1.y);
      mma_f16_acc(dk[b][0], sa, qt0.x, qt0.y);
      mma_f16_acc(dk[b][1], sa, qt1.x, qt1.y);
    }
  }
#pragma unroll
  for (int b = 0; b < RB; ++b)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = j0 + rstride * b + g + 8 * r;
      if (row < L) {
        float* dst = dbase + (size_t)row * 384 + 128 + hg + 2 * t;
        *reinterpret_cast<float2*>(dst) = make_float2(tf32_rn(dk[b][0][2 * r] * sc.cK), tf32_rn(dk[b][0][2 * r + 1] * sc.cK));
        *reinterpret_cast<float2*>(dst + 8) = make_float2(tf32_rn(dk[b][1][2 * r] * sc.cK), tf32_rn(dk[b][1][2 * r + 1] * sc.cK));
        *reinterpret_cast<float2*>(dst + 128) = make_float2(tf32_rn(dv[b][0][2 * r] * sc.cV), tf32_rn(dv[b][0][2 * r + 1] * sc.cV));
        *reinterpret_cast<float2*>(dst + 136) = make_float2(tf32_rn(dv[b][1][2 * r] * sc.cV), tf32_rn(dv[b][1][2 * r + 1] * sc.cV));
      }
    }
}

// Staging: 4 lanes share a row (16 B = 4 dims each), so a warp-wide 16 B load covers 8 rows x 64 B: every sector
// it touches is used in full. One chunk = 4 passes of 40 rows x the five matrices = 20 loads in flight per thread.
struct BwdChunk {
  float4 q[4], k[4], v[4], g[4], o[4];
  float lse[4];
};
template <int RPP>   // rows per pass = threads / 4
__device__ __forceinline__ void bwd_load_chunk(const float* __restrict__ base, const float* __restrict__ obase,
                                               const float* __restrict__ dobase, const float* __restrict__ lse_c, int row0,
                                               int L, int hg, int h, BwdChunk& d) {
  const int sub = threadIdx.x & 3, rloc = threadIdx.x >> 2;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int r = row0 + rloc + RPP * p;
    const bool ok = r < L;
    const int rr = min(r, L - 1);
    const float* qp = base + (size_t)rr * 384 + hg + sub * 4;
    d.q[p] = ok ? *reinterpret_cast<const float4*>(qp) : z;
    d.k[p] = ok ? *reinterpret_cast<const float4*>(qp + 128) : z;
    d.v[p] = ok ? *reinterpret_cast<const float4*>(qp + 256) : z;
    d.g[p] = ok ? *reinterpret_cast<const float4*>(dobase + (size_t)rr * 128 + hg + sub * 4) : z;
    d.o[p] = ok ? *reinterpret_cast<const float4*>(obase + (size_t)rr * 128 + hg + sub * 4) : z;
    d.lse[p] = (ok && sub == 0) ? lse_c[(size_t)rr * 8 + h] : 0.f;
  }
}
__device__ __forceinline__ float quad_sum(float x) {
  x += __shfl_xor_sync(0xffffffffu, x, 1);
  x += __shfl_xor_sync(0xffffffffu, x, 2);
  return x;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
// FP16 row image + transposed image of the 4 dims a lane holds of row r; every lane of the warp calls this (rows
// 2i and 2i + 1 meet by shuffle: they share the words of the transposed image)
__device__ __forceinline__ void bwd_store_images(const float4& v, float scale, int r, int Lr, uint32_t* rowimg,
                                                 uint32_t* timg, int TS) {
  const int sub = threadIdx.x & 3;
  const float x0 = v.x * scale, x1 = v.y * scale, x2 = v.z * scale, x3 = v.w * scale;
  const bool act = r < Lr;
  if (act) {
    uint32_t* w = rowimg + r * 8 + (sub < 2 ? 4 * sub : 4 * (sub - 2) + 1);
    w[0] = pack_h2(x0, x1);
    w[2] = pack_h2(x2, x3);
  }
  if (timg) {
    const int odd = r & 1;
    const float o0 = __shfl_xor_sync(0xffffffffu, odd ? x0 : x2, 4), o1 = __shfl_xor_sync(0xffffffffu, odd ? x1 : x3, 4);
    if (act) {
      uint32_t* w = timg + (4 * sub + 2 * odd) * TS + 8 * (r >> 4) + 2 * ((r & 7) >> 1) + ((r & 15) >> 3);
      w[0] = odd ? pack_h2(o0, x2) : pack_h2(x0, o0);
      w[TS] = odd ? pack_h2(o1, x3) : pack_h2(x1, o1);
    }
  }
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, (NW == 5 ? 3 : 16 / NW)) attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ att,
                                                                   const float* __restrict__ lse, const float* __restrict__ datt,
                                                                   const float* __restrict__ mask, int mS, int m0, int L,
                                                                   int pf_dist, float* __restrict__ dqkv, int msplit, int mgap) {
  extern __shared__ float sm[];
  __shared__ unsigned nrm_bits[4];          // largest squared row norm of Q, K, V, dO
  __shared__ unsigned long long live_bits_s;
  __shared__ int n_live_s;
  const int Lr = (L + 15) & ~15, ngrp = Lr >> 4;
  BwdSmem S;
  S.TS = ts_stride(Lr);
  S.Qh = reinterpret_cast<uint32_t*>(sm);
  S.Kh = S.Qh + Lr * 8;
  S.Vh = S.Kh + Lr * 8;
  S.Gh = S.Vh + Lr * 8;
  S.Qt = S.Gh + Lr * 8;
  S.Kt = S.Qt + 16 * S.TS;
  S.Gt = S.Kt + 16 * S.TS;
  S.Lq = reinterpret_cast<float4*>(S.Gt + 16 * S.TS);
  S.Dq = S.Lq + (Lr >> 3) * 4;
  S.Ms = reinterpret_cast<float*>(S.Dq + (Lr >> 3) * 4);
  S.Lm = S.Ms + Lr;
  S.Dm = S.Lm + Lr;
  S.lidx = reinterpret_cast<int*>(S.Dm + Lr);
  // heads are the fastest-varying block index: the 8 CTAs of a state run side by side and share the 128-byte
  // lines / DRAM pages of its rows (a head only touches 64 B of each 1536 B qkv row)
  const int c = blockIdx.x >> 3, h = blockIdx.x & 7, hg = h * 16;
  const float* base = qkv + (size_t)c * L * 384;
  const float* obase = att + (size_t)c * L * 128;
  const float* dobase = datt + (size_t)c * L * 128;
  const float* lse_c = lse + (size_t)c * L * 8;
  float* dbase = dqkv + (size_t)c * L * 384;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  constexpr int BWD_THREADS = NW * 32, BWD_WARPS = NW, BWD_RPP = NW * 8;
  const bool single = Lr <= 4 * BWD_RPP;   // one chunk: the rows wait in registers for the scales
  ATTN_STAMP(0);
  BwdChunk d;
  float nq = 0.f, nk = 0.f, nv = 0.f, ng = 0.f;
  if (single) bwd_load_chunk<BWD_RPP>(base, obase, dobase, lse_c, 0, L, hg, h, d);
  if (tid < 4) nrm_bits[tid] = 0u;
  for (int j = tid; j < Lr; j += BWD_THREADS)
    S.Ms[j] = j < L ? mask[(size_t)c * mS + m0 + (j < msplit ? j : j + mgap)] * -BIG_NUMBER : -INFINITY;
  for (int row0 = 0; row0 < Lr; row0 += 4 * BWD_RPP) {
    if (!single) bwd_load_chunk<BWD_RPP>(base, obase, dobase, lse_c, row0, L, hg, h, d);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      nq = fmaxf(nq, dot4(d.q[p], d.q[p])), nk = fmaxf(nk, dot4(d.k[p], d.k[p]));
      nv = fmaxf(nv, dot4(d.v[p], d.v[p])), ng = fmaxf(ng, dot4(d.g[p], d.g[p]));
    }
  }
  // (per-lane maxima of quarter-row sums of squares: 4 x the largest is an upper bound of the largest row norm^2)
  nq = 4.f * __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(nq)));
  nk = 4.f * __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(nk)));
  nv = 4.f * __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(nv)));
  ng = 4.f * __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(ng)));
  __syncthreads();   // nrm_bits zeroed, Ms written
  if (lane == 0) {
    atomicMax(&nrm_bits[0], __float_as_uint(nq));
    atomicMax(&nrm_bits[1], __float_as_uint(nk));
    atomicMax(&nrm_bits[2], __float_as_uint(nv));
    atomicMax(&nrm_bits[3], __float_as_uint(ng));
  }
  if (warp == 0) {
    // ordered list + bit mask of the live 16-key groups (see build_live_list for the exactness argument)
    int n = 0;
    unsigned long long bits = 0ull;
    for (int m0_ = 0; m0_ < ngrp; m0_ += 32) {
      const int m = m0_ + lane;
      bool l = false;
      if (m < ngrp)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 x = *reinterpret_cast<const float4*>(S.Ms + m * 16 + q * 4);
          l = l || x.x == 0.f || x.y == 0.f || x.z == 0.f || x.w == 0.f;
        }
      const unsigned bal = __ballot_sync(0xffffffffu, l);
      if (l) S.lidx[n + __popc(bal & ((1u << lane) - 1u))] = m;
      n += __popc(bal);
      if (m0_ < 64) bits |= (unsigned long long)bal << m0_;
    }
    if (n == 0) {
      // nothing unmasked: uniform attention P = 1 / L (see build_live_list); the score scale is zeroed below and
      // the log-sum-exp log2(L), common to all rows, takes the place of the mask term
      for (int m = lane; m < ngrp; m += 32) S.lidx[m] = m;
      const float fill = -log2f((float)L);
      for (int j = lane; j < L; j += 32) S.Ms[j] = fill;
      n = -ngrp;
      bits = ~0ull;
    }
    if (lane == 0) n_live_s = n, live_bits_s = bits;
  }
  __syncthreads();
  ATTN_STAMP(1);
  float q_dn, q_up, k_dn, k_up, v_dn, v_up, g_dn, g_up;
  // Q, K only feed fp32 accumulators: use FP16's range. dS' = P (dO' V'^T - D') is itself an FP16 operand and is
  // bounded by 2 |dO'| |V'| <= 2^15.
  pow2_scale(__uint_as_float(nrm_bits[0]), 14, q_dn, q_up);
  pow2_scale(__uint_as_float(nrm_bits[1]), 14, k_dn, k_up);
  pow2_scale(__uint_as_float(nrm_bits[2]), 7, v_dn, v_up);
  pow2_scale(__uint_as_float(nrm_bits[3]), 7, g_dn, g_up);
  const bool degenerate = n_live_s < 0;
  BwdScales sc;
  sc.cS = degenerate ? 0.f : QSCALE * q_up * k_up;
  sc.cQ = 0.25f * g_up * v_up * k_up;
  sc.cK = 0.25f * g_up * v_up * q_up;
  sc.cV = g_up;
  const float inv_cS = degenerate ? 0.f : 1.f / sc.cS, d_scale = g_dn * v_dn;
  for (int row0 = 0; row0 < Lr; row0 += 4 * BWD_RPP) {
    if (!single) bwd_load_chunk<BWD_RPP>(base, obase, dobase, lse_c, row0, L, hg, h, d);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int r = row0 + (tid >> 2) + BWD_RPP * p;
      bwd_store_images(d.q[p], q_dn, r, Lr, S.Qh, S.Qt, S.TS);
      bwd_store_images(d.k[p], k_dn, r, Lr, S.Kh, S.Kt, S.TS);
      bwd_store_images(d.v[p], v_dn, r, Lr, S.Vh, nullptr, S.TS);
      bwd_store_images(d.g[p], g_dn, r, Lr, S.Gh, S.Gt, S.TS);
      const float dd = quad_sum(dot4(d.g[p], d.o[p]));   // D = sum_d dO O of the row
      if (r < Lr && (tid & 3) == 0) {
        // padded queries: 2^(score - inf) == 0 (with the score scale zeroed, inf * 0 would be a NaN: their Q and dO
        // rows are zero, a finite start value contributes nothing either)
        const float lm = r < L ? -(d.lse[p] * LOG2E) * inv_cS : (degenerate ? 0.f : -INFINITY), dm = -dd * d_scale;
        S.Lm[r] = lm;
        S.Dm[r] = dm;
        // accumulator quads of phase B: column (query) r sits at positions (r & 1) and (r & 1) + 2 of quad [r / 8][(r & 7) / 2]
        float* lq = reinterpret_cast<float*>(S.Lq + (r >> 3) * 4 + ((r & 7) >> 1)) + (r & 1);
        float* dq = reinterpret_cast<float*>(S.Dq + (r >> 3) * 4 + ((r & 7) >> 1)) + (r & 1);
        lq[0] = lm, lq[2] = lm;
        dq[0] = dm, dq[2] = dm;
      }
    }
  }
  __syncthreads();
  ATTN_STAMP(2);
  // pull the rows of the item one wave of CTAs ahead into L2 while this one computes
  {
    const int nxt = blockIdx.x + pf_dist;
    if (nxt < (int)gridDim.x) {
      const int c2 = nxt >> 3, hg2 = (nxt & 7) * 16;
      const float* b2 = qkv + (size_t)c2 * L * 384 + hg2;
      const float* o2 = att + (size_t)c2 * L * 128 + hg2;
      const float* g2 = datt + (size_t)c2 * L * 128 + hg2;
      for (int r = tid; r < L; r += BWD_THREADS) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(b2 + (size_t)r * 384));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(b2 + (size_t)r * 384 + 128));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(b2 + (size_t)r * 384 + 256));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(o2 + (size_t)r * 128));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(g2 + (size_t)r * 128));
      }
    }
  }
  const int nlive = degenerate ? -n_live_s : n_live_s;
  const unsigned long long live_bits = live_bits_s;
  // Warp w always runs on SM sub-partition w % 4: rotate the block assignment per CTA so the sub-partitions stay
  // evenly loaded when the block count is not a multiple of the warp count.
  const int rot = ((warp + ((blockIdx.x * 0x9E3779B1u) >> 29)) % BWD_WARPS) * 16, stride = BWD_WARPS * 16;
  {
    int r0 = rot;
    for (; r0 + stride < L; r0 += 2 * stride) bwd_phase_a<2>(r0, stride, L, hg, g, t, S, nlive, sc, dbase);
    if (r0 < L) bwd_phase_a<1>(r0, stride, L, hg, g, t, S, nlive, sc, dbase);
  }
  ATTN_STAMP(3);
  {
    int j0 = rot;
    for (; j0 + stride < L; j0 += 2 * stride) bwd_phase_b<2>(j0, stride, L, Lr, hg, g, t, S, live_bits, sc, dbase);
    if (j0 < L) bwd_phase_b<1>(j0, stride, L, Lr, hg, g, t, S, live_bits, sc, dbase);
  }
  ATTN_STAMP(4);
}

}  // namespace

int attention_fwd(const float* qkv, const float* mask, int mS, int m0, int C, int L, float* att, float* lse,
                  cudaStream_t st, int mask_split, int mask_gap) {
  if (C <= 0) return TPC_OK;
  if (L < 1 || L > 512) return TPC_ERR_SHAPE;
  const int Lp = ((L + 7) / 8) * 8, Lr = (Lp + 15) & ~15, nt = Lp / 8;
  const size_t smem = (size_t)((Lr + Lp) * 16 + (L < 64 ? 32 : 16) * vt_stride(Lp) + Lp + ((nt + 3) & ~3) + 3 * Lp * RS + Lp) *
                      sizeof(float);
  // static smem of the kernel counts against the 227 KB limit too: ask for what is needed
  static size_t attr_smem[4] = {0, 0, 0, 0};
  static int ctas_per_sm[4] = {0, 0, 0, 0}, num_sms = 0;
  static size_t attr_smem5 = 0;
  const int nblk = Lr / 16;
  const int v = L < 64 ? 0 : (L == 64 ? 1 : (nblk == 9 || nblk == 10 ? 3 : 2)), nw = v == 3 ? 5 : (v == 2 ? 4 : 2);
  auto kern = v == 0 ? attn_fwd_kernel<true, 2>
                     : (v == 1 ? attn_fwd_kernel<false, 2> : (v == 2 ? attn_fwd_kernel<false, 4> : attn_fwd_kernel<false, 5>));
  (void)attr_smem5;
  if (smem > attr_smem[v] || ctas_per_sm[v] == 0) {
    TPC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem[v] = smem;
    int dev = 0;
    TPC_CHECK_CUDA(cudaGetDevice(&dev));
    TPC_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    TPC_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm[v], kern, nw * 32, smem));
    if (ctas_per_sm[v] < 1) return TPC_ERR_SHAPE;
  }
  const int nitems = C * 8;
  const int grid = nitems < num_sms * ctas_per_sm[v] ? nitems : num_sms * ctas_per_sm[v];
  kern<<<grid, nw * 32, smem, st>>>(qkv, mask, mS, m0, L, nitems, att, lse, mask_split, mask_gap);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int attention_bwd(const float* qkv, const float* att, const float* lse, const float* datt, const float* mask, int mS,
                  int m0, int C, int L, float* dqkv, cudaStream_t st, int mask_split, int mask_gap) {
  if (C <= 0) return TPC_OK;
  if (L < 1 || L > 512) return TPC_ERR_SHAPE;
  const int Lr = (L + 15) & ~15;
  const size_t smem = (size_t)(4 * Lr * 8 + 3 * 16 * ts_stride(Lr) + 2 * (Lr / 8) * 16 + 3 * Lr + Lr / 16 + 8) * sizeof(float);
  const int nblk = Lr / 16, nw = nblk >= 9 ? 5 : (nblk + 1) / 2, v = nw - 1;
  using Kern = void (*)(const float*, const float*, const float*, const float*, const float*, int, int, int, int, float*,
                        int, int);
  static const Kern kerns[5] = {attn_bwd_kernel<1>, attn_bwd_kernel<2>, attn_bwd_kernel<3>, attn_bwd_kernel<4>,
                                attn_bwd_kernel<5>};
  static size_t attr_smem[5] = {0, 0, 0, 0, 0};
  static int resident[5] = {0, 0, 0, 0, 0};   // CTAs in flight on the whole GPU = the L2 prefetch distance
  if (smem > attr_smem[v] || resident[v] == 0) {
    TPC_CHECK_CUDA(cudaFuncSetAttribute(kerns[v], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem[v] = smem;
    int dev = 0, sms = 0, per_sm = 0;
    TPC_CHECK_CUDA(cudaGetDevice(&dev));
    TPC_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    TPC_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kerns[v], nw * 32, smem));
    if (per_sm < 1) return TPC_ERR_SHAPE;
    resident[v] = sms * per_sm;
  }
  kerns[v]<<<C * 8, nw * 32, smem, st>>>(qkv, att, lse, datt, mask, mS, m0, L, resident[v], dqkv, mask_split, mask_gap);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

}  // namespace tpc
