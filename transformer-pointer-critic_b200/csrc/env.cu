// On-device ResourceV3 / KnapsackV2 environment: instance generation, feasibility mask, step + rewards,
// categorical sampling. Bit-exact restatement of the reference's fp32 NumPy arithmetic
// (environment/custom/resource_v3/env.py:115-203,387-430; reward.py:23-166; misc/utils.py:19-42;
//  environment/custom/knapsack_v2/env.py:102-185,384-408; knapsack_v2/misc/utils.py:5-32).
// Every float op is an explicit round-to-nearest intrinsic: no FMA contraction, same rounding sequence as NumPy.
#include "env.cuh"
#include "philox.cuh"

namespace tpc {

namespace {

// int / normalization_factor as the reference computes it: int32 -> truediv (float64) -> cast float32
// (env.py:207-215); equal to the fp32 division for the value range used.
__device__ __forceinline__ float norm_val(int k, int norm) { return (float)((double)k / (double)norm); }

__global__ void gen_profiles_kernel(tpc_env_config e, uint64_t seed, float* __restrict__ profiles) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= e.num_profiles) return;
  Stream s = make_stream(seed, i, 0, 0);
  if (e.kind == TPC_ENV_KNAPSACK_V2) {
    profiles[i * 2 + 0] = norm_val(s.randint(0, e.req_min_val, e.req_max_val), e.normalization_factor);
    profiles[i * 2 + 1] = norm_val(s.randint(1, e.val_min_val, e.val_max_val), e.normalization_factor);
  } else {
    for (int f = 0; f < e.num_features; ++f)
      profiles[i * e.num_features + f] = norm_val(s.randint(f, e.req_min_val, e.req_max_val), e.normalization_factor);
  }
}

constexpr int MAX_RULES = 512;

__global__ void reset_kernel(tpc_env_config e, uint64_t seed, int64_t id0, int64_t epoch,
                             const float* __restrict__ profiles, int B, int N, int R, float* __restrict__ batch,
                             float* __restrict__ bin_mask, float* __restrict__ mha_mask, float* __restrict__ is_empty) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int S = N + R, F = e.num_features;
  float* st = batch + (size_t)b * S * F;
  const int64_t id = id0 + b;
  // nodes (env.py:228-239): row 0 = EOS
  Stream sn = make_stream(seed, id, epoch, 1);
  for (int f = 0; f < F; ++f) st[f] = e.eos_code;
  for (int n = 1; n < N; ++n) {
    if (e.kind == TPC_ENV_KNAPSACK_V2) {
      st[n * F + 0] = norm_val(sn.randint(n, e.node_min_val, e.node_max_val), e.normalization_factor);
      st[n * F + 1] = 0.f;
    } else {
      for (int f = 0; f < F; ++f)
        st[n * F + f] = norm_val(sn.randint(n * F + f, e.node_min_val, e.node_max_val), e.normalization_factor);
    }
  }
  // rules (env.py:241-257): R distinct profiles in random order, or i.i.d. values when generated on the fly
  Stream sr = make_stream(seed, id, epoch, 2);
  if (e.num_profiles > 0) {
    int chosen[MAX_RULES];
    const int P = e.num_profiles;
    // Floyd's subset sampling, then a Fisher-Yates pass for a uniformly random order
    for (int i = 0; i < R; ++i) {
      const int top = P - R + i;
      int tval = sr.randint(i, 0, top + 1);
      bool dup = false;
      for (int k = 0; k < i; ++k) dup |= (chosen[k] == tval);
      chosen[i] = dup ? top : tval;
    }
    for (int i = R - 1; i > 0; --i) {
      const int j = sr.randint(R + i, 0, i + 1);
      const int tmp = chosen[i]; chosen[i] = chosen[j]; chosen[j] = tmp;
    }
    for (int r = 0; r < R; ++r)
      for (int f = 0; f < F; ++f) st[(N + r) * F + f] = profiles[chosen[r] * F + f];
  } else {
    for (int r = 0; r < R; ++r) {
      if (e.kind == TPC_ENV_KNAPSACK_V2) {
        st[(N + r) * F + 0] = norm_val(sr.randint(r * 2, e.req_min_val, e.req_max_val), e.normalization_factor);
        st[(N + r) * F + 1] = norm_val(sr.randint(r * 2 + 1, e.val_min_val, e.val_max_val), e.normalization_factor);
      } else {
        for (int f = 0; f < F; ++f)
          st[(N + r) * F + f] = norm_val(sr.randint(r * F + f, e.req_min_val, e.req_max_val), e.normalization_factor);
      }
    }
  }
  // masks (env.py:265-288)
  for (int j = 0; j < S; ++j) {
    bin_mask[(size_t)b * S + j] = j < N ? 0.f : 1.f;
    mha_mask[(size_t)b * S + j] = 0.f;
    if (is_empty) is_empty[(size_t)b * S + j] = 0.f;
  }
  if (is_empty) is_empty[(size_t)b * S] = e.eos_code;  // env.py:87-91
}

__global__ void feasible_mask_kernel(tpc_env_config e, const float* __restrict__ state, int ld,
                                     const float* __restrict__ dec, const float* __restrict__ bin_mask, int B, int S,
                                     float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * S) return;
  const int b = idx / S, j = idx % S;
  const float* row = state + (size_t)idx * ld;
  const float* d = dec + (size_t)b * e.num_features;
  float after;
  if (e.kind == TPC_ENV_KNAPSACK_V2) {
    // capacity - (load + weight) < 0, unrounded fp32 (knapsack_v2/env.py:393-399)
    const float rem = __fsub_rn(row[0], __fadd_rn(row[1], d[0]));
    after = rem < 0.f ? 1.f : 0.f;
  } else {
    float dom = INFINITY;
    for (int f = 0; f < e.num_features; ++f) {
      float x = row[f];
      if (e.reward_type == TPC_REWARD_REDUCED_NODE_USAGE) x = rhu2(x);  // remove_is_empty_dim (env.py:428-430)
      dom = fminf(dom, rhu2(__fsub_rn(x, d[f])));
    }
    after = dom < 0.f ? 1.f : 0.f;
  }
  float m = fmaxf(after, bin_mask[idx]);
  if (j == 0) m = 0.f;  // EOS is always available (env.py:419)
  out[idx] = m;
}

// one warp per episode
__global__ void step_kernel(tpc_env_config e, float* __restrict__ batch, float* __restrict__ bin_mask,
                            float* __restrict__ res_mask, float* __restrict__ mha_mask, float* __restrict__ is_empty,
                            const int32_t* __restrict__ bin_ids, int step, int B, int N, int R,
                            float* __restrict__ reward) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  const int S = N + R, F = e.num_features;
  float* st = batch + (size_t)b * S * F;
  const int bin = bin_ids[b];
  const int req = step;
  float nodes[4], reqs[4], rem[4];
  bool eos = true;
  for (int f = 0; f < F; ++f) {
    nodes[f] = st[bin * F + f];
    reqs[f] = st[req * F + f];
    eos = eos && (nodes[f] == e.eos_code);
  }
  float is_full;
  if (e.kind == TPC_ENV_KNAPSACK_V2) {
    rem[0] = rhu2(nodes[0]);
    rem[1] = rhu2(__fadd_rn(nodes[1], reqs[0]));
    is_full = rhu2(__fsub_rn(rem[0], rem[1])) == 0.f ? 1.f : 0.f;
  } else {
    float dom = INFINITY;
    for (int f = 0; f < F; ++f) {
      rem[f] = rhu2(__fsub_rn(nodes[f], reqs[f]));
      dom = fminf(dom, rem[f]);
    }
    is_full = dom == 0.f ? 1.f : 0.f;
  }
  __syncwarp();
  if (lane == 0) {
    for (int f = 0; f < F; ++f) st[bin * F + f] = rem[f];
    for (int f = 0; f < F; ++f) st[f] = e.eos_code;          // keep EOS node intact (env.py:139)
    if (res_mask) res_mask[(size_t)b * S + req] = 1.f;
    bin_mask[(size_t)b * S + bin] = is_full;
    bin_mask[(size_t)b * S] = 0.f;
    mha_mask[(size_t)b * S + req] = 1.f;
    if (e.mask_nodes_in_mha) mha_mask[(size_t)b * S + bin] = is_full;
    mha_mask[(size_t)b * S] = 0.f;
  }
  __syncwarp();
  const float eosf = eos ? 1.f : 0.f;
  float rw = 0.f;
  if (e.kind == TPC_ENV_KNAPSACK_V2) {
    rw = __fmul_rn(reqs[1], __fsub_rn(1.f, eosf));             // value * (1 - eos) (knapsack_v2/reward.py:36-40)
  } else if (e.reward_type == TPC_REWARD_GREEDY) {
    rw = eos ? 0.f : 1.f;                                       // 1 - is_eos (reward.py:42-44)
  } else if (e.reward_type == TPC_REWARD_SINGLE_NODE_DOMINANT) {
    float dom = INFINITY;
    for (int f = 0; f < F; ++f) dom = fminf(dom, __fsub_rn(nodes[f], reqs[f]));  // unrounded (reward.py:73)
    rw = __fadd_rn(__fmul_rn(dom, __fsub_rn(1.f, eosf)), __fmul_rn(eosf, e.rejection_penalty));
  } else if (e.reward_type == TPC_REWARD_GLOBAL_DOMINANT) {
    float dom = INFINITY;                                       // min over real nodes of the UPDATED state
    for (int i = F + lane; i < N * F; i += 32) dom = fminf(dom, st[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dom = fminf(dom, __shfl_xor_sync(0xffffffffu, dom, o));
    rw = __fadd_rn(__fmul_rn(dom, __fsub_rn(1.f, eosf)), __fmul_rn(eosf, e.rejection_penalty));
  } else {  // reduced_node_usage (reward.py:148-164)
    float used = 0.f;
    if (lane == 0) {
      used = is_empty[(size_t)b * S + bin];
      is_empty[(size_t)b * S + bin] = 1.f;
      is_empty[(size_t)b * S] = e.eos_code;
    }
    used = __shfl_sync(0xffffffffu, used, 0);
    rw = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, used), e.use_new_node_penalty), __fsub_rn(1.f, eosf)),
                   __fmul_rn(eosf, e.rejection_penalty));
  }
  if (lane == 0) reward[b] = rw;
}

__global__ void sample_kernel(const float* __restrict__ probs, int B, int S, uint64_t seed, int64_t id0, int64_t draw,
                              int32_t* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  Stream s = make_stream(seed, id0 + b, draw, 3);
  const float u = (float)(s.draw(0) >> 8) * (1.0f / 16777216.0f);  // [0,1) with 24 random bits
  const float* p = probs + (size_t)b * S;
  float cum = 0.f;
  int pick = -1, last = 0;
  for (int j = 0; j < S; ++j) {
    const float pj = p[j];
    if (pj > 0.f) last = j;
    cum = __fadd_rn(cum, pj);
    if (pick < 0 && u < cum && pj > 0.f) pick = j;
  }
  out[b] = pick >= 0 ? pick : last;
}

}  // namespace

int env_generate_profiles(const tpc_env_config& e, uint64_t seed, float* profiles, cudaStream_t st) {
  if (e.num_profiles <= 0) return TPC_OK;
  gen_profiles_kernel<<<ceil_div(e.num_profiles, 128), 128, 0, st>>>(e, seed, profiles);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int env_reset(const tpc_env_config& e, uint64_t seed, int64_t id0, int64_t epoch, const float* profiles, int B, int N,
              int R, float* batch, float* bin_mask, float* mha_mask, float* is_empty, cudaStream_t st) {
  if (R > MAX_RULES || (e.num_profiles > 0 && R > e.num_profiles)) return TPC_ERR_SHAPE;
  if (e.num_features < 1 || e.num_features > 4) return TPC_ERR_SHAPE;
  reset_kernel<<<ceil_div(B, 64), 64, 0, st>>>(e, seed, id0, epoch, profiles, B, N, R, batch, bin_mask, mha_mask,
                                                is_empty);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int env_feasible_mask(const tpc_env_config& e, const float* state, int ld, const float* dec, const float* bin_mask,
                      int B, int S, float* out, cudaStream_t st) {
  feasible_mask_kernel<<<ceil_div(B * S, 256), 256, 0, st>>>(e, state, ld, dec, bin_mask, B, S, out);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int env_step(const tpc_env_config& e, float* batch, float* bin_mask, float* res_mask, float* mha_mask, float* is_empty,
             const int32_t* bin_ids, int step, int B, int N, int R, float* reward, cudaStream_t st) {
  if (e.reward_type == TPC_REWARD_REDUCED_NODE_USAGE && e.kind == TPC_ENV_RESOURCE_V3 && !is_empty) return TPC_ERR_ARG;
  step_kernel<<<ceil_div(B * 32, 256), 256, 0, st>>>(e, batch, bin_mask, res_mask, mha_mask, is_empty, bin_ids, step, B,
                                                     N, R, reward);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int sample_categorical(const float* probs, int B, int S, uint64_t seed, int64_t id0, int64_t draw, int32_t* out,
                       cudaStream_t st) {
  sample_kernel<<<ceil_div(B, 128), 128, 0, st>>>(probs, B, S, seed, id0, draw, out);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

}  // namespace tpc

using namespace tpc;

extern "C" {

int tpc_env_generate_profiles(tpc_handle* h, const tpc_env_config* e, uint64_t seed, float* profiles, void* stream) {
  if (!h || !e || !profiles) return TPC_ERR_ARG;
  return env_generate_profiles(*e, seed, profiles, (cudaStream_t)stream);
}

int tpc_env_reset(tpc_handle* h, const tpc_env_config* e, uint64_t seed, int64_t episode_id0, int64_t epoch,
                  const float* profiles, int B, int num_bins, int num_resources, float* batch, float* bin_net_mask,
                  float* mha_mask, float* is_empty, void* stream) {
  if (!h || !e || !batch || !bin_net_mask || !mha_mask) return TPC_ERR_ARG;
  if (e->num_profiles > 0 && !profiles) return TPC_ERR_ARG;
  return env_reset(*e, seed, episode_id0, epoch, profiles, B, num_bins, num_resources, batch, bin_net_mask, mha_mask,
                   is_empty, (cudaStream_t)stream);
}

int tpc_env_feasible_mask(tpc_handle* h, const tpc_env_config* e, const float* state, int ld_state,
                          const float* dec_input, const float* bin_net_mask, int B, int S, float* out_mask,
                          void* stream) {
  if (!h || !e || !state || !dec_input || !bin_net_mask || !out_mask) return TPC_ERR_ARG;
  return env_feasible_mask(*e, state, ld_state, dec_input, bin_net_mask, B, S, out_mask, (cudaStream_t)stream);
}

int tpc_env_step(tpc_handle* h, const tpc_env_config* e, float* batch, float* bin_net_mask, float* resource_net_mask,
                 float* mha_mask, float* is_empty, const int32_t* bin_ids, int decoding_step, int B, int num_bins,
                 int num_resources, float* reward, void* stream) {
  if (!h || !e || !batch || !bin_net_mask || !mha_mask || !bin_ids || !reward) return TPC_ERR_ARG;
  return env_step(*e, batch, bin_net_mask, resource_net_mask, mha_mask, is_empty, bin_ids, decoding_step, B, num_bins,
                  num_resources, reward, (cudaStream_t)stream);
}

int tpc_sample_categorical(tpc_handle* h, const float* probs, int B, int S, uint64_t seed, int64_t episode_id0,
                           int64_t draw_index, int32_t* out, void* stream) {
  if (!h || !probs || !out) return TPC_ERR_ARG;
  return sample_categorical(probs, B, S, seed, episode_id0, draw_index, out, (cudaStream_t)stream);
}

}  // extern "C"
