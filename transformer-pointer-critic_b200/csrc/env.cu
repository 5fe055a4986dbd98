// On-device ResourceV3 / KnapsackV2 environment: instance generation, feasibility mask, step + rewards,
// categorical sampling. Bit-exact restatement of the reference's fp32 NumPy arithmetic
// (environment/custom/resource_v3/env.py:115-203,387-430; reward.py:23-166; misc/utils.py:19-42;
//  environment/custom/knapsack_v2/env.py:102-185,384-408; knapsack_v2/misc/utils.py:5-32).
// Every float op is an explicit round-to-nearest intrinsic: no FMA contraction, same rounding sequence as NumPy.
#include "env.cuh"
#include "env_dev.cuh"

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
  out[idx] = feasible_elem(e, state + (size_t)idx * ld, dec + (size_t)b * e.num_features, bin_mask[idx], j);
}

// one warp per episode
__global__ void step_kernel(tpc_env_config e, float* __restrict__ batch, float* __restrict__ bin_mask,
                            float* __restrict__ res_mask, float* __restrict__ mha_mask, float* __restrict__ is_empty,
                            const int32_t* __restrict__ bin_ids, int step, int B, int N, int R,
                            float* __restrict__ reward) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  const float rw = env_step_warp(e, batch, bin_mask, res_mask, mha_mask, is_empty, b, bin_ids[b], step, N, R, lane);
  if (lane == 0) reward[b] = rw;
}

__global__ void sample_kernel(const float* __restrict__ probs, int B, int S, uint64_t seed, int64_t id0, int64_t draw,
                              int32_t* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  out[b] = sample_pick(probs + (size_t)b * S, S, seed, id0 + b, draw);
}

__global__ void cvrp_reset_kernel(uint64_t seed, int64_t id0, int64_t epoch, int B, int Nn, int V, int max_demand,
                                  int capacity, float* __restrict__ batch, float* __restrict__ backpack,
                                  float* __restrict__ item, float* __restrict__ mha) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int S = Nn + V;
  Stream sr = make_stream(seed, id0 + b, epoch, 5);
  float* st = batch + (size_t)b * S * 3;
  for (int i = 0; i < Nn; ++i) {
    st[i * 3 + 0] = (float)sr.randint(i * 3, 0, 100);
    st[i * 3 + 1] = (float)sr.randint(i * 3 + 1, 0, 100);
    st[i * 3 + 2] = i == 0 ? 0.f : (float)sr.randint(i * 3 + 2, 1, max_demand + 1);
  }
  for (int i = Nn; i < S; ++i) {
    st[i * 3 + 0] = st[0];
    st[i * 3 + 1] = st[1];
    st[i * 3 + 2] = (float)capacity;
  }
  for (int j = 0; j < S; ++j) {
    backpack[(size_t)b * S + j] = j < Nn ? 1.f : 0.f;
    if (item) item[(size_t)b * S + j] = (j >= Nn || j == 0) ? 1.f : 0.f;
    mha[(size_t)b * S + j] = 0.f;
  }
}

__global__ void cvrp_feasible_kernel(const float* __restrict__ state, const float* __restrict__ items,
                                     const float* __restrict__ backpack, int B, int S, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * S) return;
  const int b = idx / S;
  out[idx] = cvrp_feasible_elem(state + (size_t)idx * 3, items[(size_t)b * 3 + 2], backpack[idx]);
}

__global__ void cvrp_step_kernel(float* __restrict__ batch, float* __restrict__ backpack, float* __restrict__ item,
                                 float* __restrict__ mha, const int32_t* __restrict__ vids,
                                 const int32_t* __restrict__ nids, int B, int Nn, int V, int visited,
                                 float* __restrict__ reward) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  const int S = Nn + V;
  const float rw = cvrp_step_warp(batch, backpack, item, mha, b, vids[b], nids[b], S, lane);
  if (lane == 0) {
    reward[b] = rw;
    if (item && visited + 1 == Nn - 1) item[(size_t)b * S] = 0.f;   // every node visited: the depot opens (:132-134)
  }
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
  if (e.kind == TPC_ENV_CVRP) {
    if (ld != 3) return TPC_ERR_SHAPE;
    cvrp_feasible_kernel<<<ceil_div(B * S, 256), 256, 0, st>>>(state, dec, bin_mask, B, S, out);
  } else {
    feasible_mask_kernel<<<ceil_div(B * S, 256), 256, 0, st>>>(e, state, ld, dec, bin_mask, B, S, out);
  }
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

int tpc_cvrp_reset(tpc_handle* h, uint64_t seed, int64_t episode_id0, int64_t epoch, int B, int num_nodes,
                   int num_vehicles, int max_demand, int capacity, float* batch, float* backpack_net_mask,
                   float* item_net_mask, float* mha_mask, void* stream) {
  if (!h || !batch || !backpack_net_mask || !mha_mask) return TPC_ERR_ARG;
  if (B < 0 || num_nodes < 2 || num_vehicles < 1 || max_demand < 1 || capacity < max_demand) return TPC_ERR_SHAPE;
  if (B == 0) return TPC_OK;
  cvrp_reset_kernel<<<ceil_div(B, 64), 64, 0, (cudaStream_t)stream>>>(seed, episode_id0, epoch, B, num_nodes, num_vehicles,
                                                                    max_demand, capacity, batch, backpack_net_mask,
                                                                    item_net_mask, mha_mask);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int tpc_cvrp_feasible_mask(tpc_handle* h, const float* state, const float* items, const float* backpack_net_mask, int B,
                           int S, float* out_mask, void* stream) {
  if (!h || !state || !items || !backpack_net_mask || !out_mask) return TPC_ERR_ARG;
  if (B < 0 || S < 1) return TPC_ERR_SHAPE;
  if (B == 0) return TPC_OK;
  cvrp_feasible_kernel<<<ceil_div(B * S, 256), 256, 0, (cudaStream_t)stream>>>(state, items, backpack_net_mask, B, S,
                                                                             out_mask);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int tpc_cvrp_step(tpc_handle* h, float* batch, float* backpack_net_mask, float* item_net_mask, float* mha_mask,
                  const int32_t* vehicle_ids, const int32_t* node_ids, int B, int num_nodes, int num_vehicles,
                  int visited_nodes, float* reward, void* stream) {
  if (!h || !batch || !backpack_net_mask || !mha_mask || !vehicle_ids || !node_ids || !reward) return TPC_ERR_ARG;
  if (B < 0 || num_nodes < 2 || num_vehicles < 1) return TPC_ERR_SHAPE;
  if (B == 0) return TPC_OK;
  cvrp_step_kernel<<<ceil_div(B * 32, 256), 256, 0, (cudaStream_t)stream>>>(batch, backpack_net_mask, item_net_mask, mha_mask,
                                                                          vehicle_ids, node_ids, B, num_nodes,
                                                                          num_vehicles, visited_nodes, reward);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

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
