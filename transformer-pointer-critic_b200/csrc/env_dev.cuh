// Device functions of the environment step, shared by the stand-alone kernels (env.cu) and the fused decision
// kernel of the rollout (decide.cu): one definition, so both paths execute the same fp32 operation sequence.
// Reference: environment/custom/resource_v3/env.py:115-203,387-430; reward.py:23-166; misc/utils.py:19-42;
// environment/custom/knapsack_v2/env.py:102-185,384-408; agents/agent.py:275-280 (sampling).
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace tpc {

// build_feasible_mask for position j of one episode: row = state[j][0:F], d = decoder input (the rule to place)
__device__ __forceinline__ float feasible_elem(const tpc_env_config& e, const float* __restrict__ row,
                                               const float* __restrict__ d, float bin_mask_j, int j) {
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
  float m = fmaxf(after, bin_mask_j);
  if (j == 0) m = 0.f;  // EOS is always available (env.py:419)
  return m;
}

// env.step of episode b by ONE WARP (all 32 lanes call it): places rule `req` (= decoding step) on node `bin`,
// updates state and masks in place, returns the reward (valid on every lane).
__device__ __forceinline__ float env_step_warp(const tpc_env_config& e, float* __restrict__ batch,
                                               float* __restrict__ bin_mask, float* __restrict__ res_mask,
                                               float* __restrict__ mha_mask, float* __restrict__ is_empty, int b, int bin,
                                               int req, int N, int R, int lane) {
  const int S = N + R, F = e.num_features;
  bin = min(max(bin, 0), S - 1);   // memory safety only: the host wrappers reject out-of-range ids (IndexError in the reference)
  float* st = batch + (size_t)b * S * F;
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
  return rw;
}

// ---- CVRP (environment/custom/vrp/env.py) ------------------------------------------------------------------------
// build_feasible_mask for position j (:270-301): w = demand of the served node
__device__ __forceinline__ float cvrp_feasible_elem(const float* __restrict__ row, float w, float backpack_mask_j) {
  const float tot = __fmul_rn(__fsub_rn(row[2], w), __fsub_rn(1.f, backpack_mask_j));
  return fmaxf(tot < 0.f ? 1.f : 0.f, backpack_mask_j);
}
// step(vehicle, node) of episode b by one warp (:67-146): returns the reward -round(distance) on every lane.
// item_mask may be nullptr (the single-pointer rollout serves the nodes in sequence order and never reads it).
__device__ __forceinline__ float cvrp_step_warp(float* __restrict__ batch, float* __restrict__ backpack_mask,
                                                float* __restrict__ item_mask, float* __restrict__ mha_mask, int b,
                                                int veh, int node, int S, int lane) {
  veh = min(max(veh, 0), S - 1);
  node = min(max(node, 0), S - 1);
  float* st = batch + (size_t)b * S * 3;
  const float nx = st[node * 3], ny = st[node * 3 + 1], nd = st[node * 3 + 2];
  const float vx = st[veh * 3], vy = st[veh * 3 + 1], vc = st[veh * 3 + 2];
  const float xd = __fsub_rn(vx, nx), yd = __fsub_rn(vy, ny);
  // round(sqrt(xd*xd + yd*yd)): float32 products and sum, float64 sqrt, Python round (half to even)
  const float d2 = __fadd_rn(__fmul_rn(xd, xd), __fmul_rn(yd, yd));
  const float rw = -(float)rint(sqrt((double)d2));
  __syncwarp();
  if (lane == 0) {
    st[veh * 3] = nx;
    st[veh * 3 + 1] = ny;
    st[veh * 3 + 2] = __fsub_rn(vc, nd);
    if (nd != 0.f) {
      if (item_mask) item_mask[(size_t)b * S + node] = 1.f;
      mha_mask[(size_t)b * S + node] = 1.f;
    } else {
      backpack_mask[(size_t)b * S + veh] = 1.f;
      mha_mask[(size_t)b * S + veh] = 1.f;
    }
  }
  __syncwarp();
  return rw;
}
// adapter: token index of the node served at step t
__host__ __device__ inline int cvrp_request(int t, int num_nodes) { return t < num_nodes - 1 ? t + 1 : 0; }

// tfp Categorical(probs).sample() substitute: inverse-CDF draw on the Philox uniform (seed, episode id, draw index);
// a sequential fp32 running sum over p[0..S) (p may live in shared or global memory)
__device__ __forceinline__ int sample_pick(const float* p, int S, uint64_t seed, int64_t id, int64_t draw) {
  Stream s = make_stream(seed, id, draw, 3);
  const float u = (float)(s.draw(0) >> 8) * (1.0f / 16777216.0f);  // [0,1) with 24 random bits
  float cum = 0.f;
  int pick = -1, last = 0;
  for (int j = 0; j < S; ++j) {
    const float pj = p[j];
    if (pj > 0.f) last = j;
    cum = __fadd_rn(cum, pj);
    if (pick < 0 && u < cum && pj > 0.f) pick = j;
  }
  return pick >= 0 ? pick : last;
}

}  // namespace tpc
