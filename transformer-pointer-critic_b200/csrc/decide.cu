// Fused decision kernel of the rollout: everything of one decoding step that follows the encoder / decoder
// projections, one CTA per episode --
//   pointer head: V . tanh(W1 dec + W2 enc), tanh-clipped logits, feasibility mask, softmax   (pointer_attention.py:40-66)
//   pick: greedy argmax or Philox inverse-CDF categorical draw                                (agents/agent.py:275-280)
//   environment step on device: resource decrement, masks, reward                              (env.py:115-203, reward.py)
//   Agent.store of the NEXT step's inputs: state / mha-mask / is_empty copies, next decoder
//   input (the next rule) and its feasible mask                                                (agents/trainer.py:49-76, env.py:387-422)
// It replaces six launches and three device-to-device copies per step (pointer_fwd, sample, step, scatter_reward,
// gather_dec, feasible_mask). The arithmetic is the device functions of pointer_dev.cuh / env_dev.cuh, i.e. the
// stand-alone kernels' own code: the step-wise API and the fused rollout produce identical bits.
#include "decide.cuh"
#include "env_dev.cuh"
#include "pointer_dev.cuh"

namespace tpc {

namespace {

__global__ void __launch_bounds__(256) decide_step_kernel(const DecideArgs a) {
  __shared__ float sl[PTR_MAX_S];
  __shared__ float red[8];
  __shared__ int redi[8];
  __shared__ int pick_s;
  const int c = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = a.N + a.R, F = a.env.num_features;

  const int Sc = S - a.win.gap;   // compact tokens (the window leaves out rules placed at earlier steps)
  const int best = pointer_head(c, a.w1d, a.kvw, a.V, a.bV, a.clipC, a.has_clip, a.ptr_mask, Sc, a.score, a.logits,
                                nullptr, sl, red, redi, a.win);
  if (threadIdx.x == 0) {
    // the draw runs over the compact probabilities: the tokens left out have probability exactly 0, so the inverse-CDF
    // pick over the full row is the same token
    const int pick = a.greedy ? best : a.win.full(sample_pick(sl, Sc, a.seed, a.id0 + c, a.draw + (a.draw0 ? *a.draw0 : 0)));
    a.actions[c] = pick;
    pick_s = pick;
  }
  __syncthreads();
  if (warp == 0) {
    const float rw = a.env.kind == TPC_ENV_CVRP
                         ? cvrp_step_warp(a.batch, a.bin_mask, nullptr, a.mha_mask, c, pick_s, a.req, S, lane)
                         : env_step_warp(a.env, a.batch, a.bin_mask, nullptr, a.mha_mask, a.is_empty, c, pick_s, a.req,
                                         a.N, a.R, lane);
    if (lane == 0) a.rewards[(size_t)c * a.T + a.t] = rw;
  }
  if (a.t + 1 >= a.T) return;
  __syncthreads();   // the step's global writes are visible to the whole CTA

  // inputs of step t + 1, stored the way Agent.store keeps them
  const float* st = a.batch + (size_t)c * S * F;
  const float* dec = st + (size_t)a.req_next * F;
  if (threadIdx.x < F) a.dec_next[(size_t)c * F + threadIdx.x] = dec[threadIdx.x];
  if (a.states_next)
    for (int i = threadIdx.x; i < S * F; i += 256) a.states_next[(size_t)c * S * F + i] = st[i];
  for (int j = threadIdx.x; j < S; j += 256) {
    const size_t o = (size_t)c * S + j;
    if (a.mha_next) a.mha_next[o] = a.mha_mask[o];
    if (a.is_empties_next) a.is_empties_next[o] = a.is_empty[o];
    a.pmask_next[o] = a.env.kind == TPC_ENV_CVRP ? cvrp_feasible_elem(st + (size_t)j * F, dec[2], a.bin_mask[o])
                                                 : feasible_elem(a.env, st + (size_t)j * F, dec, a.bin_mask[o], j);
  }
}

}  // namespace

int decide_step(const DecideArgs& a, int C, cudaStream_t st) {
  if (C <= 0) return TPC_OK;
  if (a.N + a.R > PTR_MAX_S) return TPC_ERR_SHAPE;
  decide_step_kernel<<<C, 256, 0, st>>>(a);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

}  // namespace tpc
