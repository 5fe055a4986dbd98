// Fused per-step decision kernel of the rollout (decide.cu).
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace tpc {

struct DecideArgs {
  // pointer head inputs (see pointer_fwd in kernels.cuh)
  const float *w1d, *kvw, *V, *bV;
  float clipC;
  int has_clip;
  const float* ptr_mask;   // feasible mask of THIS step (C, S)
  float *score, *logits;   // optional outputs
  TokWin win;              // compact token window of kvw (rule tokens placed at earlier steps are left out)
  // pick
  int greedy;
  uint64_t seed;
  int64_t id0, draw;
  const int64_t* draw0;    // optional device scalar added to `draw` (graph replay of the rollout: epoch * T lives here)
  int32_t* actions;        // (C) slot of this step
  // environment (live state, in place)
  tpc_env_config env;
  float *batch, *bin_mask, *mha_mask, *is_empty;
  int N, R, t;             // first-segment tokens (nodes incl. EOS), second-segment tokens, this step
  int T;                   // steps of an episode (ResourceV3 / KnapsackV2: R; CVRP adapter: N - 1 + R)
  int req, req_next;       // token index of the request served at this step / at the next one (the decoder input)
  float* rewards;          // (C, T)
  // inputs of step t + 1 (written unless t + 1 == R); the three *_next trajectory slots may be NULL
  float *dec_next, *pmask_next, *states_next, *mha_next, *is_empties_next;
};

int decide_step(const DecideArgs& a, int C, cudaStream_t st);

}  // namespace tpc
