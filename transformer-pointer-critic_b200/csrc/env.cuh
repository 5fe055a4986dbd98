// Internal launchers of the environment kernels (env.cu).
#pragma once
#include "handle.cuh"

namespace tpc {

int env_generate_profiles(const tpc_env_config& e, uint64_t seed, float* profiles, cudaStream_t st);
int env_reset(const tpc_env_config& e, uint64_t seed, int64_t id0, int64_t epoch, const float* profiles, int B, int N,
              int R, float* batch, float* bin_mask, float* mha_mask, float* is_empty, cudaStream_t st);
int env_feasible_mask(const tpc_env_config& e, const float* state, int ld, const float* dec, const float* bin_mask,
                      int B, int S, float* out, cudaStream_t st);
int env_step(const tpc_env_config& e, float* batch, float* bin_mask, float* res_mask, float* mha_mask, float* is_empty,
             const int32_t* bin_ids, int step, int B, int N, int R, float* reward, cudaStream_t st);
int sample_categorical(const float* probs, int B, int S, uint64_t seed, int64_t id0, int64_t draw, int32_t* out,
                       cudaStream_t st);

}  // namespace tpc
