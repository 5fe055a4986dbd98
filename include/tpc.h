/* libtpc.so -- C-ABI of the B200-native rollout + A2C-update hot path of transformer-pointer-critic.
 *
 * The reference (AndreMaz/transformer-pointer-critic) is a pure-Python/TensorFlow program with no FFI; its
 * seam is the Python object protocol between agents/trainer.py, environment/custom/resource_v3/tester.py and
 * the Agent / env objects. Each entry point below names the reference interface it replaces
 * (paths relative to the reference repository). INTEGRATION.md shows the ctypes stub a maintainer adds.
 *
 * Conventions: every pointer is a DEVICE pointer owned by the caller (torch-allocated) unless marked HOST;
 * every call is asynchronous on `stream` (a cudaStream_t passed as void*), returns 0 or a negative TPC_ERR_*
 * code, never throws; no global mutable state; a handle / model is not thread-safe, different ones are.
 * All floating-point tensors are fp32, row-major; masks are fp32 with 1 = masked (as in the reference).
 */
#ifndef TPC_H_
#define TPC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPC_OK 0
#define TPC_ERR_CUDA -1
#define TPC_ERR_ARG -2
#define TPC_ERR_SHAPE -3
#define TPC_ERR_WORKSPACE -4
#define TPC_ERR_DRIVER -5

typedef struct tpc_handle tpc_handle;
typedef struct tpc_model tpc_model;

/* ------------------------------------------------------------------------------------------------------
 * configuration (configs/ResourceV3.json:90-164, configs/KnapsackV2.json; agents/models/model_factory.py:4-36;
 * environment/custom/resource_v3/env.py:21-81,314-334)
 * ---------------------------------------------------------------------------------------------------- */
typedef struct tpc_model_config {
  int actor_num_layers, actor_inner_layer_dim;     /* actor.num_layers, actor.inner_layer_dim */
  int critic_num_layers, critic_inner_layer_dim;   /* critic.num_layers, critic.inner_layer_dim */
  int dim_model, num_heads;                        /* must be 128 / 8 (all shipped configs) */
  int attention_dense_units, last_layer_units;     /* must be 128 */
  float logit_clipping_C;                          /* actor.logit_clipping_C */
  int has_logit_clipping;                          /* 0 when logit_clipping_C is null */
  int common_embedding;                            /* encoder_embedding.common (env.py:323-332) */
  int num_bin_features, num_resource_features;     /* 3/3 ResourceV3, 4/3 "cost" reward, 2/2 KnapsackV2 */
  int dec_features;                                /* decoder-input width = env num_features */
  int num_bins, num_resources;                     /* TRAINING shape: critic Flatten width (critic/model.py:65) */
} tpc_model_config;

enum { TPC_ENV_RESOURCE_V3 = 0, TPC_ENV_KNAPSACK_V2 = 1, TPC_ENV_CVRP = 2 };
enum { TPC_REWARD_GREEDY = 0, TPC_REWARD_SINGLE_NODE_DOMINANT = 1, TPC_REWARD_GLOBAL_DOMINANT = 2,
       TPC_REWARD_REDUCED_NODE_USAGE = 3 };

typedef struct tpc_env_config {
  int kind;                    /* TPC_ENV_* */
  int reward_type;             /* TPC_REWARD_* (environment/custom/resource_v3/reward.py:6-21) */
  float rejection_penalty;     /* reward.<type>.rejection_penalty */
  float use_new_node_penalty;  /* reward.reduced_node_usage.use_new_node_penalty */
  int mask_nodes_in_mha;       /* env_config.mask_nodes_in_mha (env.py:153-156) */
  int num_features;            /* 3 (ResourceV3) or 2 (KnapsackV2) */
  float eos_code;              /* EOS_CODE (-2) */
  /* instance generation (env.py:205-263): integers / normalization_factor */
  int normalization_factor;
  int node_min_val, node_max_val;  /* nodes: U{min..max-1} ; KnapsackV2: bin capacity */
  int req_min_val, req_max_val;    /* rules: U{min..max-1} ; KnapsackV2: item weight */
  int val_min_val, val_max_val;    /* KnapsackV2 only: item value */
  int num_profiles;                /* profile table rows (0 => generate_request_on_the_fly) */
} tpc_env_config;

/* ------------------------------------------------------------------------------------------------------
 * library + model objects
 * ---------------------------------------------------------------------------------------------------- */
int tpc_create(tpc_handle** out);
int tpc_destroy(tpc_handle* h);
int tpc_num_sms(const tpc_handle* h);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long tpc_launch_count(void);

/* replaces model_factory('transformer', opts) (agents/models/model_factory.py:4-36) */
int tpc_model_create(tpc_handle* h, const tpc_model_config* cfg, tpc_model** out);
int tpc_model_destroy(tpc_model* m);
/* which: 0 actor, 1 critic. Canonical flat parameter vector (Keras trainable_weights order). */
size_t tpc_param_count(const tpc_model* m, int which);
int tpc_num_tensors(const tpc_model* m, int which);
/* HOST out arrays of length tpc_num_tensors: offset and size of each canonical tensor */
int tpc_param_layout(const tpc_model* m, int which, int64_t* offsets, int64_t* sizes);
/* Name and shape of canonical tensor `index` (Keras `trainable_weights` order, e.g. "encoder.enc_layers.0.mha.wq.kernel"):
 * kernels are [rows][cols] = [in][out], 1-D tensors report cols = 0. Used by the weight interchange
 * (agents/agent.py:289-294: Keras attribute paths of a TF checkpoint). */
int tpc_param_info(const tpc_model* m, int which, int index, char* name, int name_capacity, int64_t* rows, int64_t* cols);
/* floats in the kernel-side weight image: transposed / fused copies of the forward weights with their fp16 split images
 * (see tpc_gemm), TF32-rounded copies for the backward GEMMs */
size_t tpc_shadow_count(const tpc_model* m, int which);
/* rebuild the weight image after params changed (after load_weights / every optimizer step) */
int tpc_prepare_weights(tpc_model* m, int which, const float* params, float* shadow, void* stream);

/* bytes of scratch a call needs for `nstates` states of shape (num_bins + num_resources) tokens.
 * mode 0: actor forward (rollout step)   1: critic forward   2: actor forward+backward   3: critic fwd+bwd
 *      4: tpc_rollout of nstates episodes */
size_t tpc_workspace_bytes(const tpc_model* m, int mode, int nstates, int num_bins, int num_resources);

/* workspace of tpc_rollout = tpc_workspace_bytes(m, 4, B, ...); of tpc_a2c_grads: */
size_t tpc_a2c_workspace_bytes(const tpc_model* m, int B, int num_bins, int num_resources, int chunk_states);

/* ------------------------------------------------------------------------------------------------------
 * stand-alone operators (unit-test granularity)
 * ---------------------------------------------------------------------------------------------------- */
/* D[M,N] = epi(A[M,K] @ Bt[N,K]^T + bias); flags bit0 relu, bit1 layernorm(gamma,beta -> rstd_out),
 * bit2 round outputs to TF32, bit3 atomic split-K accumulate; aux_mode 0 none / 1 add / 2 relu-gate.
 * Bt_corr == NULL: one TF32 pass (operands truncated by the tensor core). Bt_corr = split image of Bt made by
 * tpc_gemm_split_image: fp32-accurate product D = A_h.B_h + A_h.B_r + A_r.B_h with x_h = fp16(x), x_r = fp16(x - x_h)
 * and fp32 accumulation, ~2^-21 relative (the forward passes use this; the backward passes use one pass on
 * TF32-rounded gradients). Operands beyond +-65504 saturate. */
int tpc_gemm(tpc_handle* h, const float* A, int lda, const float* Bt, int ldb, float* D, int ldd, int M, int N, int K,
             const float* bias, const float* aux, int ldaux, int aux_mode, const float* gamma, const float* beta,
             float* rstd_out, int flags, int k_splits, const void* Bt_corr, void* stream);
/* image[2][N][ldb] fp16 = { B_h = fp16(Bt), B_r = fp16(Bt - B_h) }: N * ldb * 4 bytes, 16-byte aligned, ldb % 8 == 0 */
int tpc_gemm_split_image(tpc_handle* h, const float* Bt, int ldb, int N, int K, void* image, void* stream);
/* Backward GEMM with the LayerNorm backward fused into its epilogue (N = 128, one TF32 pass): with
 * dy = A[M,K] @ Bt[128,K]^T (+ aux[M,128] when aux != NULL) the gradient w.r.t. a LayerNorm output y = LN(pre),
 *   D[M,128] = dpre = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma,  xhat = (y - beta) / gamma,
 *   dgamma[128] += sum_rows(dy * xhat),  dbeta[128] += sum_rows(dy)
 * (tf.keras LayerNormalization under tf.GradientTape, agents/trainer.py:115-141; D is RN-rounded to TF32 because it
 * feeds the next backward GEMMs). dy itself is never written. */
int tpc_gemm_ln_bwd(tpc_handle* h, const float* A, int lda, const float* Bt, int ldb, float* D, int ldd, int M, int K,
                    const float* aux, int ldaux, const float* y, const float* rstd, const float* gamma,
                    const float* beta, float* dgamma, float* dbeta, void* stream);
/* dW[Kin,Nout] += X[M,Kin]^T @ dY[M,Nout];  db[Nout] += colsum(dY) when db != NULL */
int tpc_wgrad(tpc_handle* h, const float* X, int ldx, const float* dY, int ldy, float* dW, int ldw, float* db, int M,
              int Kin, int Nout, void* stream);

/* self-attention of one encoder stack on packed qkv[C*L][384] (q | k | v), 8 heads x 16; mask[c*mS + m0 + j] != 0
 * hides key j. Replaces scaled_dot_product_attention + head split/merge (common/attention.py:37-48,
 * common/utils.py:13-47) and its backward. att[C*L][128], lse[C*L][8] (log-sum-exp, saved for the backward). */
int tpc_attention_fwd(tpc_handle* h, const float* qkv, const float* mask, int mS, int m0, int C, int L, float* att,
                      float* lse, void* stream);
int tpc_attention_bwd(tpc_handle* h, const float* qkv, const float* att, const float* lse, const float* datt,
                      const float* mask, int mS, int m0, int C, int L, float* dqkv, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * model forward passes (step-wise, API-exact mode used by Agent.act / tests)
 * ---------------------------------------------------------------------------------------------------- */
/* replaces ActorTransformer.call (agents/models/transformer/actor/model.py:41-72) as used by Agent.act
 * (agents/agent.py:265-273): state (B,S,Fstate) with Fstate = num_bin_features when not common_embedding
 * (4th column = is_empty) else dec_features; dec_input (B,dec_features); ptr_mask (B,S) = feasible mask;
 * mha_mask (B,S). Outputs: logits (B,S), probs (B,S), argmax (B) int32 (argmax over probs, lowest index wins). */
int tpc_actor_forward(tpc_model* m, const float* shadow, const float* params, const float* state,
                      const float* dec_input, const float* ptr_mask, const float* mha_mask, int B, int num_bins,
                      int num_resources, float* logits, float* probs, int32_t* argmax, void* workspace,
                      size_t workspace_bytes, void* stream);
/* replaces CriticTransformer.call (agents/models/transformer/critic/model.py:55-85): values (nstates) */
int tpc_critic_forward(tpc_model* m, const float* shadow, const float* params, const float* state,
                       const float* mha_mask, int nstates, int num_bins, int num_resources, float* values,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * environment (environment/custom/resource_v3/env.py, reward.py, misc/utils.py; knapsack_v2/env.py)
 * ---------------------------------------------------------------------------------------------------- */
/* replaces generate_dataset (env.py:205-215): profiles (num_profiles, F) = U{req_min..req_max-1}/norm.
 * Philox4x32-10, key = seed. */
int tpc_env_generate_profiles(tpc_handle* h, const tpc_env_config* e, uint64_t seed, float* profiles, void* stream);
/* replaces reset()/generate_batch()/generate_masks() (env.py:83-99,217-288): batch (B,S,F) with row 0 = EOS,
 * bin_net_mask (B,S), mha_mask (B,S), is_empty (B,S) (reduced_node_usage only, may be NULL otherwise).
 * Episode b uses the Philox stream (seed, episode_id0 + b, epoch) => independent of the GPU count. */
int tpc_env_reset(tpc_handle* h, const tpc_env_config* e, uint64_t seed, int64_t episode_id0, int64_t epoch,
                  const float* profiles, int B, int num_bins, int num_resources, float* batch, float* bin_net_mask,
                  float* mha_mask, float* is_empty, void* stream);
/* replaces build_feasible_mask (env.py:387-422): state (B,S,ld_state>=F), dec_input (B,F) */
int tpc_env_feasible_mask(tpc_handle* h, const tpc_env_config* e, const float* state, int ld_state,
                          const float* dec_input, const float* bin_net_mask, int B, int S, float* out_mask,
                          void* stream);
/* replaces step (env.py:115-203): in-place on batch / bin_net_mask / resource_net_mask (may be NULL) / mha_mask /
 * is_empty; decoding_step = index of the rule being placed; reward (B). */
int tpc_env_step(tpc_handle* h, const tpc_env_config* e, float* batch, float* bin_net_mask, float* resource_net_mask,
                 float* mha_mask, float* is_empty, const int32_t* bin_ids, int decoding_step, int B, int num_bins,
                 int num_resources, float* reward, void* stream);
/* replaces tfp.distributions.Categorical(probs).sample() (agents/agent.py:275-280): inverse-CDF draw with the
 * Philox uniform (seed, episode_id0 + b, draw_index). */
int tpc_sample_categorical(tpc_handle* h, const float* probs, int B, int S, uint64_t seed, int64_t episode_id0,
                           int64_t draw_index, int32_t* out, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * fused fast path: rollout + A2C update (agents/trainer.py:37-175 behind one call each)
 * ---------------------------------------------------------------------------------------------------- */
typedef struct tpc_rollout_buffers {
  /* live environment state (in/out): as produced by tpc_env_reset */
  float* batch; float* bin_net_mask; float* mha_mask; float* is_empty;
  /* trajectory memory = Agent.store (agents/agent.py:75-95), step-major: */
  float* states;       /* (T,B,S,F)      may be NULL in testing mode */
  float* is_empties;   /* (T,B,S)        reduced_node_usage only, else NULL */
  float* dec_inputs;   /* (T,B,F) */
  float* ptr_masks;    /* (T,B,S) feasible masks returned by act */
  float* mha_masks;    /* (T,B,S) */
  int32_t* actions;    /* (T,B) */
  float* rewards;      /* (B,T) */
  float* logits;       /* (T,B,S) optional, may be NULL */
} tpc_rollout_buffers;

/* T = num_resources steps of {feasible mask -> actor forward -> greedy / sampled pick -> env step}. */
int tpc_rollout(tpc_model* m, const tpc_env_config* e, const float* actor_shadow, const float* actor_params,
                const tpc_rollout_buffers* buf, int B, int num_bins, int num_resources, int greedy, uint64_t seed,
                int64_t episode_id0, int64_t epoch, void* workspace, size_t workspace_bytes, void* stream);
/* The loop over the T decoding steps (agents/trainer.py:49-100) is replayed as ONE CUDA graph from the second call with
 * the same arguments on (every kernel argument of the loop is a function of the call's arguments; the Philox draw index
 * epoch * T is read from device memory). Returns how many rollouts of this model ran as a graph replay.
 * TPC_ROLLOUT_GRAPH=0 in the environment keeps plain stream launches. */
unsigned long long tpc_rollout_graph_replays(const tpc_model* m);

/* ------------------------------------------------------------------------------------------------------
 * CVRP (environment/custom/vrp/env.py; BASELINE config #5). State (B, Nn + V, 3) fp32: rows [0, Nn) nodes [x, y, demand]
 * (row 0 = depot), rows [Nn, Nn + V) vehicles [x, y, remaining capacity]; raw integers, un-normalised.
 * The reference Agent cannot drive this class; tpc_cvrp_step / tpc_cvrp_feasible_mask are its
 * methods as they are, and tpc_rollout with env kind TPC_ENV_CVRP is the single-pointer adapter (oracle/cvrp.py,
 * DESIGN.md): num_bins = Nn, num_resources = V, T = Nn - 1 + V steps, step t serves node t + 1 (then the depot V
 * times) and the pointer chooses the vehicle.
 * ------------------------------------------------------------------------------------------------------ */
/* CVRP.generate_batch replacement (vrp/env.py:148-206 samples one Augerat file; here: synthetic instances of the same
 * form): depot and nodes at integer coordinates U{0..99}, demands U{1..max_demand}, V vehicles at the depot with
 * `capacity`; masks as generate_masks (:208-232). Philox stream keyed by (seed, global episode id, epoch). */
int tpc_cvrp_reset(tpc_handle* h, uint64_t seed, int64_t episode_id0, int64_t epoch, int B, int num_nodes,
                   int num_vehicles, int max_demand, int capacity, float* batch, float* backpack_net_mask,
                   float* item_net_mask, float* mha_mask, void* stream);
/* CVRP.build_feasible_mask (vrp/env.py:270-301): items (B, 3) = the node rows being served. */
int tpc_cvrp_feasible_mask(tpc_handle* h, const float* state, const float* items, const float* backpack_net_mask, int B,
                           int S, float* out_mask, void* stream);
/* CVRP.step(vehicle_ids, node_ids) (vrp/env.py:67-146) in place; visited_nodes = the env's counter BEFORE this step
 * (the depot becomes selectable in item_net_mask once it reaches Nn - 2); reward (B). item_net_mask may be NULL. */
int tpc_cvrp_step(tpc_handle* h, float* batch, float* backpack_net_mask, float* item_net_mask, float* mha_mask,
                  const int32_t* vehicle_ids, const int32_t* node_ids, int B, int num_nodes, int num_vehicles,
                  int visited_nodes, float* reward, void* stream);

typedef struct tpc_a2c_hyper {
  float gamma, values_loss_coefficient, entropy_coefficient;  /* agents/agent.py:31-33 */
  int chunk_states;       /* states per forward/backward chunk (0 = default) */
  int total_states;       /* T*B of the WHOLE job (all ranks): denominator of both mean losses */
  int phases;             /* 0 = critic then actor (trainer.py:115-141); TPC_A2C_CRITIC / TPC_A2C_ACTOR run one tape only,
                             so that the caller can start the critic-gradient all-reduce while the actor tape runs.
                             The actor phase reads the `advantages` the critic phase wrote. */
} tpc_a2c_hyper;
#define TPC_A2C_CRITIC 1
#define TPC_A2C_ACTOR 2

/* compute_discounted_rewards + compute_value_loss + compute_actor_loss + both GradientTapes
 * (agents/agent.py:110-242, agents/trainer.py:103-141). Gradients are ACCUMULATED into actor_grads /
 * critic_grads (canonical order; caller zeroes them, all-reduces them across ranks, then calls tpc_adam_apply).
 * losses[4] (device): sum over local states of {(R-V)^2 * c_v, policy_loss, total actor loss, entropy}. */
int tpc_a2c_grads(tpc_model* m, const float* actor_shadow, const float* actor_params, const float* critic_shadow,
                  const float* critic_params, const tpc_rollout_buffers* buf, int B, int num_bins, int num_resources,
                  const tpc_a2c_hyper* hp, float* actor_grads, float* critic_grads, float* values, float* advantages,
                  float* losses, void* workspace, size_t workspace_bytes, void* stream);

/* Zero `nbytes` of device memory on `stream` (gradient / loss accumulators before tpc_a2c_grads; the reference's tapes
 * start from fresh tensors, agents/trainer.py:115,125). */
int tpc_zero(tpc_handle* h, void* ptr, size_t nbytes, void* stream);

/* out[i] = in[i] * scale, n small (the 4 loss sums -> means: value_loss / mean(policy_loss) / bin_loss /
 * mean(entropy) of agents/trainer.py:159-162). in == out is allowed. */
int tpc_scale(tpc_handle* h, const float* in, int n, float scale, float* out, void* stream);

/* Roofline denominator for the projection GEMMs ("measure a tcgen05 kind::tf32 GEMM peak"):
 * every SM issues `iters` x 4 tcgen05.mma kind::tf32 (M=128, N=256, K=8) on operands resident in shared memory, timed
 * with CUDA events on `stream` (the call synchronises). *tflops_host = dense TF32 TFLOP/s at the clocks of that moment. */
int tpc_tf32_peak_probe(tpc_handle* h, int iters, float* tflops_host, void* stream);

/* Keras Adam with per-variable clipnorm (agents/agent.py:39-49): clipnorm <= 0 disables it.
 * step is 1-based. grad_scale multiplies the gradients first (1/world_size after a sum all-reduce). */
int tpc_adam_apply(tpc_model* m, int which, float* params, const float* grads, float* adam_m, float* adam_v, int step,
                   float lr, float clipnorm, float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * tester: batched heuristics and solution statistics (environment/custom/resource_v3/tester.py:117-255)
 * One warp per instance; the reference solves one instance at a time with Python Node / Resource objects and
 * asserts batch == 1 (heuristic/base_heuristic.py:39,59).
 * state: (B, >= num_bins + num_resources rows, >= 3 features) fp32, instance stride ld_instance and row stride
 * ld_row in floats; row 0 is the EOS node. assign (B, num_resources) int32: node id per rule in the ORIGINAL
 * rule order, 0 = rejected. remaining (B, num_bins, 3): node resources after the placement.
 * ---------------------------------------------------------------------------------------------------- */
/* replaces DominantResourceHeuristic.solve (heuristic/dominant_heuristic.py:30-71): rules in (stable) order of
 * max(CPU,RAM,MEM), each to the feasible node with the largest / smallest dominant remainder (node.py:36-70). */
int tpc_heuristic_dominant(tpc_handle* h, const float* state, int ld_instance, int ld_row, int B, int num_bins,
                           int num_resources, int resource_sort_descending, int node_sort_descending, int32_t* assign,
                           float* remaining, void* stream);
/* replaces RandomHeuristic.solve (heuristic/random_heuristic.py:28-65): the i-th random.randrange(n) call of
 * instance b is floor(u * n / 2^32) with u = word i of the Philox4x32-10 stream (seed, instance_id0 + b, epoch,
 * kind 4). */
int tpc_heuristic_random(tpc_handle* h, const float* state, int ld_instance, int ld_row, int B, int num_bins,
                         int num_resources, uint64_t seed, int64_t instance_id0, int64_t epoch, int32_t* assign,
                         float* remaining, void* stream);
/* replaces compute_stats (misc/utils.py:56-81) for every instance: dominant = min remaining resource over the
 * real nodes (unrounded), rejected = rules at node 0, empty_nodes = real nodes without a rule. assign is read as
 * assign[b * stride_instance + j * stride_rule] (the net's step-major actions (T,B): strides 1 and B). */
int tpc_solution_stats(tpc_handle* h, const int32_t* assign, int64_t stride_instance, int64_t stride_rule,
                       const float* remaining, int ld_instance, int ld_row, int B, int num_bins, int num_resources,
                       float* dominant, int32_t* rejected, int32_t* empty_nodes, void* stream);
/* replaces gather_stats_from_solutions (misc/utils.py:97-159): heur_* are (num_heuristics, B). Adds to
 * results[0..2] the dominant {won, draw, lost} and to results[3..5] the rejected {won, draw, lost} counts of the
 * net against the best heuristic; net_dominant_rounded (B, optional) = round_half_up(net_dominant, 2). */
int tpc_compare_solutions(tpc_handle* h, const float* net_dominant, const int32_t* net_rejected,
                          const float* heur_dominant, const int32_t* heur_rejected, int num_heuristics, int B,
                          float* net_dominant_rounded, int32_t* results, void* stream);

/* KnapsackV2 tester (environment/custom/knapsack_v2/tester.py, heuristic/*, misc/utils.py:41-95). state rows: bins
 * (capacity, current_load) with row 0 = EOS bin, then items (weight, value). assign (B, num_items) int32 = bin per
 * item (0 = rejected), order (B, num_items) = item placed k-th, bins (B, num_bins, 2) = (capacity, load) afterwards,
 * values (B, num_bins) = Bin.current_value (fp32 sums in placement order). */
/* replaces WasteReductionHeuristic.solve (heuristic/waste_reduction.py:26-72, bin.py:34-60, item.py:14) */
int tpc_knapsack_waste_reduction(tpc_handle* h, const float* state, int ld_instance, int ld_row, int B, int num_bins,
                                 int num_items, int item_sort_descending, int bin_sort_descending, int32_t* assign,
                                 int32_t* order, float* bins, float* values, void* stream);
/* replaces RandomHeuristic.solve (knapsack_v2/heuristic/random_heuristic.py:24-52); draws as tpc_heuristic_random */
int tpc_knapsack_random(tpc_handle* h, const float* state, int ld_instance, int ld_row, int B, int num_bins,
                        int num_items, uint64_t seed, int64_t instance_id0, int64_t epoch, int32_t* assign,
                        int32_t* order, float* bins, float* values, void* stream);
/* Bin.current_value per bin for a placement made in item order (the net's history, knapsack_v2/env.py); assign is
 * read as assign[b * stride_instance + j * stride_item]; item values come from state */
int tpc_knapsack_bin_values(tpc_handle* h, const int32_t* assign, int64_t stride_instance, int64_t stride_item,
                            const float* state, int ld_instance, int ld_row, int B, int num_bins, int num_items,
                            float* values, void* stream);
/* replaces compute_stats (knapsack_v2/misc/utils.py:41-58): reward = sum of the real bins' values, empty bins,
 * rejected items, value collected by the EOS bin. The won / draw / lost verdict (misc/utils.py:60-95) is
 * tpc_compare_solutions with the rewards in the `dominant` slots. */
int tpc_knapsack_stats(tpc_handle* h, const int32_t* assign, int64_t stride_instance, int64_t stride_item,
                       const float* values, int B, int num_bins, int num_items, float* reward, int32_t* empty_bins,
                       int32_t* rejected, float* rejected_value, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * step-wise loss arithmetic (the API-exact mode of Agent; tpc_a2c_grads does the same inside the fused update)
 * ---------------------------------------------------------------------------------------------------- */
/* replaces Agent.compute_discounted_rewards (agents/agent.py:110-124): rewards, returns (B,T); bootstrap (B) = value
 * after the last step, NULL = 0 (agents/trainer.py:103-105). */
int tpc_discounted_returns(tpc_handle* h, const float* rewards, const float* bootstrap, int B, int T, float gamma,
                           float* returns, void* stream);
/* replaces the arithmetic of Agent.compute_value_loss (agents/agent.py:155-164): returns_v, values, advantages (n)
 * step-major; loss[0] = c_v * mean((R - V)^2). */
int tpc_value_loss(tpc_handle* h, const float* returns_v, const float* values, int n, float c_v, float* advantages,
                   float* loss, void* stream);
/* replaces the arithmetic of Agent.compute_actor_loss (agents/agent.py:189-240): policy_loss (n) = sparse softmax
 * cross-entropy of the taken action, entropy (n) (both optional); sums[3] = mean policy loss,
 * mean(policy_loss * advantage - c_e * entropy), mean entropy. */
int tpc_actor_loss(tpc_handle* h, const float* logits, const int32_t* actions, const float* advantages, int n, int S,
                   float c_e, float* policy_loss, float* entropy, float* sums, void* stream);
/* replaces the batch statistics of the training loop (agents/trainer.py:89-100): out[3] = min, max, mean over the
 * batch of the per-episode returns (sum of the T step rewards of rewards (B,T)). */
int tpc_episode_reward_stats(tpc_handle* h, const float* rewards, int B, int T, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TPC_H_ */
