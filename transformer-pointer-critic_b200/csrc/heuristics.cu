// Batched placement heuristics and solution statistics of the ResourceV3 tester, one warp per instance.
// Bit-exact restatement of the reference's per-instance Python objects:
//   DominantResourceHeuristic.solve / place_single_resource  environment/custom/resource_v3/heuristic/dominant_heuristic.py:30-71
//   RandomHeuristic.solve / place_single_resource            environment/custom/resource_v3/heuristic/random_heuristic.py:28-65
//   Node.compute_dominant_resource / insert_req              environment/custom/resource_v3/node.py:36-70
//   compute_stats, gather_stats_from_solutions               environment/custom/resource_v3/misc/utils.py:56-159
// The reference asserts batch == 1 (base_heuristic.py:39,59); here every instance of the batch is solved.
// Float ops are explicit round-to-nearest intrinsics in the reference's order (no FMA contraction).
#include "handle.cuh"
#include "philox.cuh"

namespace tpc {

namespace {

constexpr int WARPS = 4;

// Node.compute_dominant_resource (node.py:36-41): min over the 3 features of round_half_up(remaining - request, 2)
__device__ __forceinline__ float dominant_diff(const float* rem, const float* req) {
  const float d0 = rhu2(__fsub_rn(rem[0], req[0]));
  const float d1 = rhu2(__fsub_rn(rem[1], req[1]));
  const float d2 = rhu2(__fsub_rn(rem[2], req[2]));
  return fminf(fminf(d0, d1), d2);
}

// shared memory of one warp: rem[N][3] | rules[R][3] | list_a[R] | list_b[N]
__device__ __forceinline__ void load_instance(const float* __restrict__ inst, int ld_row, int N, int R, float* rem,
                                              float* rules, int lane) {
  for (int i = lane; i < N * 3; i += 32) rem[i] = inst[(size_t)(i / 3) * ld_row + (i % 3)];
  for (int i = lane; i < R * 3; i += 32) rules[i] = inst[(size_t)(N + i / 3) * ld_row + (i % 3)];
  __syncwarp();
}

__global__ void __launch_bounds__(WARPS * 32)
dominant_kernel(const float* __restrict__ state, int ld_inst, int ld_row, int B, int N, int R, int res_desc,
                int node_desc, int32_t* __restrict__ assign, float* __restrict__ remaining) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * WARPS + warp;
  if (b >= B) return;
  const int per_warp = 3 * N + 3 * R + R;
  float* rem = smem + (size_t)warp * per_warp;
  float* rules = rem + 3 * N;
  int* order = reinterpret_cast<int*>(rules + 3 * R);
  load_instance(state + (size_t)b * ld_inst, ld_row, N, R, rem, rules, lane);

  // sorted(resource_list, key=max(CPU, RAM, MEM), reverse=res_desc): stable in both directions, so the position of
  // rule j is the number of rules that sort strictly before it plus the equal-keyed rules with a smaller index
  for (int j = lane; j < R; j += 32) {
    const float kj = fmaxf(fmaxf(rules[3 * j], rules[3 * j + 1]), rules[3 * j + 2]);
    int pos = 0;
    for (int i = 0; i < R; ++i) {
      const float ki = fmaxf(fmaxf(rules[3 * i], rules[3 * i + 1]), rules[3 * i + 2]);
      pos += (res_desc ? ki > kj : ki < kj) || (ki == kj && i < j);
    }
    order[pos] = j;
  }
  __syncwarp();

  int32_t* out = assign + (size_t)b * R;
  for (int t = 0; t < R; ++t) {
    const int j = order[t];
    const float* req = rules + 3 * j;
    // first feasible node of the (stably) sorted diff list = feasible node with the extreme diff, lowest id on ties
    float best = 0.f;
    int best_n = 0;  // 0 = none: the rule goes to the EOS node
    for (int n = 1 + lane; n < N; n += 32) {
      const float d = dominant_diff(rem + 3 * n, req);
      if (d >= 0.f && (best_n == 0 || (node_desc ? d > best : d < best))) {
        best = d;
        best_n = n;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int on = __shfl_xor_sync(0xffffffffu, best_n, o);
      const bool take = on != 0 && (best_n == 0 || (node_desc ? ob > best : ob < best) || (ob == best && on < best_n));
      if (take) {
        best = ob;
        best_n = on;
      }
    }
    if (best_n != 0 && lane < 3) rem[3 * best_n + lane] = rhu2(__fsub_rn(rem[3 * best_n + lane], req[lane]));  // node.py:64-66
    if (lane == 0) out[j] = best_n;
    __syncwarp();
  }
  float* ro = remaining + (size_t)b * N * 3;
  for (int i = lane; i < N * 3; i += 32) ro[i] = rem[i];
}

// remove element idx of a list held in shared memory (Python list.pop(idx)); all lanes return the popped value
__device__ __forceinline__ int list_pop(int* list, int len, int idx, int lane) {
  const int v = list[idx];
  __syncwarp();
  for (int k0 = idx; k0 < len - 1; k0 += 32) {
    const int k = k0 + lane;
    const int nxt = k < len - 1 ? list[k + 1] : 0;
    __syncwarp();
    if (k < len - 1) list[k] = nxt;
    __syncwarp();
  }
  return v;
}

__global__ void __launch_bounds__(WARPS * 32)
random_kernel(const float* __restrict__ state, int ld_inst, int ld_row, int B, int N, int R, uint64_t seed,
              int64_t id0, int64_t epoch, int32_t* __restrict__ assign, float* __restrict__ remaining) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * WARPS + warp;
  if (b >= B) return;
  const int per_warp = 3 * N + 3 * R + R + N;
  float* rem = smem + (size_t)warp * per_warp;
  float* rules = rem + 3 * N;
  int* left = reinterpret_cast<int*>(rules + 3 * R);
  int* nodes = left + R;
  load_instance(state + (size_t)b * ld_inst, ld_row, N, R, rem, rules, lane);
  for (int j = lane; j < R; j += 32) left[j] = j;
  __syncwarp();
  // one Philox stream per instance; draw i answers the i-th random.randrange(n) call as floor(u32 * n / 2^32).
  // Every lane evaluates the same stream so no broadcast is needed.
  Stream s = make_stream(seed, id0 + b, epoch, 4);
  uint32_t draw = 0;
  int32_t* out = assign + (size_t)b * R;
  for (int n_left = R; n_left > 0; --n_left) {
    const int j = list_pop(left, n_left, (int)__umulhi(s.draw(draw++), (uint32_t)n_left), lane);
    const float* req = rules + 3 * j;
    for (int n = lane; n < N - 1; n += 32) nodes[n] = n + 1;  // copy_list = [] + node_list (random_heuristic.py:39)
    __syncwarp();
    int placed = 0;
    for (int n_nodes = N - 1; n_nodes > 0; --n_nodes) {
      const int n = list_pop(nodes, n_nodes, (int)__umulhi(s.draw(draw++), (uint32_t)n_nodes), lane);
      if (dominant_diff(rem + 3 * n, req) >= 0.f) {  // node.can_fit_resource (node.py:43-46)
        __syncwarp();
        if (lane < 3) rem[3 * n + lane] = rhu2(__fsub_rn(rem[3 * n + lane], req[lane]));
        placed = n;
        break;
      }
    }
    __syncwarp();
    if (lane == 0) out[j] = placed;
  }
  float* ro = remaining + (size_t)b * N * 3;
  for (int i = lane; i < N * 3; i += 32) ro[i] = rem[i];
}

// compute_stats (misc/utils.py:56-81) of one finished placement per warp
__global__ void __launch_bounds__(WARPS * 32)
stats_kernel(const int32_t* __restrict__ assign, long long sa_inst, long long sa_rule, const float* __restrict__ remaining,
             int ld_inst, int ld_row, int B, int N, int R, float* __restrict__ dominant, int32_t* __restrict__ rejected,
             int32_t* __restrict__ empty_nodes) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * WARPS + warp;
  if (b >= B) return;
  int* used = reinterpret_cast<int*>(smem) + (size_t)warp * N;
  for (int n = lane; n < N; n += 32) used[n] = 0;
  __syncwarp();
  int rej = 0;
  for (int j = lane; j < R; j += 32) {
    const int a = assign[(size_t)b * sa_inst + (size_t)j * sa_rule];
    rej += a == 0;
    if (a >= 0 && a < N) used[a] = 1;  // benign race: every writer stores 1
  }
  __syncwarp();
  const float* rm = remaining + (size_t)b * ld_inst;
  float mn = INFINITY;
  int empty = 0;
  for (int n = 1 + lane; n < N; n += 32) {
    const float* r = rm + (size_t)n * ld_row;
    mn = fminf(mn, fminf(fminf(r[0], r[1]), r[2]));
    empty += used[n] == 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    empty += __shfl_xor_sync(0xffffffffu, empty, o);
    rej += __shfl_xor_sync(0xffffffffu, rej, o);
  }
  if (lane == 0) {
    dominant[b] = mn;
    rejected[b] = rej;
    empty_nodes[b] = empty;
  }
}

// gather_stats_from_solutions (misc/utils.py:97-159): net vs the best heuristic, per instance.
// results[0..2] += dominant {won, draw, lost}; results[3..5] += rejected {won, draw, lost}.
__global__ void compare_kernel(const float* __restrict__ net_dom, const int32_t* __restrict__ net_rej,
                               const float* __restrict__ heur_dom, const int32_t* __restrict__ heur_rej, int H, int B,
                               float* __restrict__ net_dom_rounded, int32_t* __restrict__ results) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  int code_d = -1, code_r = -1;
  if (b < B) {
    const float nd = rhu2(net_dom[b]);  // utils.py:107
    const int nr = net_rej[b];
    float max_dom = -1.f;               // utils.py:116-117
    int min_rej = 9999;
    for (int h = 0; h < H; ++h) {
      max_dom = fmaxf(max_dom, rhu2(heur_dom[(size_t)h * B + b]));
      min_rej = min(min_rej, heur_rej[(size_t)h * B + b]);
    }
    code_d = max_dom > nd ? 2 : (max_dom == nd ? 1 : 0);
    code_r = min_rej < nr ? 2 : (min_rej == nr ? 1 : 0);
    if (net_dom_rounded) net_dom_rounded[b] = nd;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const unsigned md = __ballot_sync(0xffffffffu, code_d == c), mr = __ballot_sync(0xffffffffu, code_r == c);
    if ((threadIdx.x & 31) == 0) {
      if (md) atomicAdd(results + c, __popc(md));
      if (mr) atomicAdd(results + 3 + c, __popc(mr));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// KnapsackV2 (environment/custom/knapsack_v2/heuristic/{waste_reduction,random_heuristic}.py, bin.py:34-60, item.py,
// misc/utils.py:41-58). state rows: bins (capacity, current_load), items (weight, value).
// shared memory of one warp: cap[N] | load[N] | val[N] | w[R] | v[R] | list_a[R] | list_b[N]
// ------------------------------------------------------------------------------------------------------------------
struct KnSmem {
  float *cap, *load, *val, *w, *v;
  int *la, *lb;
};
__device__ __forceinline__ KnSmem kn_carve(float* base, int N, int R) {
  KnSmem k;
  k.cap = base; k.load = k.cap + N; k.val = k.load + N; k.w = k.val + N; k.v = k.w + R;
  k.la = reinterpret_cast<int*>(k.v + R); k.lb = k.la + R;
  return k;
}
__host__ __device__ inline int kn_floats(int N, int R) { return 4 * N + 3 * R; }

__device__ __forceinline__ void kn_load(const float* __restrict__ inst, int ld_row, int N, int R, const KnSmem& k, int lane) {
  for (int n = lane; n < N; n += 32) {
    k.cap[n] = inst[(size_t)n * ld_row];
    k.load[n] = inst[(size_t)n * ld_row + 1];
    k.val[n] = 0.f;
  }
  for (int j = lane; j < R; j += 32) {
    k.w[j] = inst[(size_t)(N + j) * ld_row];
    k.v[j] = inst[(size_t)(N + j) * ld_row + 1];
  }
  __syncwarp();
}
// Bin.compute_remaining_capacity (bin.py:44-47)
__device__ __forceinline__ float kn_remaining(float cap, float load, float w) {
  return rhu2(__fsub_rn(cap, rhu2(__fadd_rn(load, w))));
}
// Bin.insert_item (bin.py:49-60) by lane 0; bin 0 (EOS) only collects the value
__device__ __forceinline__ void kn_insert(const KnSmem& k, int n, int j) {
  if (n != 0) k.load[n] = rhu2(__fadd_rn(k.load[n], k.w[j]));
  k.val[n] = __fadd_rn(k.val[n], k.v[j]);
}
__device__ __forceinline__ void kn_store(const KnSmem& k, int N, float* __restrict__ bins_out, float* __restrict__ values_out,
                                         int lane) {
  for (int n = lane; n < N; n += 32) {
    bins_out[2 * n] = k.cap[n];
    bins_out[2 * n + 1] = k.load[n];
    values_out[n] = k.val[n];
  }
}

__global__ void __launch_bounds__(WARPS * 32)
knapsack_wr_kernel(const float* __restrict__ state, int ld_inst, int ld_row, int B, int N, int R, int item_desc,
                   int bin_desc, int32_t* __restrict__ assign, int32_t* __restrict__ order_out,
                   float* __restrict__ bins_out, float* __restrict__ values_out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * WARPS + warp;
  if (b >= B) return;
  const KnSmem k = kn_carve(smem + (size_t)warp * kn_floats(N, R), N, R);
  kn_load(state + (size_t)b * ld_inst, ld_row, N, R, k, lane);
  // sorted(item_list, key=value / weight, reverse=item_desc), stable (waste_reduction.py:36-40, item.py:14)
  for (int j = lane; j < R; j += 32) {
    const float kj = __fdiv_rn(k.v[j], k.w[j]);
    int pos = 0;
    for (int i = 0; i < R; ++i) {
      const float ki = __fdiv_rn(k.v[i], k.w[i]);
      pos += (item_desc ? ki > kj : ki < kj) || (ki == kj && i < j);
    }
    k.la[pos] = j;
  }
  __syncwarp();
  for (int t = 0; t < R; ++t) {
    const int j = k.la[t];
    const float w = k.w[j];
    float best = 0.f;
    int best_n = 0;
    for (int n = 1 + lane; n < N; n += 32) {
      const float d = kn_remaining(k.cap[n], k.load[n], w);
      if (d >= 0.f && (best_n == 0 || (bin_desc ? d > best : d < best))) {
        best = d;
        best_n = n;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int on = __shfl_xor_sync(0xffffffffu, best_n, o);
      const bool take = on != 0 && (best_n == 0 || (bin_desc ? ob > best : ob < best) || (ob == best && on < best_n));
      if (take) {
        best = ob;
        best_n = on;
      }
    }
    if (lane == 0) {
      kn_insert(k, best_n, j);
      assign[(size_t)b * R + j] = best_n;
      order_out[(size_t)b * R + t] = j;
    }
    __syncwarp();
  }
  kn_store(k, N, bins_out + (size_t)b * N * 2, values_out + (size_t)b * N, lane);
}

__global__ void __launch_bounds__(WARPS * 32)
knapsack_random_kernel(const float* __restrict__ state, int ld_inst, int ld_row, int B, int N, int R, uint64_t seed,
                       int64_t id0, int64_t epoch, int32_t* __restrict__ assign, int32_t* __restrict__ order_out,
                       float* __restrict__ bins_out, float* __restrict__ values_out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * WARPS + warp;
  if (b >= B) return;
  const KnSmem k = kn_carve(smem + (size_t)warp * kn_floats(N, R), N, R);
  kn_load(state + (size_t)b * ld_inst, ld_row, N, R, k, lane);
  for (int j = lane; j < R; j += 32) k.la[j] = j;
  __syncwarp();
  Stream s = make_stream(seed, id0 + b, epoch, 4);
  uint32_t draw = 0;
  for (int n_left = R; n_left > 0; --n_left) {
    const int j = list_pop(k.la, n_left, (int)__umulhi(s.draw(draw++), (uint32_t)n_left), lane);
    for (int n = lane; n < N - 1; n += 32) k.lb[n] = n + 1;
    __syncwarp();
    int placed = 0;
    for (int n_bins = N - 1; n_bins > 0; --n_bins) {
      const int n = list_pop(k.lb, n_bins, (int)__umulhi(s.draw(draw++), (uint32_t)n_bins), lane);
      if (kn_remaining(k.cap[n], k.load[n], k.w[j]) >= 0.f) {   // Bin.can_fit_item (bin.py:40-42)
        placed = n;
        break;
      }
    }
    __syncwarp();
    if (lane == 0) {
      kn_insert(k, placed, j);
      assign[(size_t)b * R + j] = placed;
      order_out[(size_t)b * R + (R - n_left)] = j;
    }
    __syncwarp();
  }
  kn_store(k, N, bins_out + (size_t)b * N * 2, values_out + (size_t)b * N, lane);
}

// Bin.current_value of every bin for a placement made in item order (the net's decoding order): lane = bin, items
// scanned in order, so every bin's fp32 sum has the reference's operation sequence.
__global__ void __launch_bounds__(WARPS * 32)
knapsack_bin_values_kernel(const int32_t* __restrict__ assign, long long sa_inst, long long sa_item,
                           const float* __restrict__ state, int ld_inst, int ld_row, int B, int N, int R,
                           float* __restrict__ values_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * WARPS + warp;
  if (b >= B) return;
  const float* inst = state + (size_t)b * ld_inst;
  for (int n = lane; n < N; n += 32) {
    float acc = 0.f;
    for (int j = 0; j < R; ++j)
      if (assign[(size_t)b * sa_inst + (size_t)j * sa_item] == n) acc = __fadd_rn(acc, inst[(size_t)(N + j) * ld_row + 1]);
    values_out[(size_t)b * N + n] = acc;
  }
}

// compute_stats (knapsack_v2/misc/utils.py:41-58)
__global__ void __launch_bounds__(WARPS * 32)
knapsack_stats_kernel(const int32_t* __restrict__ assign, long long sa_inst, long long sa_item,
                      const float* __restrict__ values, int B, int N, int R, float* __restrict__ reward,
                      int32_t* __restrict__ empty_bins, int32_t* __restrict__ rejected, float* __restrict__ rejected_value) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * WARPS + warp;
  if (b >= B) return;
  int* used = reinterpret_cast<int*>(smem) + (size_t)warp * N;
  for (int n = lane; n < N; n += 32) used[n] = 0;
  __syncwarp();
  int rej = 0;
  for (int j = lane; j < R; j += 32) {
    const int a = assign[(size_t)b * sa_inst + (size_t)j * sa_item];
    rej += a == 0;
    if (a >= 0 && a < N) used[a] = 1;
  }
  __syncwarp();
  int empty = 0;
  for (int n = 1 + lane; n < N; n += 32) empty += used[n] == 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    empty += __shfl_xor_sync(0xffffffffu, empty, o);
    rej += __shfl_xor_sync(0xffffffffu, rej, o);
  }
  if (lane == 0) {
    float r = 0.f;   // reward = 0; reward += node.current_value, bins in order
    for (int n = 1; n < N; ++n) r = __fadd_rn(r, values[(size_t)b * N + n]);
    reward[b] = r;
    empty_bins[b] = empty;
    rejected[b] = rej;
    rejected_value[b] = values[(size_t)b * N];
  }
}

int check_shape(int B, int N, int R, size_t smem_bytes) {
  if (B < 0 || N < 2 || R < 1) return TPC_ERR_SHAPE;
  if (smem_bytes > 200 * 1024) return TPC_ERR_SHAPE;
  return TPC_OK;
}

template <typename K>
int opt_in_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) TPC_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return TPC_OK;
}

}  // namespace

}  // namespace tpc

using namespace tpc;

extern "C" {

int tpc_heuristic_dominant(tpc_handle* h, const float* state, int ld_instance, int ld_row, int B, int num_bins,
                           int num_resources, int resource_sort_descending, int node_sort_descending, int32_t* assign,
                           float* remaining, void* stream) {
  if (!h || !state || !assign || !remaining) return TPC_ERR_ARG;
  const int N = num_bins, R = num_resources;
  const size_t smem = (size_t)WARPS * (3 * N + 3 * R + R) * sizeof(float);
  TPC_TRY(check_shape(B, N, R, smem));
  if (B == 0) return TPC_OK;
  TPC_TRY(opt_in_smem(dominant_kernel, smem));
  dominant_kernel<<<ceil_div(B, WARPS), WARPS * 32, smem, (cudaStream_t)stream>>>(
      state, ld_instance, ld_row, B, N, R, resource_sort_descending != 0, node_sort_descending != 0, assign, remaining);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int tpc_heuristic_random(tpc_handle* h, const float* state, int ld_instance, int ld_row, int B, int num_bins,
                         int num_resources, uint64_t seed, int64_t instance_id0, int64_t epoch, int32_t* assign,
                         float* remaining, void* stream) {
  if (!h || !state || !assign || !remaining) return TPC_ERR_ARG;
  const int N = num_bins, R = num_resources;
  const size_t smem = (size_t)WARPS * (3 * N + 3 * R + R + N) * sizeof(float);
  TPC_TRY(check_shape(B, N, R, smem));
  if (B == 0) return TPC_OK;
  TPC_TRY(opt_in_smem(random_kernel, smem));
  random_kernel<<<ceil_div(B, WARPS), WARPS * 32, smem, (cudaStream_t)stream>>>(
      state, ld_instance, ld_row, B, N, R, seed, instance_id0, epoch, assign, remaining);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int tpc_solution_stats(tpc_handle* h, const int32_t* assign, int64_t stride_instance, int64_t stride_rule,
                       const float* remaining, int ld_instance, int ld_row, int B, int num_bins, int num_resources,
                       float* dominant, int32_t* rejected, int32_t* empty_nodes, void* stream) {
  if (!h || !assign || !remaining || !dominant || !rejected || !empty_nodes) return TPC_ERR_ARG;
  const size_t smem = (size_t)WARPS * num_bins * sizeof(int);
  TPC_TRY(check_shape(B, num_bins, num_resources, smem));
  if (B == 0) return TPC_OK;
  TPC_TRY(opt_in_smem(stats_kernel, smem));
  stats_kernel<<<ceil_div(B, WARPS), WARPS * 32, smem, (cudaStream_t)stream>>>(
      assign, stride_instance, stride_rule, remaining, ld_instance, ld_row, B, num_bins, num_resources, dominant,
      rejected, empty_nodes);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int tpc_compare_solutions(tpc_handle* h, const float* net_dominant, const int32_t* net_rejected,
                          const float* heur_dominant, const int32_t* heur_rejected, int num_heuristics, int B,
                          float* net_dominant_rounded, int32_t* results, void* stream) {
  if (!h || !net_dominant || !net_rejected || !results || num_heuristics < 0) return TPC_ERR_ARG;
  if (num_heuristics > 0 && (!heur_dominant || !heur_rejected)) return TPC_ERR_ARG;
  if (B <= 0) return B == 0 ? TPC_OK : TPC_ERR_SHAPE;
  compare_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(net_dominant, net_rejected, heur_dominant,
                                                                      heur_rejected, num_heuristics, B,
                                                                      net_dominant_rounded, results);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int tpc_knapsack_waste_reduction(tpc_handle* h, const float* state, int ld_instance, int ld_row, int B, int num_bins,
                                 int num_items, int item_sort_descending, int bin_sort_descending, int32_t* assign,
                                 int32_t* order, float* bins, float* values, void* stream) {
  if (!h || !state || !assign || !order || !bins || !values) return TPC_ERR_ARG;
  const size_t smem = (size_t)WARPS * kn_floats(num_bins, num_items) * sizeof(float);
  TPC_TRY(check_shape(B, num_bins, num_items, smem));
  if (B == 0) return TPC_OK;
  TPC_TRY(opt_in_smem(knapsack_wr_kernel, smem));
  knapsack_wr_kernel<<<ceil_div(B, WARPS), WARPS * 32, smem, (cudaStream_t)stream>>>(
      state, ld_instance, ld_row, B, num_bins, num_items, item_sort_descending != 0, bin_sort_descending != 0, assign,
      order, bins, values);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int tpc_knapsack_random(tpc_handle* h, const float* state, int ld_instance, int ld_row, int B, int num_bins,
                        int num_items, uint64_t seed, int64_t instance_id0, int64_t epoch, int32_t* assign,
                        int32_t* order, float* bins, float* values, void* stream) {
  if (!h || !state || !assign || !order || !bins || !values) return TPC_ERR_ARG;
  const size_t smem = (size_t)WARPS * kn_floats(num_bins, num_items) * sizeof(float);
  TPC_TRY(check_shape(B, num_bins, num_items, smem));
  if (B == 0) return TPC_OK;
  TPC_TRY(opt_in_smem(knapsack_random_kernel, smem));
  knapsack_random_kernel<<<ceil_div(B, WARPS), WARPS * 32, smem, (cudaStream_t)stream>>>(
      state, ld_instance, ld_row, B, num_bins, num_items, seed, instance_id0, epoch, assign, order, bins, values);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int tpc_knapsack_bin_values(tpc_handle* h, const int32_t* assign, int64_t stride_instance, int64_t stride_item,
                            const float* state, int ld_instance, int ld_row, int B, int num_bins, int num_items,
                            float* values, void* stream) {
  if (!h || !assign || !state || !values) return TPC_ERR_ARG;
  TPC_TRY(check_shape(B, num_bins, num_items, 0));
  if (B == 0) return TPC_OK;
  knapsack_bin_values_kernel<<<ceil_div(B, WARPS), WARPS * 32, 0, (cudaStream_t)stream>>>(
      assign, stride_instance, stride_item, state, ld_instance, ld_row, B, num_bins, num_items, values);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int tpc_knapsack_stats(tpc_handle* h, const int32_t* assign, int64_t stride_instance, int64_t stride_item,
                       const float* values, int B, int num_bins, int num_items, float* reward, int32_t* empty_bins,
                       int32_t* rejected, float* rejected_value, void* stream) {
  if (!h || !assign || !values || !reward || !empty_bins || !rejected || !rejected_value) return TPC_ERR_ARG;
  const size_t smem = (size_t)WARPS * num_bins * sizeof(int);
  TPC_TRY(check_shape(B, num_bins, num_items, smem));
  if (B == 0) return TPC_OK;
  TPC_TRY(opt_in_smem(knapsack_stats_kernel, smem));
  knapsack_stats_kernel<<<ceil_div(B, WARPS), WARPS * 32, smem, (cudaStream_t)stream>>>(
      assign, stride_instance, stride_item, values, B, num_bins, num_items, reward, empty_bins, rejected, rejected_value);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

}  // extern "C"
