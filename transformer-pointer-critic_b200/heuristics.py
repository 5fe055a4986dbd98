"""Batched placement heuristics and solution statistics for the ResourceV3 tester, with the reference's names and
option keys (environment/custom/resource_v3/heuristic/{factory,base_heuristic,dominant_heuristic,random_heuristic}.py,
misc/utils.py:56-159). The reference solves ONE instance with Python Node / Resource objects and asserts
batch == 1 (base_heuristic.py:39,59); here `solve(state)` takes the whole (B, S, F) batch and runs one warp per
instance on the device (libtpc: tpc_heuristic_*, tpc_solution_stats, tpc_compare_solutions). No CPU fallback.

`solver.solution` (a list of Node objects in the reference) becomes a `Solution` holding device tensors:
assign (B, R) int32 (node id per rule, 0 = rejected) and remaining (B, N, 3).
The CPLEX solvers are out of scope (CPLEX is not installable): asking for them raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import lib as L

_handles: dict = {}


def _handle(device: torch.device):
    """One libtpc handle per device, created on first use."""
    if not torch.cuda.is_available():
        raise L.TpcError("tpc_b200 heuristics need a CUDA device (sm_100a); there is no CPU fallback")
    key = torch.device(device).index or 0
    if key not in _handles:
        lib = L.load()
        h = C.c_void_p()
        with torch.cuda.device(key):
            L.check(lib.tpc_create(C.byref(h)), "tpc_create")
        _handles[key] = (lib, h)
    return _handles[key]


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


@dataclass
class Solution:
    assign: torch.Tensor      # (B, R) int32
    remaining: torch.Tensor   # (B, N, 3) fp32


def _as_device_state(state, device) -> torch.Tensor:
    t = state if torch.is_tensor(state) else torch.as_tensor(np.asarray(state, dtype=np.float32))
    t = t.to(device=device, dtype=torch.float32)
    if t.dim() != 3 or t.shape[2] < 3:
        raise ValueError(f"state must be (batch, nodes + rules, >= 3 features), got {tuple(t.shape)}")
    return t.contiguous()


class BaseHeuristic:
    """heuristic/base_heuristic.py:6-33."""

    def __init__(self, num_nodes: int, normalization_factor: int, device="cuda:0"):
        self.name = "base_heuristic"
        self.is_optimal = 0
        self.resource_batch_id = 0
        self.num_nodes = num_nodes
        self.normalization_factor = normalization_factor
        self.device = torch.device(device)
        self.solution = []

    def reset(self):
        self.resource_batch_id = 0
        self.solution = []

    def solve(self, state):
        raise NotImplementedError('Method "solve" not implemented')

    def _outputs(self, state: torch.Tensor):
        B, S = state.shape[0], state.shape[1]
        R = S - self.num_nodes
        if R < 1:
            raise ValueError("state holds no rules")
        assign = torch.empty(B, R, dtype=torch.int32, device=self.device)
        remaining = torch.empty(B, self.num_nodes, 3, dtype=torch.float32, device=self.device)
        return B, R, assign, remaining

    def stats(self):
        """compute_stats (misc/utils.py:56-81) of every instance: (dominant (B,) fp32, rejected (B,), empty (B,))."""
        return compute_stats(self.solution)


class DominantResourceHeuristic(BaseHeuristic):
    """heuristic/dominant_heuristic.py:13-71."""

    def __init__(self, num_nodes: int, normalization_factor: int, opts: dict, device="cuda:0"):
        super().__init__(num_nodes, normalization_factor, device)
        self.resource_sort_descending = opts["resource_sort_descending"]
        self.node_sort_descending = opts["node_sort_descending"]
        self.generate_name()

    def generate_name(self):
        self.name = f"dominant_resource_ASC_{self.resource_sort_descending}_node_ASC_{self.node_sort_descending}"

    def solve(self, state):
        st = _as_device_state(state, self.device)
        lib, h = _handle(self.device)
        B, R, assign, remaining = self._outputs(st)
        L.check(lib.tpc_heuristic_dominant(h, L.ptr(st), st.stride(0), st.stride(1), B, self.num_nodes, R,
                                           1 if self.resource_sort_descending else 0,
                                           1 if self.node_sort_descending else 0, L.ptr(assign), L.ptr(remaining),
                                           _stream(self.device)), "tpc_heuristic_dominant")
        self.resource_batch_id += 1
        self.solution = Solution(assign, remaining)


class RandomHeuristic(BaseHeuristic):
    """heuristic/random_heuristic.py:14-65. Python's `random` stream becomes a Philox stream per instance keyed by
    (seed, instance_id0 + b, number of earlier solve() calls)."""

    def __init__(self, num_nodes: int, normalization_factor: int, opts: dict, device="cuda:0", seed: int = 1235,
                 instance_id0: int = 0):
        super().__init__(num_nodes, normalization_factor, device)
        self.seed, self.instance_id0 = int(seed), int(instance_id0)
        self.generate_name()

    def generate_name(self):
        self.name = "random"

    def solve(self, state):
        st = _as_device_state(state, self.device)
        lib, h = _handle(self.device)
        B, R, assign, remaining = self._outputs(st)
        L.check(lib.tpc_heuristic_random(h, L.ptr(st), st.stride(0), st.stride(1), B, self.num_nodes, R, self.seed,
                                         self.instance_id0, self.resource_batch_id, L.ptr(assign), L.ptr(remaining),
                                         _stream(self.device)), "tpc_heuristic_random")
        self.resource_batch_id += 1
        self.solution = Solution(assign, remaining)


def generate_dominant_combos(num_nodes: int, normalization_factor: int, opts: dict, device="cuda:0"):
    """heuristic/factory.py:44-71."""
    if not opts["generate_params_combos"]:
        return [DominantResourceHeuristic(num_nodes, normalization_factor, opts, device)]
    return [DominantResourceHeuristic(num_nodes, normalization_factor,
                                      {"resource_sort_descending": r, "node_sort_descending": n}, device)
            for r in (True, False) for n in (True, False)]


def heuristic_factory(num_nodes: int, normalization_factor: int, opts: dict, device="cuda:0", seed: int = 1235,
                      instance_id0: int = 0):
    """heuristic/factory.py:16-42 (same order: dominant combos, random, [cplex])."""
    for key in ("cplex_greedy_and_critical", "cplex_node_reduction"):
        if opts.get(key, {}).get("use"):
            raise NotImplementedError(f"heuristic '{key}' needs IBM CPLEX and is out of scope of tpc_b200")
    return generate_dominant_combos(num_nodes, normalization_factor, opts["dominant_resource"], device) + \
        [RandomHeuristic(num_nodes, normalization_factor, opts["random"], device, seed, instance_id0)]


def compute_stats(solution: Solution):
    """compute_stats (misc/utils.py:56-81), batched: (dominant (B,) fp32 unrounded, rejected (B,), empty (B,))."""
    return solution_stats(solution.assign, solution.assign.stride(0), solution.assign.stride(1), solution.remaining)


def solution_stats(assign: torch.Tensor, stride_instance: int, stride_rule: int, remaining: torch.Tensor):
    """`remaining` is (B, >= N rows, >= 3 features) with row 0 = EOS node and N = remaining.shape[1] rows scanned
    (pass a narrowed view of the env batch for the net's solution)."""
    device = remaining.device
    lib, h = _handle(device)
    B, N = remaining.shape[0], remaining.shape[1]
    R = assign.numel() // B
    dominant = torch.empty(B, dtype=torch.float32, device=device)
    rejected = torch.empty(B, dtype=torch.int32, device=device)
    empty = torch.empty(B, dtype=torch.int32, device=device)
    L.check(lib.tpc_solution_stats(h, L.ptr(assign), stride_instance, stride_rule, L.ptr(remaining),
                                   remaining.stride(0), remaining.stride(1), B, N, R, L.ptr(dominant),
                                   L.ptr(rejected), L.ptr(empty), _stream(device)), "tpc_solution_stats")
    return dominant, rejected, empty


def net_solution_stats(bufs, num_bins: int):
    """compute_stats of the net's placements straight from the rollout buffers: actions (T, B) step-major and the
    final environment state (B, S, F) (what env.history holds in the reference, env.py:354-384)."""
    return solution_stats(bufs.actions, 1, bufs.actions.stride(0), bufs.batch[:, :num_bins, :])


def compare_solutions(net_dominant, net_rejected, heur_dominant, heur_rejected):
    """gather_stats_from_solutions' verdict (misc/utils.py:116-159) for every instance, summed over the batch.
    heur_* are lists of (B,) tensors. Returns (net_dominant rounded (B,), results int32[6] device tensor:
    dominant won/draw/lost, rejected won/draw/lost)."""
    device = net_dominant.device
    lib, h = _handle(device)
    B, H = net_dominant.shape[0], len(heur_dominant)
    hd = torch.stack(list(heur_dominant)).contiguous() if H else None
    hr = torch.stack(list(heur_rejected)).contiguous() if H else None
    rounded = torch.empty(B, dtype=torch.float32, device=device)
    results = torch.zeros(6, dtype=torch.int32, device=device)
    L.check(lib.tpc_compare_solutions(h, L.ptr(net_dominant), L.ptr(net_rejected), L.ptr(hd), L.ptr(hr), H, B,
                                      L.ptr(rounded), L.ptr(results), _stream(device)), "tpc_compare_solutions")
    return rounded, results


# ---------------------------------------------------------------------------------------------------------------
# KnapsackV2 (environment/custom/knapsack_v2/heuristic/*, misc/utils.py:41-95)
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class KnapsackSolution:
    assign: torch.Tensor      # (B, R) int32: bin per item, 0 = rejected
    order: torch.Tensor       # (B, R) int32: item placed k-th
    bins: torch.Tensor        # (B, N, 2) fp32 (capacity, load)
    values: torch.Tensor      # (B, N) fp32 Bin.current_value


class _KnapsackBase(BaseHeuristic):
    """knapsack_v2/heuristic/base_heuristic.py (num_bins instead of num_nodes, no normalization factor)."""

    def __init__(self, num_bins: int, device="cuda:0"):
        super().__init__(num_bins, 1, device)
        self.num_bins = num_bins

    def _state(self, state):
        t = state if torch.is_tensor(state) else torch.as_tensor(np.asarray(state, dtype=np.float32))
        t = t.to(device=self.device, dtype=torch.float32)
        if t.dim() != 3 or t.shape[2] < 2 or t.shape[1] <= self.num_bins:
            raise ValueError(f"state must be (batch, bins + items, >= 2 features), got {tuple(t.shape)}")
        return t.contiguous()

    def _alloc(self, st):
        B, R = st.shape[0], st.shape[1] - self.num_bins
        i32 = lambda *s: torch.empty(*s, dtype=torch.int32, device=self.device)
        f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.device)
        return B, R, KnapsackSolution(i32(B, R), i32(B, R), f32(B, self.num_bins, 2), f32(B, self.num_bins))

    def stats(self):
        """compute_stats (knapsack_v2/misc/utils.py:41-58): (reward, empty bins, rejected items, rejected value)."""
        sol = self.solution
        return knapsack_stats(sol.assign, sol.assign.stride(0), sol.assign.stride(1), sol.values)


class WasteReductionHeuristic(_KnapsackBase):
    """knapsack_v2/heuristic/waste_reduction.py:13-72."""

    def __init__(self, num_bins: int, opts: dict, device="cuda:0"):
        super().__init__(num_bins, device)
        self.item_sort_descending = opts["item_sort_descending"]
        self.bin_sort_descending = opts["bin_sort_descending"]
        self.name = f"item_ASC_{self.item_sort_descending}_bin_ASC_{self.bin_sort_descending}"

    def solve(self, state):
        st = self._state(state)
        lib, h = _handle(self.device)
        B, R, sol = self._alloc(st)
        L.check(lib.tpc_knapsack_waste_reduction(h, L.ptr(st), st.stride(0), st.stride(1), B, self.num_bins, R,
                                                 1 if self.item_sort_descending else 0,
                                                 1 if self.bin_sort_descending else 0, L.ptr(sol.assign),
                                                 L.ptr(sol.order), L.ptr(sol.bins), L.ptr(sol.values),
                                                 _stream(self.device)), "tpc_knapsack_waste_reduction")
        self.resource_batch_id += 1
        self.solution = sol


class KnapsackRandomHeuristic(_KnapsackBase):
    """knapsack_v2/heuristic/random_heuristic.py:13-52 (Philox stream per instance, see RandomHeuristic)."""

    def __init__(self, num_bins: int, opts: dict, device="cuda:0", seed: int = 1235, instance_id0: int = 0):
        super().__init__(num_bins, device)
        self.seed, self.instance_id0 = int(seed), int(instance_id0)
        self.name = "random"

    def solve(self, state):
        st = self._state(state)
        lib, h = _handle(self.device)
        B, R, sol = self._alloc(st)
        L.check(lib.tpc_knapsack_random(h, L.ptr(st), st.stride(0), st.stride(1), B, self.num_bins, R, self.seed,
                                        self.instance_id0, self.resource_batch_id, L.ptr(sol.assign), L.ptr(sol.order),
                                        L.ptr(sol.bins), L.ptr(sol.values), _stream(self.device)), "tpc_knapsack_random")
        self.resource_batch_id += 1
        self.solution = sol


def knapsack_heuristic_factory(num_bins: int, opts: dict, device="cuda:0", seed: int = 1235, instance_id0: int = 0):
    """knapsack_v2/heuristic/factory.py:5-50: waste-reduction combos, random. The reference also appends an ORTools
    solver (heuristic/or_tools.py); Google OR-Tools is not installable here, so it is left out."""
    wr = opts["waste_reduction"]
    if wr["generate_params_combos"]:
        solvers = [WasteReductionHeuristic(num_bins, {"item_sort_descending": i, "bin_sort_descending": b}, device)
                   for i in (True, False) for b in (True, False)]
    else:
        solvers = [WasteReductionHeuristic(num_bins, wr, device)]
    return solvers + [KnapsackRandomHeuristic(num_bins, opts["random"], device, seed, instance_id0)]


def knapsack_stats(assign, stride_instance, stride_item, values):
    device = values.device
    lib, h = _handle(device)
    B, N = values.shape
    R = assign.numel() // B
    reward = torch.empty(B, dtype=torch.float32, device=device)
    empty = torch.empty(B, dtype=torch.int32, device=device)
    rejected = torch.empty(B, dtype=torch.int32, device=device)
    rej_value = torch.empty(B, dtype=torch.float32, device=device)
    L.check(lib.tpc_knapsack_stats(h, L.ptr(assign), stride_instance, stride_item, L.ptr(values), B, N, R,
                                   L.ptr(reward), L.ptr(empty), L.ptr(rejected), L.ptr(rej_value), _stream(device)),
            "tpc_knapsack_stats")
    return reward, empty, rejected, rej_value


def knapsack_net_stats(bufs, initial_state: torch.Tensor, num_bins: int):
    """compute_stats of the net's placement: per-bin values summed in decoding order (what env.history holds in the
    reference), then the same statistics as the heuristics."""
    device = initial_state.device
    lib, h = _handle(device)
    B, R = bufs.actions.shape[1], bufs.actions.shape[0]
    st = initial_state.contiguous()
    values = torch.empty(B, num_bins, dtype=torch.float32, device=device)
    L.check(lib.tpc_knapsack_bin_values(h, L.ptr(bufs.actions), 1, bufs.actions.stride(0), L.ptr(st), st.stride(0),
                                        st.stride(1), B, num_bins, R, L.ptr(values), _stream(device)),
            "tpc_knapsack_bin_values")
    return knapsack_stats(bufs.actions, 1, bufs.actions.stride(0), values), values
