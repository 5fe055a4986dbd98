"""Golden placements of the UNMODIFIED reference KnapsackV2 heuristics
(environment/custom/knapsack_v2/heuristic/{waste_reduction,random_heuristic}.py).

Run in the build container only (needs /root/reference):

    python tests/golden/make_knapsack_heuristic_golden.py

WasteReductionHeuristic is deterministic: its four (item order, bin order) variants are recorded on random
KnapsackV2-shaped instances. RandomHeuristic draws from Python's `random`: the draw sequence is recorded with the
placement so the oracle can replay it. ORTools (heuristic/or_tools.py) needs Google OR-Tools, which is not installed.
Output: tests/golden/heuristic_knapsack_v2.npz
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
sys.path.insert(0, "/root/reference")

from environment.custom.knapsack_v2.heuristic.waste_reduction import WasteReductionHeuristic  # noqa: E402
from environment.custom.knapsack_v2.heuristic import random_heuristic as rh  # noqa: E402
from environment.custom.knapsack_v2.misc.utils import compute_stats  # noqa: E402


def instance(rng, n_bins, n_items, cap_lo, cap_hi):
    s = np.zeros((n_bins + n_items, 2), dtype=np.float32)
    s[0] = -2
    s[1:n_bins, 0] = (rng.integers(cap_lo, cap_hi, size=n_bins - 1) / 100).astype(np.float32)   # capacity, load 0
    s[n_bins:, 0] = (rng.integers(1, 30, size=n_items) / 100).astype(np.float32)                # weight
    s[n_bins:, 1] = (rng.integers(1, 100, size=n_items) / 100).astype(np.float32)               # value
    return s


def solution_arrays(solver, state, n_bins):
    n_items = state.shape[0] - n_bins
    assign = np.full(n_items, -1, dtype=np.int32)
    bins = np.zeros((n_bins, 2), dtype=np.float32)
    values = np.zeros(n_bins, dtype=np.float32)
    for b in solver.solution:
        bins[b.id] = [b.capacity[0], b.current_load[0]]
        values[b.id] = b.current_value[0]
        for it in b.item_list:
            assign[it.id] = b.id
    assert (assign >= 0).all()
    reward, empty, rejected, rejected_value = compute_stats(solver.solution)
    return assign, bins, values, np.array([float(np.float32(reward)[0]) if np.ndim(reward) else float(reward), empty,
                                           rejected, float(rejected_value[0])], dtype=np.float64)


def main():
    rng = np.random.default_rng(20260202)
    out = {}
    cases = [(3, 5, 10, 100), (6, 10, 10, 100), (11, 20, 10, 100), (11, 20, 10, 40), (21, 60, 30, 80), (51, 100, 10, 100),
             (51, 100, 5, 30), (2, 7, 10, 50)]
    idx = 0
    for (n_bins, n_items, lo, hi) in cases:
        for rep in range(4):
            st = instance(rng, n_bins, n_items, lo, hi)
            out[f"c{idx}_state"] = st
            out[f"c{idx}_bins"] = np.array(n_bins)
            for i_desc in (True, False):
                for b_desc in (True, False):
                    solver = WasteReductionHeuristic(n_bins, {"item_sort_descending": i_desc, "bin_sort_descending": b_desc})
                    solver.solve(st[None])
                    a, b, v, s = solution_arrays(solver, st, n_bins)
                    tag = f"c{idx}_wr_{int(i_desc)}{int(b_desc)}"
                    out[tag + "_assign"], out[tag + "_bins"], out[tag + "_values"], out[tag + "_stats"] = a, b, v, s
            draws = []
            real = random.randrange

            def rec(n, _real=real, _d=draws):
                v = _real(n)
                _d.append((n, v))
                return v
            random.seed(2000 + idx)
            rh.random.randrange = rec
            try:
                solver = rh.RandomHeuristic(n_bins, {})
                solver.solve(st[None])
            finally:
                rh.random.randrange = real
            a, b, v, s = solution_arrays(solver, st, n_bins)
            out[f"c{idx}_rnd_assign"], out[f"c{idx}_rnd_bins"], out[f"c{idx}_rnd_values"] = a, b, v
            out[f"c{idx}_rnd_stats"] = s
            out[f"c{idx}_rnd_draws"] = np.array(draws, dtype=np.int32)
            idx += 1
    out["num_cases"] = np.array(idx)
    np.savez_compressed(os.path.join(HERE, "heuristic_knapsack_v2.npz"), **out)
    print("cases", idx)


if __name__ == "__main__":
    main()
