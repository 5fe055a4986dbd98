"""Golden placements of the UNMODIFIED reference heuristics (environment/custom/resource_v3/heuristic/*).

Run in the build container only (needs /root/reference):

    PYTHONPATH=oracle/tf_shim:/root/reference python tests/golden/make_heuristic_golden.py

DominantResourceHeuristic is deterministic: all four (resource order, node order) variants are recorded on random
ResourceV3-shaped instances (values k/100 like the environment's), plus the reference's own unit-test instance
(tests/unit/environment/resource_v3/heuristic_test.py:26-34). RandomHeuristic draws from Python's `random`: the
draw sequence is recorded with the placements so the oracle can replay it.
Output: tests/golden/heuristic_resource_v3.npz
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
sys.path.insert(0, "/root/reference")

from environment.custom.resource_v3.heuristic.dominant_heuristic import DominantResourceHeuristic  # noqa: E402
from environment.custom.resource_v3.heuristic import random_heuristic as rh  # noqa: E402
from environment.custom.resource_v3.misc.utils import compute_stats  # noqa: E402


def instance(rng, n_nodes, n_rules, node_lo, node_hi):
    s = np.zeros((n_nodes + n_rules, 3), dtype=np.float32)
    s[0] = -2
    s[1:n_nodes] = (rng.integers(node_lo, node_hi, size=(n_nodes - 1, 3)) / 100).astype(np.float32)
    s[n_nodes:] = (rng.integers(1, 30, size=(n_rules, 3)) / 100).astype(np.float32)
    return s


def solution_arrays(solver, state, n_nodes):
    n_rules = state.shape[0] - n_nodes
    assign = np.full(n_rules, -1, dtype=np.int32)
    rem = np.zeros((n_nodes, 3), dtype=np.float32)
    for node in solver.solution:
        rem[node.id] = [node.remaining_CPU[0], node.remaining_RAM[0], node.remaining_MEM[0]]
        for req in node.req_list:
            assign[req.id] = node.id
    assert (assign >= 0).all()
    dom, rej, empty = compute_stats(solver.solution)
    return assign, rem, np.array([float(dom[0]), rej, empty], dtype=np.float64)


def main():
    rng = np.random.default_rng(20260101)
    out = {}
    cases = [(4, 6, 0, 100), (6, 10, 0, 100), (11, 20, 0, 100), (11, 20, 0, 30), (21, 40, 20, 60), (51, 100, 0, 100),
             (51, 100, 0, 40), (8, 30, 0, 20), (3, 5, 0, 100)]
    idx = 0
    for (n_nodes, n_rules, lo, hi) in cases:
        for rep in range(4):
            st = instance(rng, n_nodes, n_rules, lo, hi)
            out[f"c{idx}_state"] = st
            out[f"c{idx}_nodes"] = np.array(n_nodes)
            for rd in (True, False):
                for nd in (True, False):
                    solver = DominantResourceHeuristic(n_nodes, 100, {"resource_sort_descending": rd, "node_sort_descending": nd})
                    solver.solve(st[None])
                    a, r, s = solution_arrays(solver, st, n_nodes)
                    out[f"c{idx}_dom_{int(rd)}{int(nd)}_assign"] = a
                    out[f"c{idx}_dom_{int(rd)}{int(nd)}_rem"] = r
                    out[f"c{idx}_dom_{int(rd)}{int(nd)}_stats"] = s
            # random heuristic with its draw sequence recorded
            draws = []
            real = random.randrange

            def rec(n, _real=real, _d=draws):
                v = _real(n)
                _d.append((n, v))
                return v
            random.seed(1000 + idx)
            rh.random.randrange = rec
            try:
                solver = rh.RandomHeuristic(n_nodes, 100, {})
                solver.solve(st[None])
            finally:
                rh.random.randrange = real
            a, r, s = solution_arrays(solver, st, n_nodes)
            out[f"c{idx}_rnd_assign"] = a
            out[f"c{idx}_rnd_rem"] = r
            out[f"c{idx}_rnd_stats"] = s
            out[f"c{idx}_rnd_draws"] = np.array(draws, dtype=np.int32)
            idx += 1
    # the reference's own unit-test instance (un-normalised integers)
    st = np.array([[-2, -2, -2], [1, 2, 3], [5, 2, 6], [2, 1, 4], [3, 5, 8]], dtype=np.float32)
    solver = DominantResourceHeuristic(3, 100, {"resource_sort_descending": True, "node_sort_descending": True})
    solver.solve(st[None])
    a, r, s = solution_arrays(solver, st, 3)
    out["kat_state"], out["kat_assign"], out["kat_rem"] = st, a, r
    out["num_cases"] = np.array(idx)
    np.savez_compressed(os.path.join(HERE, "heuristic_resource_v3.npz"), **out)
    print("cases", idx, "kat assign", a)


if __name__ == "__main__":
    main()
