"""CPU restatement (test infrastructure only) of the reference's placement heuristics for ResourceV3:

  * DominantResourceHeuristic  environment/custom/resource_v3/heuristic/dominant_heuristic.py:13-88
  * RandomHeuristic            environment/custom/resource_v3/heuristic/random_heuristic.py:14-65
  * Node bookkeeping           environment/custom/resource_v3/node.py:36-78 (fp32 round_half_up arithmetic)
  * compute_stats              environment/custom/resource_v3/misc/utils.py:56-81

Pinned by tests/golden/heuristic_*.npz, produced by running the UNMODIFIED reference classes
(tests/golden/make_heuristic_golden.py). The random heuristic draws from Python's `random` in the reference; here the
draws come from a caller-supplied function so that the device kernel's Philox stream can be replayed exactly.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def rhu2(x):
    """round_half_up(x, 2) in fp32 (misc/utils.py:19-22)."""
    x = np.asarray(x, dtype=F32)
    return (np.floor(x * F32(100) + F32(0.5)) / F32(100)).astype(F32)


def dominant_diff(rem, req):
    """node.compute_dominant_resource (node.py:36-41): min over features of rhu(remaining - request)."""
    return rhu2(rem - req).min(axis=-1)


def dominant_heuristic(state, num_nodes, resource_sort_descending, node_sort_descending):
    """state (S,3) fp32 of ONE instance (row 0 = EOS node). Returns (assign (R,) int32 node id per rule in the
    original rule order, 0 = rejected; remaining (num_nodes,3) fp32)."""
    state = np.asarray(state, dtype=F32)
    rem = state[:num_nodes].copy()
    rules = state[num_nodes:]
    keys = rules.max(axis=1)
    # Python's sorted() is stable, and with reverse=True equal elements keep their original order as well
    order = sorted(range(len(rules)), key=lambda j: keys[j], reverse=bool(resource_sort_descending))
    assign = np.zeros(len(rules), dtype=np.int32)
    for j in order:
        diffs = dominant_diff(rem[1:], rules[j])
        cand = sorted(range(len(diffs)), key=lambda n: diffs[n], reverse=bool(node_sort_descending))
        for n in cand:
            if diffs[n] >= 0:
                rem[1 + n] = rhu2(rem[1 + n] - rules[j])
                assign[j] = 1 + n
                break
    return assign, rem


def random_heuristic(state, num_nodes, randrange):
    """`randrange(n)` returns the next uniform integer in [0, n) (random_heuristic.py:33-63 call order: one draw per
    rule pick, then one per node tried)."""
    state = np.asarray(state, dtype=F32)
    rem = state[:num_nodes].copy()
    rules = state[num_nodes:]
    left = list(range(len(rules)))
    assign = np.zeros(len(rules), dtype=np.int32)
    while left:
        j = left.pop(randrange(len(left)))
        nodes = list(range(1, num_nodes))
        while nodes:
            n = nodes.pop(randrange(len(nodes)))
            if dominant_diff(rem[n], rules[j]) >= 0:
                rem[n] = rhu2(rem[n] - rules[j])
                assign[j] = n
                break
    return assign, rem


def solution_stats(assign, rem):
    """compute_stats (misc/utils.py:56-81) on a finished placement: (dominant, rejected, empty nodes)."""
    num_nodes = rem.shape[0]
    rejected = int((assign == 0).sum())
    dominant = F32(rem[1:].min())
    used = np.zeros(num_nodes, dtype=bool)
    used[assign] = True
    empty = int((~used[1:]).sum())
    return dominant, rejected, empty
