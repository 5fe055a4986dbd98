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


# ---- Philox4x32-10 (Salmon et al. 2011): replays the device kernel's random.randrange substitute -----------------
_M0, _M1, _W0, _W1, _MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF


def philox4x32_10(counter, key):
    c = [int(x) & _MASK for x in counter]
    k0, k1 = int(key[0]) & _MASK, int(key[1]) & _MASK
    for _ in range(10):
        p0, p1 = _M0 * c[0], _M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k0, p1 & _MASK, (p0 >> 32) ^ c[3] ^ k1, p0 & _MASK]
        k0, k1 = (k0 + _W0) & _MASK, (k1 + _W1) & _MASK
    return c


def philox_randrange(seed, instance_id, epoch, kind=4):
    """randrange(n) of instance `instance_id`: call i returns floor(u_i * n / 2^32), u_i = word i % 4 of Philox block
    i // 4 with counter (id lo, id hi, epoch, kind << 28 | block) and key (seed lo, seed hi) -- the layout of the
    device streams (csrc/philox.cuh)."""
    state = {"i": 0, "blk": -1, "cur": None}

    def randrange(n):
        i = state["i"]
        state["i"] += 1
        b = i >> 2
        if b != state["blk"]:
            state["cur"] = philox4x32_10((instance_id & _MASK, (instance_id >> 32) & _MASK, epoch & _MASK,
                                          ((kind << 28) | b) & _MASK), (seed & _MASK, (seed >> 32) & _MASK))
            state["blk"] = b
        return (state["cur"][i & 3] * n) >> 32
    return randrange


def compare_solutions(net_dominant, net_rejected, heur_dominant, heur_rejected):
    """gather_stats_from_solutions' verdict (misc/utils.py:107-157) for one instance:
    (rounded net dominant, dominant code, rejected code), codes 0 won / 1 draw / 2 lost."""
    nd = rhu2(net_dominant)
    max_dom, min_rej = F32(-1), 9999
    for d, r in zip(heur_dominant, heur_rejected):
        max_dom = max(max_dom, rhu2(d))
        min_rej = min(min_rej, int(r))
    cd = 2 if max_dom > nd else (1 if max_dom == nd else 0)
    cr = 2 if min_rej < net_rejected else (1 if min_rej == net_rejected else 0)
    return nd, cd, cr


# ---- KnapsackV2 (environment/custom/knapsack_v2/heuristic/*, bin.py, item.py, misc/utils.py:41-58) ---------------
# state (S, 2): bin rows (capacity, current_load), row 0 = EOS bin; item rows (weight, value).
def _kn_remaining(cap, load, w):
    """Bin.compute_remaining_capacity (bin.py:44-47): rhu(capacity - rhu(load + weight))."""
    return rhu2(F32(cap) - rhu2(F32(load) + F32(w)))


def _kn_insert(bins, values, n, item):
    """Bin.insert_item (bin.py:49-60): bin 0 (EOS) only collects the value."""
    if n != 0:
        bins[n, 1] = rhu2(bins[n, 1] + item[0])
    values[n] = F32(values[n] + item[1])


def knapsack_waste_reduction(state, num_bins, item_sort_descending, bin_sort_descending):
    """WasteReductionHeuristic.solve (waste_reduction.py:26-72). Returns (assign (R,), order (R,) = placement order,
    bins (num_bins, 2) after the placement, values (num_bins,) = Bin.current_value)."""
    state = np.asarray(state, dtype=F32)
    bins, items = state[:num_bins].copy(), state[num_bins:]
    values = np.zeros(num_bins, dtype=F32)
    ratio = (items[:, 1] / items[:, 0]).astype(F32)                       # Item.ratio (item.py:14)
    order = sorted(range(len(items)), key=lambda j: ratio[j], reverse=bool(item_sort_descending))
    assign = np.zeros(len(items), dtype=np.int32)
    for j in order:
        diffs = [_kn_remaining(bins[n, 0], bins[n, 1], items[j, 0]) for n in range(1, num_bins)]
        cand = sorted(range(len(diffs)), key=lambda n: diffs[n], reverse=bool(bin_sort_descending))
        pick = 0
        for n in cand:
            if diffs[n] >= 0:
                pick = 1 + n
                break
        _kn_insert(bins, values, pick, items[j])
        assign[j] = pick
    return assign, np.asarray(order, dtype=np.int32), bins, values


def knapsack_random(state, num_bins, randrange):
    """RandomHeuristic.solve (knapsack_v2/heuristic/random_heuristic.py:24-52)."""
    state = np.asarray(state, dtype=F32)
    bins, items = state[:num_bins].copy(), state[num_bins:]
    values = np.zeros(num_bins, dtype=F32)
    left = list(range(len(items)))
    assign = np.zeros(len(items), dtype=np.int32)
    order = []
    while left:
        j = left.pop(randrange(len(left)))
        cand = list(range(1, num_bins))
        pick = 0
        while cand:
            n = cand.pop(randrange(len(cand)))
            if _kn_remaining(bins[n, 0], bins[n, 1], items[j, 0]) >= 0:      # Bin.can_fit_item (bin.py:40-42)
                pick = n
                break
        _kn_insert(bins, values, pick, items[j])
        assign[j] = pick
        order.append(j)
    return assign, np.asarray(order, dtype=np.int32), bins, values


def knapsack_bin_values(assign, state, num_bins, order=None):
    """Bin.current_value of every bin when the items are inserted in `order` (default: item order = the net's
    decoding order): sequential fp32 sums."""
    items = np.asarray(state, dtype=F32)[num_bins:]
    values = np.zeros(num_bins, dtype=F32)
    for j in (range(len(items)) if order is None else order):
        values[assign[j]] = F32(values[assign[j]] + items[j, 1])
    return values


def knapsack_solution_stats(assign, values):
    """compute_stats (knapsack_v2/misc/utils.py:41-58): (reward, empty bins, rejected items, rejected value)."""
    num_bins = len(values)
    reward = F32(0)
    for n in range(1, num_bins):
        reward = F32(reward + values[n])
    used = np.zeros(num_bins, dtype=bool)
    used[assign] = True
    return reward, int((~used[1:]).sum()), int((np.asarray(assign) == 0).sum()), F32(values[0])
