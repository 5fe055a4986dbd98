"""Size-independent properties of the heuristic oracle on random instances (the same invariants hold for the device
kernels, which are compared with this oracle bit for bit in tests/test_gpu_heuristics.py)."""
import numpy as np
import pytest

from oracle import heuristics as oh


def _instance(rng, n_nodes, n_rules, lo, hi):
    s = np.zeros((n_nodes + n_rules, 3), dtype=np.float32)
    s[0] = -2
    s[1:n_nodes] = (rng.integers(lo, hi, size=(n_nodes - 1, 3)) / 100).astype(np.float32)
    s[n_nodes:] = (rng.integers(1, 30, size=(n_rules, 3)) / 100).astype(np.float32)
    return s


@pytest.mark.parametrize("seed", range(6))
def test_placements_conserve_capacity_and_never_overload(seed):
    rng = np.random.default_rng(seed)
    n_nodes, n_rules = int(rng.integers(2, 20)), int(rng.integers(1, 40))
    st = _instance(rng, n_nodes, n_rules, 0, 60)
    k0 = np.round(st.astype(np.float64) * 100).astype(np.int64)
    sols = [oh.dominant_heuristic(st, n_nodes, rd, nd) for rd in (True, False) for nd in (True, False)]
    sols.append(oh.random_heuristic(st, n_nodes, oh.philox_randrange(seed, 0, 0)))
    for assign, rem in sols:
        k1 = np.round(rem.astype(np.float64) * 100).astype(np.int64)
        assert assign.shape == (n_rules,) and assign.min() >= 0 and assign.max() < n_nodes
        assert (k1[1:] >= 0).all()                                          # no node is ever overloaded
        placed = np.zeros((n_nodes, 3), dtype=np.int64)
        np.add.at(placed, assign, k0[n_nodes:])
        np.testing.assert_array_equal((k0[:n_nodes] - k1)[1:], placed[1:])  # capacity removed == demands placed
        dom, rej, empty = oh.solution_stats(assign, rem)
        assert rej == int((assign == 0).sum()) and 0 <= empty <= n_nodes - 1
        assert float(dom) == float(rem[1:].min())
        # a rejected rule fits no node in the FINAL state either (capacities only shrink)
        for j in np.nonzero(assign == 0)[0]:
            assert (oh.dominant_diff(rem[1:], st[n_nodes + j]) < 0).all()


@pytest.mark.parametrize("seed", range(4))
def test_knapsack_placements_respect_capacity_and_account_for_every_value(seed):
    rng = np.random.default_rng(100 + seed)
    n_bins, n_items = int(rng.integers(2, 15)), int(rng.integers(1, 40))
    st = np.zeros((n_bins + n_items, 2), dtype=np.float32)
    st[0] = -2
    st[1:n_bins, 0] = (rng.integers(5, 80, size=n_bins - 1) / 100).astype(np.float32)
    st[n_bins:, 0] = (rng.integers(1, 30, size=n_items) / 100).astype(np.float32)
    st[n_bins:, 1] = (rng.integers(1, 100, size=n_items) / 100).astype(np.float32)
    sols = [oh.knapsack_waste_reduction(st, n_bins, i, b) for i in (True, False) for b in (True, False)]
    sols.append(oh.knapsack_random(st, n_bins, oh.philox_randrange(seed, 1, 0)))
    for assign, order, bins, values in sols:
        assert sorted(order.tolist()) == list(range(n_items))
        assert (bins[1:, 1] <= bins[1:, 0] + 1e-6).all()                    # load never exceeds capacity
        load = np.zeros(n_bins)
        np.add.at(load, assign, np.round(st[n_bins:, 0].astype(np.float64) * 100))
        np.testing.assert_array_equal(np.round(bins[1:, 1].astype(np.float64) * 100), load[1:])
        reward, empty, rejected, rej_value = oh.knapsack_solution_stats(assign, values)
        total = st[n_bins:, 1].astype(np.float64).sum()
        assert abs(float(reward) + float(rej_value) - total) < 1e-3        # every item's value lands somewhere
        assert rejected == int((assign == 0).sum())
