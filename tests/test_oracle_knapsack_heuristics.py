"""The KnapsackV2 heuristic oracle (oracle/heuristics.py) against placements of the unmodified reference classes
(tests/golden/heuristic_knapsack_v2.npz, generator: tests/golden/make_knapsack_heuristic_golden.py)."""
import os

import numpy as np
import pytest

from oracle import heuristics as oh

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "heuristic_knapsack_v2.npz"))
N_CASES = int(G["num_cases"])


def _check(tag, a, order, bins, values, st, n_bins):
    np.testing.assert_array_equal(a, G[tag + "_assign"])
    np.testing.assert_array_equal(bins[1:], G[tag + "_bins"][1:])           # bit-exact fp32 load bookkeeping
    np.testing.assert_array_equal(values, G[tag + "_values"])               # per-bin value sums in placement order
    np.testing.assert_array_equal(values, oh.knapsack_bin_values(a, st, n_bins, order))
    reward, empty, rejected, rej_value = oh.knapsack_solution_stats(a, values)
    ref = G[tag + "_stats"]
    assert (float(reward), empty, rejected, float(rej_value)) == (float(np.float32(ref[0])), int(ref[1]), int(ref[2]),
                                                                  float(np.float32(ref[3])))


@pytest.mark.parametrize("idx", range(N_CASES))
def test_waste_reduction_matches_reference(idx):
    st, n_bins = G[f"c{idx}_state"], int(G[f"c{idx}_bins"])
    for i_desc in (1, 0):
        for b_desc in (1, 0):
            a, order, bins, values = oh.knapsack_waste_reduction(st, n_bins, i_desc, b_desc)
            _check(f"c{idx}_wr_{i_desc}{b_desc}", a, order, bins, values, st, n_bins)


@pytest.mark.parametrize("idx", range(N_CASES))
def test_random_replays_reference_draws(idx):
    st, n_bins = G[f"c{idx}_state"], int(G[f"c{idx}_bins"])
    it = iter([tuple(x) for x in G[f"c{idx}_rnd_draws"]])

    def randrange(n):
        m, v = next(it)
        assert m == n, "draw sequence out of step with the reference"
        return int(v)
    a, order, bins, values = oh.knapsack_random(st, n_bins, randrange)
    _check(f"c{idx}_rnd", a, order, bins, values, st, n_bins)
    assert next(it, None) is None


KN_KAT = np.array([[-2.0, -2.0], [0.1, 0.0], [0.5, 0.0], [0.2, 0.1], [0.3, 0.5], [0.1, 0.4], [0.9, 0.4]], dtype=np.float32)


def test_reference_unit_test_instance():
    """tests/unit/environment/knapsack_v2/heuristic_test.py:236-300 (WasteReduction, items descending / bins
    ascending): reward 1.0, no empty bin, one rejected item worth 0.4, item counts per bin 1 / 1 / 2."""
    a, order, bins, values = oh.knapsack_waste_reduction(KN_KAT, 3, True, False)
    reward, empty, rejected, rej_value = oh.knapsack_solution_stats(a, values)
    assert round(float(reward), 2) == 1.0 and (empty, rejected) == (0, 1) and rej_value == np.float32(0.4)
    assert [int((a == n).sum()) for n in range(3)] == [1, 1, 2]
