"""The heuristic oracle (oracle/heuristics.py) against placements produced by the unmodified reference classes
(tests/golden/heuristic_resource_v3.npz, generator: tests/golden/make_heuristic_golden.py)."""
import os

import numpy as np
import pytest

from oracle import heuristics as oh

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "heuristic_resource_v3.npz"))
N_CASES = int(G["num_cases"])


@pytest.mark.parametrize("idx", range(N_CASES))
def test_dominant_matches_reference(idx):
    st, n_nodes = G[f"c{idx}_state"], int(G[f"c{idx}_nodes"])
    for rd in (1, 0):
        for nd in (1, 0):
            a, r = oh.dominant_heuristic(st, n_nodes, rd, nd)
            np.testing.assert_array_equal(a, G[f"c{idx}_dom_{rd}{nd}_assign"])
            np.testing.assert_array_equal(r[1:], G[f"c{idx}_dom_{rd}{nd}_rem"][1:])   # bit-exact fp32 bookkeeping
            dom, rej, empty = oh.solution_stats(a, r)
            ref = G[f"c{idx}_dom_{rd}{nd}_stats"]
            assert (float(dom), rej, empty) == (float(np.float32(ref[0])), int(ref[1]), int(ref[2]))


@pytest.mark.parametrize("idx", range(N_CASES))
def test_random_replays_reference_draws(idx):
    st, n_nodes = G[f"c{idx}_state"], int(G[f"c{idx}_nodes"])
    draws = [tuple(x) for x in G[f"c{idx}_rnd_draws"]]
    it = iter(draws)

    def randrange(n):
        m, v = next(it)
        assert m == n, "draw sequence out of step with the reference"
        return int(v)
    a, r = oh.random_heuristic(st, n_nodes, randrange)
    np.testing.assert_array_equal(a, G[f"c{idx}_rnd_assign"])
    np.testing.assert_array_equal(r[1:], G[f"c{idx}_rnd_rem"][1:])
    assert next(it, None) is None


def test_reference_unit_test_instance():
    """tests/unit/environment/resource_v3/heuristic_test.py:166-199: [3,5,8] is rejected, [2,1,4] goes to node 2."""
    a, r = oh.dominant_heuristic(G["kat_state"], 3, True, True)
    np.testing.assert_array_equal(a, G["kat_assign"])
    np.testing.assert_array_equal(a, [2, 0])
    np.testing.assert_array_equal(r[1:], G["kat_rem"][1:])
