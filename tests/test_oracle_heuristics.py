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


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32-10."""
    assert oh.philox4x32_10((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oh.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oh.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    rr = oh.philox_randrange(1235, 7, 0)
    draws = [rr(n) for n in (1, 2, 3, 50, 100, 7, 7, 7, 7)]
    assert draws[0] == 0 and all(0 <= d < n for d, n in zip(draws, (1, 2, 3, 50, 100, 7, 7, 7, 7)))


def test_compare_solutions_follows_reference_thresholds():
    """misc/utils.py:116-157: heuristics start from max_dominant = -1 / min_rejected = 9999."""
    nd, cd, cr = oh.compare_solutions(np.float32(0.104999), 3, [np.float32(0.10), np.float32(0.05)], [4, 3])
    assert float(nd) == float(np.float32(0.10)) and (cd, cr) == (1, 1)
    assert oh.compare_solutions(np.float32(0.2), 1, [np.float32(0.3)], [0])[1:] == (2, 2)
    assert oh.compare_solutions(np.float32(0.2), 1, [np.float32(0.1)], [2])[1:] == (0, 0)
    assert oh.compare_solutions(np.float32(0.0), 0, [], [])[1:] == (0, 0)     # no heuristic: -1 < 0 and 9999 > 0
