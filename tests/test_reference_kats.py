"""The reference's own known-answer tests for the hot path, restated against the NumPy oracle.
Sources: tests/unit/environment/resource_v3/env_test.py:75-96,173-262,316-499;
reward_test.py:9-328; utils_test.py:10-77 (values lifted verbatim; data, not code)."""
import numpy as np
import pytest

from oracle import env as oe

EOS = np.full((1, 3), -2, dtype="float32")
BIN_MASK = np.array([[0., 0., 0., 0., 1., 1.], [0., 0., 0., 0., 1., 1.]], dtype="float32")


def _cfg(reward="greedy", mask_nodes=False):
    return {"mask_nodes_in_mha": mask_nodes, "decimal_precision": 2, "batch_size": 2, "num_features": 3,
            "profiles_sample_size": 2, "node_sample_size": 3, "EOS_CODE": -2,
            "reward": {"type": reward, "greedy": {}, "single_node_dominant": {"rejection_penalty": -2},
                       "global_dominant": {"rejection_penalty": -2},
                       "reduced_node_usage": {"rejection_penalty": -2, "use_new_node_penalty": -1}}}


def _state(rows0, rows1):
    return np.array([[[-2., -2., -2.]] + rows0, [[-2., -2., -2.]] + rows1], dtype="float32")


FEASIBLE_CASES = [
    # env_test.py:316-352 all unmasked
    (_state([[.8, .9, .8], [.8, .8, .8], [.9, .8, .9], [.1, .1, .1], [.3, .3, .3]],
            [[.8, .9, .9], [.9, .9, .8], [.8, .9, .8], [.2, .2, .2], [.5, .5, .5]]),
     [[.1, .1, .1], [.2, .2, .2]], [[0, 0, 0, 0, 1, 1], [0, 0, 0, 0, 1, 1]]),
    # env_test.py:354-390 all masked
    (_state([[0, 0, 0], [.1, .1, .1], [.2, .2, .2], [.1, .1, .1], [.3, .3, .3]],
            [[.2, .2, .2], [.3, .3, .3], [.4, .4, .4], [.2, .2, .2], [.5, .5, .5]]),
     [[.3, .3, .3], [.5, .5, .5]], [[0, 1, 1, 1, 1, 1], [0, 1, 1, 1, 1, 1]]),
    # env_test.py:392-428 2nd and 3rd unmasked
    (_state([[0, 0, 0], [.4, .4, .4], [.3, .3, .3], [.1, .1, .1], [.3, .3, .3]],
            [[.2, .2, .2], [.5, .5, .5], [.6, .7, .8], [.2, .2, .2], [.5, .5, .5]]),
     [[.3, .3, .3], [.5, .5, .5]], [[0, 1, 0, 0, 1, 1], [0, 1, 0, 0, 1, 1]]),
    # env_test.py:430-466 mask 2 nodes
    (_state([[0, 0, 0], [.4, .4, .4], [.03, .03, .03], [.1, .1, .1], [.3, .3, .3]],
            [[.7, .7, .7], [.3, .3, .3], [.1, .1, .1], [.2, .2, .2], [.5, .5, .5]]),
     [[.3, .3, .3], [.5, .5, .5]], [[0, 1, 0, 1, 1, 1], [0, 0, 1, 1, 1, 1]]),
    # env_test.py:468-499 single resource overloaded
    (_state([[.9, 0, .8], [.7, .8, 0], [.3, .3, .3], [.1, .1, .1], [.3, .3, .3]],
            [[.7, .7, .7], [.3, 0, .3], [0, .7, .7], [.2, .2, .2], [.5, .5, .5]]),
     [[.3, .3, .3], [.5, .5, .5]], [[0, 1, 1, 0, 1, 1], [0, 0, 1, 1, 1, 1]]),
]


@pytest.mark.parametrize("state,req,expected", FEASIBLE_CASES)
def test_build_feasible_mask_kats(state, req, expected):
    env = oe.ResourceV3Oracle(_cfg())
    env.load(state)
    got = env.build_feasible_mask(state, np.expand_dims(np.array(req, dtype="float32"), 1), BIN_MASK.copy())
    assert got.tolist() == np.array(expected, dtype="float32").tolist()


def test_initial_masks_and_step_kat():
    """env_test.py:75-96 (initial masks) and :173-262 (step)."""
    state = np.array([[[-2, -2, -2], [.84, .9, .85], [.89, .87, .89], [.94, .86, .95], [.29, .17, .13], [.19, .29, .22]],
                      [[-2, -2, -2], [.89, .95, .97], [.99, .95, .87], [.89, .9, .82], [.26, .23, .22], [.1, .19, .25]]],
                     dtype="float32")
    env = oe.ResourceV3Oracle(_cfg())
    s, dec, bin_mask, mha = env.load(state)
    assert bin_mask.tolist() == BIN_MASK.tolist()
    assert mha.tolist() == np.zeros((2, 1, 1, 6), dtype="float32").tolist()
    assert dec.shape == (2, 1, 3) and s.shape == (2, 6, 3)
    ids = [0, 2]
    expected = state.copy()
    # NB the reference test aliases `fake_state` with `env.batch` (env_test.py:216, 228-231), so its state assertion
    # compares an array with itself; the value the env really produces is the ROUNDED remainder (env.py:133-137).
    expected[[0, 1], ids] = oe.round_half_up(state[[0, 1], ids] - state[[0, 1], [4, 4]], 2)
    expected[[0, 1], 0] = EOS
    nxt, ndec, rewards, done, info = env.step(ids, BIN_MASK.copy())
    assert nxt.tolist() == expected.tolist()
    res = np.array([[1, 1, 1, 1, 1, 0], [1, 1, 1, 1, 1, 0]], dtype="float32")
    assert info["resource_net_mask"].tolist() == res.tolist()
    assert info["bin_net_mask"].tolist() == BIN_MASK.tolist()
    m = np.zeros((2, 1, 1, 6), dtype="float32")
    m[:, :, :, 4] = 1
    assert info["mha_used_mask"].tolist() == m.tolist()
    assert rewards.shape == (2, 1) and not done
    assert ndec.tolist() == state[:, 5:6].tolist()          # env_test.py:597-692 next decoder input
    _, _, _, done, _ = env.step([1, 1], BIN_MASK.copy())
    assert done                                              # env_test.py:303-314


def _reward_inputs(ids, reqs):
    og = np.array([[[-2., -2., -2.], [100., 200., 300.], [400., 500., 600.], [10., 20., 30.], [40., 50., 60.]]] * len(ids),
                  dtype="float32")
    idx = np.arange(len(ids))
    nodes, rq = og[idx, ids], og[idx, reqs]
    upd = og.copy()
    upd[idx, ids] = nodes - rq
    return og, upd, nodes, rq


def test_reward_kats():
    """reward_test.py:9-70 (greedy), :72-135 (single node dominant), :137-200 (global dominant)."""
    og, upd, nodes, rq = _reward_inputs([0, 2], [3, 4])
    assert oe.GreedyReward({}, EOS).compute_reward(upd, og, 3, nodes, rq, None, [0, 2]).tolist() == [0, 1]
    r = oe.SingleNodeDominantReward({"rejection_penalty": -1}, EOS).compute_reward(upd, og, 3, nodes, rq, None, [0, 2])
    assert r.tolist() == [-1, 360]
    r = oe.GlobalDominantReward({"rejection_penalty": -1}, EOS).compute_reward(upd, og, 3, nodes, rq, None, [0, 2])
    assert r.tolist() == [-1, 100]


def test_reduced_node_usage_kat_with_is_empty_mutation():
    """reward_test.py:202-328."""
    og, upd, nodes, rq = _reward_inputs([0, 2, 1], [3, 4, 4])
    rw = oe.ReducedNodeUsage({"rejection_penalty": -10, "use_new_node_penalty": -1}, EOS)
    is_empty = np.zeros((3, 5, 1), dtype="float32")
    r = rw.compute_reward(upd, og, 3, nodes, rq, None, [0, 2, 1], is_empty)
    assert is_empty[:, :, 0].tolist() == [[-2, 0, 0, 0, 0], [-2, 0, 1, 0, 0], [-2, 1, 0, 0, 0]]
    assert r.tolist() == [-10, -1, -1]
    r = rw.compute_reward(upd, og, 3, nodes, rq, None, [0, 2, 2], is_empty)
    assert is_empty[:, :, 0].tolist() == [[-2, 0, 0, 0, 0], [-2, 0, 1, 0, 0], [-2, 1, 1, 0, 0]]
    assert r.tolist()[1] == 0    # node already used: no new-node penalty


def test_round_half_up_and_reshapes():
    """utils_test.py:10-77."""
    assert oe.round_half_up(0.4099999964237213, 2) == 0.41
    assert oe.round_half_up(0.34999999962747097, 2) == 0.35
    assert oe.round_half_up(0.05999999865889549, 2) == 0.06
    x = np.array([[1, 2, 3], [4, 5, 6]], dtype="float32")
    v = oe.reshape_into_vertical_format(x, 6)
    assert v.tolist() == [[1], [4], [2], [5], [3], [6]]
    assert oe.reshape_into_horizontal_format(v, 2, 3).tolist() == x.tolist()


def test_four_feature_state_for_cost_reward():
    """env_test.py:794-820: reduced_node_usage appends is_empty as a 4th feature; EOS row = -2."""
    env = oe.ResourceV3Oracle(_cfg("reduced_node_usage"))
    state = _state([[.5, .5, .5], [.5, .5, .5], [.5, .5, .5], [.1, .1, .1], [.1, .1, .1]],
                   [[.5, .5, .5], [.5, .5, .5], [.5, .5, .5], [.1, .1, .1], [.1, .1, .1]])
    s, dec, _, _ = env.load(state)
    assert s.shape == (2, 6, 4) and dec.shape == (2, 1, 3)
    assert np.all(s[:, 0, 3] == -2) and np.all(s[:, 1:, 3] == 0)


def test_integer_hundredths_equivalence():
    """on the value domain the fp32 round_half_up arithmetic equals exact integer arithmetic."""
    ka, kb = np.meshgrid(np.arange(-200, 101), np.arange(0, 101), indexing="ij")
    a = (ka / 100).astype(np.float32)
    b = (kb / 100).astype(np.float32)
    got = oe.round_half_up(a - b, 2)
    want = ((ka - kb) / 100).astype(np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
