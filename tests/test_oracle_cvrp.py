"""oracle/cvrp.py against trajectories recorded from the UNMODIFIED reference CVRP class
(tests/golden/make_cvrp_golden.py): every tensor crossing the env API, bit for bit."""
import glob
import os

import numpy as np
import pytest

from oracle.cvrp import CvrpOracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "cvrp_*.npz")))


@pytest.mark.parametrize("case", CASES)
def test_cvrp_oracle_replays_reference_trajectory(case):
    z = np.load(os.path.join(GOLDEN, f"{case}.npz"))
    env = CvrpOracle(int(z["num_nodes"]), int(z["num_vehicles"]))
    state, bp, item, mha = env.load(z["state"][0])
    T = z["state"].shape[0]
    assert T == env.steps
    for t in range(T):
        assert np.array_equal(state, z["state"][t]) and np.array_equal(bp, z["backpack_net_mask"][t])
        assert np.array_equal(item, z["item_net_mask"][t]) and np.array_equal(mha, z["mha_mask"][t])
        node = env.request_node(t)
        assert np.all(z["node_ids"][t] == node)
        items = state[:, node]
        assert np.array_equal(items, z["items"][t])
        assert np.array_equal(env.build_feasible_mask(state, items, bp), z["feasible_mask"][t])
        state, reward, done, info = env.step(z["vehicle_ids"][t], z["node_ids"][t])
        assert np.array_equal(reward.reshape(-1), z["reward"][t]) and done == bool(z["done"][t])
        bp, item, mha = info["backpack_net_mask"], info["item_net_mask"], info["mha_used_mask"]
        assert np.array_equal(state, z["next_state"][t]) and np.array_equal(bp, z["next_backpack_net_mask"][t])
        assert np.array_equal(item, z["next_item_net_mask"][t]) and np.array_equal(mha, z["next_mha_mask"][t])
    assert done


def test_cvrp_invariants():
    """Capacity conservation and the done rule on a random instance."""
    rng = np.random.default_rng(0)
    Nn, V, B = 9, 4, 7
    batch = np.zeros((B, Nn + V, 3), np.float32)
    batch[:, :Nn, :2] = rng.integers(0, 100, (B, Nn, 2))
    batch[:, 1:Nn, 2] = rng.integers(1, 25, (B, Nn - 1))
    batch[:, Nn:, :2] = batch[:, :1, :2]
    batch[:, Nn:, 2] = 100
    env = CvrpOracle(Nn, V)
    state, bp, item, mha = env.load(batch)
    total = 0
    for t in range(env.steps):
        node = env.request_node(t)
        feas = env.build_feasible_mask(state, state[:, node], bp)
        v = np.array([rng.choice(np.flatnonzero(feas[b] == 0)) for b in range(B)])
        assert np.all(v >= Nn)
        state, r, done, info = env.step(v, np.full(B, node))
        bp = info["backpack_net_mask"]
        total += r
    assert done and np.all(bp == 1) and np.all(total < 0)
    assert np.array_equal(100 * V - state[:, Nn:, 2].sum(1), batch[:, 1:Nn, 2].sum(1))
    assert np.all(state[:, Nn:, :2] == batch[:, :1, :2])            # every vehicle is back at the depot
