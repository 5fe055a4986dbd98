"""CVRP on the device (BASELINE config #5): the reference class's own step / feasible-mask semantics bit for bit against
trajectories recorded from the UNMODIFIED environment/custom/vrp/env.py (tests/golden/cvrp_*.npz), and the fused
single-pointer rollout replayed through the oracle. Everything goes through libtpc.so."""
import glob
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "cvrp_*.npz")))


def _env(Nn, V, B):
    from tpc_b200.env import CVRP
    return CVRP("CVRP", {"batch_size": B, "node_sample_size": Nn, "vehicle_sample_size": V, "num_features": 3})


@pytest.mark.parametrize("case", CASES)
def test_stepwise_cvrp_matches_reference_goldens(case):
    z = np.load(os.path.join(GOLDEN, f"{case}.npz"))
    Nn, V, B = int(z["num_nodes"]), int(z["num_vehicles"]), z["state"].shape[1]
    env = _env(Nn, V, B)
    state, bp, item, mha = env.load_batch(z["state"][0])
    for t in range(z["state"].shape[0]):
        assert np.array_equal(state, z["state"][t]) and np.array_equal(bp, z["backpack_net_mask"][t])
        assert np.array_equal(item, z["item_net_mask"][t]) and np.array_equal(mha, z["mha_mask"][t])
        node = env.request_node(t)
        feas = env.build_feasible_mask(state, state[:, node], bp)
        assert np.array_equal(feas, z["feasible_mask"][t])
        state, reward, done, info = env.step(z["vehicle_ids"][t], z["node_ids"][t])
        assert np.array_equal(np.asarray(reward).reshape(-1), z["reward"][t]) and done == bool(z["done"][t])
        bp, item, mha = info["backpack_net_mask"], info["item_net_mask"], info["mha_used_mask"]
        assert np.array_equal(state, z["next_state"][t]) and np.array_equal(bp, z["next_backpack_net_mask"][t])
        assert np.array_equal(item, z["next_item_net_mask"][t]) and np.array_equal(mha, z["next_mha_mask"][t])
    assert done


def test_cvrp_step_rejects_overload_and_bad_indices():
    env = _env(6, 2, 3)
    state, bp, item, mha = env.reset()
    with pytest.raises(IndexError):
        env.step([6, 6, 99], [1, 1, 1])
    batch = state.copy()
    batch[:, 6:, 2] = 1.0                              # capacity 1 < every demand
    env.load_batch(batch)
    with pytest.raises(AssertionError):
        env.step([6, 6, 6], [1, 1, 1])


@pytest.mark.parametrize("Nn,V,B,greedy", [(12, 4, 64, True), (32, 5, 256, False)])
def test_fused_cvrp_rollout_replays_through_the_oracle(Nn, V, B, greedy):
    """configs/CVRP.json shape (32 nodes, 5 vehicles): the fused device rollout of the single-pointer adapter, replayed
    step by step through oracle/cvrp.py -- stored states / masks / rewards bit-exact, every pick feasible, logits of the
    actor within 1e-3 of the oracle model."""
    import torch
    from oracle import model as om
    from oracle.cvrp import CvrpOracle
    from tpc_b200.agent import Agent
    from tpc_b200.configs import get_configs
    agent_cfg, _, env_cfg, _, _, _ = get_configs("CVRP", "tpc")
    env_cfg = json.loads(json.dumps(env_cfg))
    # capacity 200: V * (200 - 24) exceeds the largest possible total demand, so no sampled trajectory can reach a node
    # that fits no vehicle (the reference asserts there, vrp/env.py:97-98; a dead end has no defined continuation)
    env_cfg.update(batch_size=B, node_sample_size=Nn, vehicle_sample_size=V, vehicle_capacity=200)
    from tpc_b200.env import env_factory
    env, tester = env_factory("custom", "CVRP", env_cfg)
    assert tester is None
    agent_cfg = env.add_stats_to_agent_config(json.loads(json.dumps(agent_cfg)))
    agent = Agent("transformer", agent_cfg, "cuda:0", num_features=3, seed=4)
    bufs = env.device_buffers(store=True, keep_logits=True)
    env.reset_into(bufs)
    init = bufs.batch.cpu().numpy().copy()
    agent.rollout(env, bufs, greedy=greedy)
    torch.cuda.synchronize()
    T, S = env.steps, Nn + V
    assert bufs.actions.shape[0] == T and bufs.rewards.shape == (B, T)
    ora = CvrpOracle(Nn, V)
    state, bp, item, mha = ora.load(init)
    acts = bufs.actions.cpu().numpy()
    ca = om.NetCfg(num_layers=1, dff=128)
    fa = torch.tensor(agent.bin_actor.get_flat())
    sub = np.arange(0, B, max(1, B // 8))
    worst = 0.0
    for t in range(T):
        node = ora.request_node(t)
        feas = ora.build_feasible_mask(state, state[:, node], bp)
        assert np.array_equal(bufs.states[t].cpu().numpy(), state)
        assert np.array_equal(bufs.ptr_masks[t].cpu().numpy(), feas)
        assert np.array_equal(bufs.mha_masks[t].cpu().numpy(), mha.reshape(B, S))
        assert np.array_equal(bufs.dec_inputs[t].cpu().numpy(), state[:, node])
        a = acts[t]
        assert np.all(feas[np.arange(B), a] == 0) and np.all(a >= Nn)
        ol, _, oi = om.actor_forward(fa, ca, torch.tensor(state[sub]), torch.tensor(state[sub][:, node:node + 1]),
                                     torch.tensor(feas[sub]), torch.tensor(mha[sub]), Nn)
        unm = feas[sub] == 0
        cl = bufs.logits[t].cpu().numpy()[sub]
        worst = max(worst, float(np.abs(cl[unm] - ol.numpy()[unm]).max() / np.abs(ol.numpy()[unm]).max()))
        state, rw, done, info = ora.step(a, np.full(B, node))
        assert np.array_equal(bufs.rewards[:, t].cpu().numpy(), rw.reshape(B))
        bp, mha = info["backpack_net_mask"], info["mha_used_mask"]
    assert done and np.array_equal(bufs.batch.cpu().numpy(), state) and worst < 1e-3
    assert np.all(bp == 1)                                   # every vehicle returned to the depot


_INSTANCE = """NAME : T-n9-k2
COMMENT : (hand made, Min no of trucks: 2, Optimal value: 77)
TYPE : CVRP
DIMENSION : 9
EDGE_WEIGHT_TYPE : EUC_2D
CAPACITY : 20
NODE_COORD_SECTION
 1 40 50
 2 10 20
 3 30 45
 4 70 80
 5 95 10
 6 20 90
 7 55 5
 8 60 60
 9 5 5
DEMAND_SECTION
1 0
2 3
3 3
4 3
5 3
6 3
7 3
8 3
9 3
DEPOT_SECTION
 1
 -1
EOF
"""


def test_cvrp_batches_sampled_from_an_instance_file(tmp_path):
    """vrp/env.py:27-45,148-206: with `dir_path` / `problem_name` pointing at an instance, every episode is the depot, a
    sample of the instance's customers and vehicles of the instance's capacity; step-wise API and fused rollout run on it
    (total demand 24 against 2 x 20: no node can be left without a vehicle)."""
    import torch
    from oracle.cvrp import CvrpOracle
    from tpc_b200.agent import Agent
    from tpc_b200.configs import get_configs
    from tpc_b200.env import env_factory
    with open(tmp_path / "T-n9-k2.vrp", "w") as f:
        f.write(_INSTANCE)
    with open(tmp_path / "opt-T-n9-k2", "w") as f:
        f.write("Route #1: 1 2 8 5\nRoute #2: 3 7 6 4\ncost 77\n")
    agent_cfg, _, env_cfg, _, _, _ = get_configs("CVRP", "tpc")
    env_cfg = json.loads(json.dumps(env_cfg))
    Nn, V, B = 6, 2, 16
    env_cfg.update(batch_size=B, node_sample_size=Nn, vehicle_sample_size=V, dir_path=str(tmp_path), problem_name="T-n9-k2")
    env, _ = env_factory("custom", "CVRP", env_cfg)
    assert env.instance_name == "T-n9-k2" and env.optimum_value == 77 and env.vehicle_capacity == 20
    state, bp, item, mha = env.reset()
    customers = {(10, 20), (30, 45), (70, 80), (95, 10), (20, 90), (55, 5), (60, 60), (5, 5)}
    assert state.shape == (B, Nn + V, 3) and np.all(state[:, 0] == np.array([40, 50, 0], np.float32))
    assert np.all(state[:, Nn:] == np.array([40, 50, 20], np.float32)) and np.all(state[:, 1:Nn, 2] == 3)
    for b in range(B):
        rows = {(int(x), int(y)) for x, y in state[b, 1:Nn, :2]}
        assert len(rows) == Nn - 1 and rows <= customers
    assert np.all(bp[:, :Nn] == 1) and np.all(bp[:, Nn:] == 0) and np.all(item[:, 0] == 1) and np.all(mha == 0)
    # one reference step: vehicle Nn serves node 1
    nstate, reward, done, info = env.step(np.full(B, Nn), np.full(B, 1))
    want = -np.round(np.sqrt(((state[:, 1, :2] - state[:, Nn, :2]).astype(np.float64) ** 2).sum(1)))
    assert np.array_equal(np.asarray(reward).reshape(-1), want.astype(np.float32)) and not done
    assert np.all(nstate[:, Nn, 2] == 17) and np.all(nstate[:, Nn, :2] == state[:, 1, :2])
    # fused greedy rollout on instance batches, replayed through the oracle
    agent_cfg = env.add_stats_to_agent_config(json.loads(json.dumps(agent_cfg)))
    agent = Agent("transformer", agent_cfg, "cuda:0", num_features=3, seed=6)
    bufs = env.device_buffers(store=True)
    env.reset_into(bufs)
    init = bufs.batch.cpu().numpy().copy()
    assert np.all(init[:, 0] == np.array([40, 50, 0], np.float32)) and not np.array_equal(init, state)   # a new draw
    agent.rollout(env, bufs, greedy=True)
    torch.cuda.synchronize()
    ora = CvrpOracle(Nn, V)
    st, obp, _, omha = ora.load(init)
    acts = bufs.actions.cpu().numpy()
    for t in range(env.steps):
        node = ora.request_node(t)
        assert np.array_equal(bufs.states[t].cpu().numpy(), st)
        st, rw, done, info = ora.step(acts[t], np.full(B, node))
        assert np.array_equal(bufs.rewards[:, t].cpu().numpy(), rw.reshape(B))
    assert done and np.array_equal(bufs.batch.cpu().numpy(), st)
