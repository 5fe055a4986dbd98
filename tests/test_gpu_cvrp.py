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
