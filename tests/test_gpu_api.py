"""GPU tests of the reference-facing Python API (Agent / env / trainer / tester) and full-size property tests."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _make(batch=8, nodes=5, rules=7, reward="greedy"):
    import torch
    from tpc_b200.agent import Agent
    from tpc_b200.configs import get_configs
    from tpc_b200.env import ResourceEnvironmentV3
    agent_cfg, trainer_cfg, env_cfg, tester_cfg, _, _ = get_configs("ResourceV3", "tpc")
    env_cfg = json.loads(json.dumps(env_cfg))
    env_cfg.update(batch_size=batch, node_sample_size=nodes, profiles_sample_size=rules)
    env_cfg["reward"]["type"] = reward
    env = ResourceEnvironmentV3("ResourceV3", env_cfg, "cuda:0")
    env.add_stats_to_agent_config(agent_cfg)
    agent = Agent("transformer", agent_cfg, "cuda:0", num_features=3, seed=1)
    return env, agent, trainer_cfg, tester_cfg, env_cfg


def test_stepwise_api_matches_oracle_and_fused_update():
    """Drive the reference's trainer protocol step by step (act / env.step / store), then check that
    (a) every env tensor equals the NumPy oracle bit for bit, (b) compute_value_loss / compute_actor_loss agree
    with the fused device update on the same memory within 1e-3, (c) act-time probs == loss-time probs to 5
    decimals (agents_test.py:217-220)."""
    import torch
    from oracle.env import ResourceV3Oracle
    env, agent, *_ , env_cfg = _make()
    state, dec, bin_mask, mha = env.reset()
    ora = ResourceV3Oracle(env_cfg)
    ora.load(state)
    act_probs = []
    done, t = False, 0
    while not done:
        bin_id, feas, probs = agent.act(state, dec, bin_mask, mha, env.build_feasible_mask)
        o_state, o_dec, o_bm, o_mh = ora.state()
        assert np.array_equal(feas, ora.build_feasible_mask(o_state, o_dec, o_bm))
        nstate, ndec, reward, done, info = env.step(bin_id, feas)
        os_, od_, orw, odone, oinfo = ora.step(np.asarray(bin_id), feas)
        assert np.array_equal(nstate, os_) and np.array_equal(np.asarray(reward), np.asarray(orw, dtype=np.float32))
        assert np.array_equal(info["bin_net_mask"], oinfo["bin_net_mask"])
        assert np.array_equal(info["mha_used_mask"], oinfo["mha_used_mask"]) and done == odone
        agent.store(state.copy(), dec.copy(), feas.copy(), mha.copy(), bin_id.numpy().copy(), reward.numpy().copy(), t)
        act_probs.append(np.asarray(probs))
        state, bin_mask, mha = nstate, info["bin_net_mask"], info["mha_used_mask"]
        if not done:
            dec = ndec.copy()
        t += 1
    assert t == env.profiles_sample_size
    R = agent.compute_discounted_rewards(np.zeros((agent.batch_size, 1), np.float32))
    vloss, values, adv = agent.compute_value_loss(R)
    total, _, entropy, ploss, probs = agent.compute_actor_loss(agent.bin_actor, agent.bin_masks, agent.bins,
                                                               agent.actor_decoder_input, adv)
    np.testing.assert_almost_equal(np.concatenate(act_probs), np.asarray(probs), decimal=5)
    # independent host recomputation (float64) of what the device loss kernels return (agents/agent.py:110-242)
    rw = agent.rewards.astype(np.float64)
    Rref = np.zeros_like(rw)
    boot = np.zeros(rw.shape[0])
    for step in reversed(range(rw.shape[1])):
        boot = rw[:, step] + agent.gamma * boot
        Rref[:, step] = boot
    np.testing.assert_allclose(R, Rref, rtol=1e-6, atol=1e-6)
    n = values.shape[0]
    Rv = np.concatenate([Rref[:, t_] for t_ in range(rw.shape[1])])             # reshape_into_vertical_format
    v64 = np.asarray(values, dtype=np.float64).reshape(-1)
    np.testing.assert_allclose(np.asarray(adv).reshape(-1), Rv - v64, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(vloss, agent.values_loss_coefficient * np.mean((Rv - v64) ** 2), rtol=1e-5)
    lg = np.asarray(agent.bin_actor(np.concatenate(agent.states), np.concatenate(agent.actor_decoder_input),
                                    np.concatenate(agent.bin_masks), True, agent.num_bins,
                                    enc_padding_mask=np.concatenate(agent.mha_masks),
                                    dec_padding_mask=np.concatenate(agent.mha_masks))[0], dtype=np.float64)
    mx = lg.max(axis=-1, keepdims=True)
    logp = lg - (mx + np.log(np.exp(lg - mx).sum(axis=-1, keepdims=True)))
    acts = np.concatenate([np.asarray(a) for a in agent.bins]).astype(np.int64).reshape(-1)
    pl_ref = -logp[np.arange(n), acts]
    ent_ref = -(np.exp(logp) * logp).sum(axis=-1)
    np.testing.assert_allclose(np.asarray(ploss), pl_ref, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(np.asarray(entropy), ent_ref, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(total, np.mean(pl_ref * (Rv - v64) - agent.entropy_coefficient * ent_ref), rtol=1e-4, atol=1e-5)
    bufs = agent.memory_to_device()
    losses = agent.update(bufs).cpu().numpy()
    ref = np.array([vloss, np.mean(ploss), total, np.mean(entropy)], dtype=np.float64)
    assert np.all(np.abs(losses - ref) <= 1e-3 * np.maximum(1.0, np.abs(ref))), (losses, ref)


def test_trainer_and_tester_run(tmp_path):
    from tpc_b200.tester import test as tester
    from tpc_b200.trainer import trainer
    env, agent, trainer_cfg, tester_cfg, _ = _make(batch=16, nodes=5, rules=8)
    trainer_cfg = dict(trainer_cfg, n_iterations=3)
    before = agent.bin_actor.get_flat()
    out = trainer(env, agent, trainer_cfg, False, str(tmp_path))
    assert len(out) == 7 and all(len(x) == 3 for x in out)
    assert all(np.isfinite(v) for buf in out for v in buf)
    assert not np.array_equal(before, agent.bin_actor.get_flat())          # weights moved
    path = os.path.join(str(tmp_path), "model", "actor.npz")
    assert os.path.exists(path)                                            # trainer.py:178-184
    saved = agent.bin_actor.get_flat()
    agent.init_weights(5)
    agent.load_weights(os.path.join(str(tmp_path), "model", "actor"))
    assert np.array_equal(saved, agent.bin_actor.get_flat())
    tcfg = json.loads(json.dumps(tester_cfg))
    tcfg["batch_size"] = 4
    tcfg["show_per_test_stats"] = False
    tcfg["testbed"].update(num_tests=1, node_sample_configs={"min": 5, "max": 10, "step": 5},
                           request_sample_configs={"min": 10, "max": 20, "step": 10})
    dominant, rejected = tester(env, agent, tcfg, str(tmp_path))     # [won, draw, lost] vs the best heuristic
    assert dominant.shape == rejected.shape == (3,) and dominant.sum() == rejected.sum() == 2 * 2 * 4
    assert os.path.exists(os.path.join(str(tmp_path), "tests", "test.csv"))


@pytest.mark.parametrize("reward", ["greedy", "reduced_node_usage"])
def test_full_size_rollout_properties(reward):
    """BASELINE full size (batch 1024, 50 nodes x 100 rules): size-independent invariants of a greedy rollout."""
    import torch
    from tpc_b200.agent import Agent
    from tpc_b200.configs import get_configs
    from tpc_b200.env import ResourceEnvironmentV3
    agent_cfg, _, env_cfg, _, _, _ = get_configs("ResourceV3", "tpc")
    env_cfg = json.loads(json.dumps(env_cfg))
    env_cfg.update(batch_size=1024, node_sample_size=50, profiles_sample_size=100)
    env_cfg["reward"]["type"] = reward
    env = ResourceEnvironmentV3("ResourceV3", env_cfg, "cuda:0")
    env.add_stats_to_agent_config(agent_cfg)
    agent = Agent("transformer", agent_cfg, "cuda:0", num_features=3, seed=2)
    bufs = env.device_buffers(store=True)
    env.reset_into(bufs)
    init = bufs.batch.clone()
    agent.rollout(env, bufs, greedy=True)
    torch.cuda.synchronize()
    N, R, B = 51, 100, 1024
    k0 = torch.round(init * 100).long()
    k1 = torch.round(bufs.batch * 100).long()
    acts = bufs.actions.long()                                  # (T,B)
    assert int(acts.min()) >= 0 and int(acts.max()) < N          # only node positions are ever picked
    assert torch.all(bufs.ptr_masks.gather(2, acts.unsqueeze(2)).squeeze(2) == 0)   # never a masked position
    assert torch.all(k1[:, 1:N] >= 0)                            # no node is ever overloaded
    assert torch.equal(k1[:, N:], k0[:, N:]) and torch.all(bufs.batch[:, 0] == -2)   # rules and EOS row untouched
    # conservation in integer hundredths: capacity removed from node n == sum of the demands placed on it
    placed = torch.zeros(B, N, 3, dtype=torch.long, device="cuda:0")
    demands = k0[:, N:, :].permute(1, 0, 2)                     # (T,B,3)
    placed.scatter_add_(1, acts.t().unsqueeze(2).expand(B, R, 3), demands.permute(1, 0, 2))
    assert torch.equal((k0 - k1)[:, 1:N], placed[:, 1:N])
    assert torch.all(bufs.mha_mask[:, N:] == 1)                  # every rule consumed exactly once
    if reward == "greedy":
        rejected = (acts == 0).sum(0)
        assert torch.equal(bufs.rewards.sum(1), (R - rejected).float())


def test_initializers_follow_get_initializer():
    """common/utils.py:49-64: `use_default_initializer` -> glorot_uniform, else U(+-1/sqrt(dims)) with dims = dim_model
    (encoder / decoder / critic head) or attention_dense_units (pointer, pointer_attention.py:13); biases 0, gamma 1."""
    from oracle import model as om
    from tpc_b200.agent import Agent
    from tpc_b200.configs import get_configs
    from tpc_b200.env import ResourceEnvironmentV3
    agent_cfg, _, env_cfg, _, _, _ = get_configs("ResourceV3", "tpc")
    env_cfg = json.loads(json.dumps(env_cfg))
    env_cfg.update(batch_size=4)
    env = ResourceEnvironmentV3("ResourceV3", env_cfg, "cuda:0")
    agent_cfg = json.loads(json.dumps(agent_cfg))
    env.add_stats_to_agent_config(agent_cfg)
    for default in (True, False):
        agent_cfg["actor"]["use_default_initializer"] = agent_cfg["critic"]["use_default_initializer"] = default
        agent = Agent("transformer", agent_cfg, "cuda:0", num_features=3, seed=3)
        for net, spec in ((agent.bin_actor, om.actor_spec(om.NetCfg(num_layers=1, dff=128))),
                          (agent.critic, om.critic_spec(om.NetCfg(num_layers=3, dff=512, tensor_size=31)))):
            flat = net.get_flat()
            offs, _ = om.spec_offsets(spec)
            for name, (o, shape) in offs.items():
                w = flat[o:o + int(np.prod(shape))]
                if name.endswith(".kernel"):
                    lim = np.sqrt(6.0 / (shape[0] + shape[1])) if default else 1.0 / np.sqrt(128)
                    assert np.abs(w).max() <= lim and np.abs(w).max() > 0.8 * lim * (w.size > 64), name
                elif name.endswith(".gamma"):
                    assert np.all(w == 1), name
                else:
                    assert np.all(w == 0), name
        agent.engine.close()


@pytest.mark.parametrize("nodes,rules", [(5, 10), (50, 100)])
def test_config2_corner_cells_replay_bit_exact_at_batch_1024(nodes, rules):
    """BASELINE config #2 (tester sweep 5-50 nodes x 10-100 rules, greedy, batch 1024): the two corner cells, 16
    episodes of each drawn out of the batch-1024 device rollout and replayed through the oracle (top-2-margin rule)."""
    from tests.gpu_checks import check_env
    check_env.check_config2_corner(nodes, rules, B=1024, n_check=16)


def test_knapsack_v2_trains_on_the_same_kernels(tmp_path):
    """BASELINE config #4: configs/KnapsackV2.json (2 features, separate embeddings, clipnorm null)."""
    import torch
    from tpc_b200.agent import Agent
    from tpc_b200.configs import get_configs
    from tpc_b200.env import KnapsackEnvironmentV2
    from tpc_b200.trainer import trainer
    from oracle.env import KnapsackV2Oracle
    agent_cfg, trainer_cfg, env_cfg, _, _, _ = get_configs("KnapsackV2", "tpc")
    env_cfg = json.loads(json.dumps(env_cfg))
    env_cfg.update(batch_size=32, bin_sample_size=5, item_sample_size=8)
    env = KnapsackEnvironmentV2("KnapsackV2", env_cfg, "cuda:0")
    env.add_stats_to_agent_config(agent_cfg)
    agent = Agent("transformer", agent_cfg, "cuda:0", num_features=2, seed=3)
    # one greedy rollout replayed through the NumPy oracle: bit-exact bookkeeping and rewards
    bufs = env.device_buffers(store=True)
    env.reset_into(bufs)
    init = bufs.batch.cpu().numpy().copy()
    agent.rollout(env, bufs, greedy=True)
    torch.cuda.synchronize()
    ora = KnapsackV2Oracle(env_cfg)
    state, dec, bm, mh = ora.load(init)
    acts = bufs.actions.cpu().numpy()
    for t in range(env.profiles_sample_size):
        feas = ora.build_feasible_mask(state, dec, bm)
        assert np.array_equal(bufs.ptr_masks[t].cpu().numpy(), feas)
        state, dec, rw, done, info = ora.step(acts[t], feas)
        assert np.array_equal(bufs.rewards[:, t].cpu().numpy(), np.asarray(rw, dtype=np.float32).reshape(-1))
        bm, mh = info["bin_net_mask"], info["mha_used_mask"]
    assert done and np.array_equal(bufs.batch.cpu().numpy(), state)
    out = trainer(env, agent, dict(trainer_cfg, n_iterations=2), False, str(tmp_path))
    assert all(np.isfinite(v) for buf in out for v in buf)


def test_runner_writes_the_reference_artefacts(tmp_path):
    """main.py:40-83: config.json, env.txt, training/logs.csv, model/actor.npz, tests/test.csv."""
    from tpc_b200.main import runner
    ov = {"trainer_config": {"n_iterations": 2},
          "env_config": {"batch_size": 8, "node_sample_size": 4, "profiles_sample_size": 6},
          "tester_config": {"batch_size": 3, "show_per_test_stats": False, "show_inference_progress": False,
                            "testbed": {"num_tests": 1, "node_sample_configs": {"min": 5, "max": 5, "step": 5},
                                        "node_available_resources": {"min": 0, "max": 100, "step": 100},
                                        "request_sample_configs": {"min": 10, "max": 10, "step": 10}}}}
    log_dir, history, (dom, rej) = runner(env_name="ResourceV3", log_root=str(tmp_path), overrides=ov)
    assert json.load(open(os.path.join(log_dir, "config.json")))["env_config"]["num_profiles"] == 1000
    assert np.loadtxt(os.path.join(log_dir, "env.txt")).shape == (1000, 3)
    logs = open(os.path.join(log_dir, "training", "logs.csv")).read().splitlines()
    assert logs[0] == "Step;Avg Reward;Max Reward;Min Reward;Value Loss;Bin Entropy;Total Bin Loss;Bin Policy Loss"
    assert len(logs) == 3 and logs[1].startswith("0;") and len(logs[1].split(";")) == 8
    assert os.path.exists(os.path.join(log_dir, "model", "actor.npz"))
    assert os.path.exists(os.path.join(log_dir, "tests", "test.csv")) and dom.sum() == rej.sum() == 3


@pytest.mark.parametrize("reward", ["single_node_dominant", "global_dominant", "reduced_node_usage"])
def test_training_with_every_reward_type(tmp_path, reward):
    """BASELINE config #3 (fair = global_dominant, cost = reduced_node_usage with the 4-feature node embedding) at a
    small shape: sampled rollouts + A2C updates run, losses stay finite, weights move, rewards follow reward.py's sign
    conventions (penalties are negative, environment/custom/resource_v3/reward.py:46-166)."""
    from tpc_b200.agent import Agent
    from tpc_b200.configs import get_configs
    from tpc_b200.env import ResourceEnvironmentV3
    from tpc_b200.trainer import trainer
    agent_cfg, trainer_cfg, env_cfg, _, _, _ = get_configs("ResourceV3", "tpc")
    env_cfg = json.loads(json.dumps(env_cfg))
    env_cfg.update(batch_size=32, node_sample_size=6, profiles_sample_size=12)
    env_cfg["reward"]["type"] = reward
    env = ResourceEnvironmentV3("ResourceV3", env_cfg, "cuda:0")
    env.add_stats_to_agent_config(agent_cfg)
    assert agent_cfg["encoder_embedding"]["common"] == (reward != "reduced_node_usage")
    agent = Agent("transformer", agent_cfg, "cuda:0", num_features=3, seed=11)
    before = agent.critic.get_flat().copy()
    out = trainer(env, agent, dict(trainer_cfg, n_iterations=3, store_model_weights=dict(
        trainer_cfg["store_model_weights"], export_weights=False)), False, str(tmp_path))
    assert all(np.isfinite(v) for buf in out for v in buf)
    assert not np.array_equal(before, agent.critic.get_flat())
    avg, mn, mx = out[0], out[1], out[2]
    assert all(a <= b + 1e-6 for a, b in zip(mn, avg)) and all(a <= b + 1e-6 for a, b in zip(avg, mx))


def test_episode_reward_stats_kernel():
    """trainer.py:89-100: min / max / mean over the batch of the per-episode returns."""
    import ctypes as C
    import torch
    from tpc_b200 import lib as L
    lib = L.load()
    h = C.c_void_p()
    assert lib.tpc_create(C.byref(h)) == 0
    rng = np.random.default_rng(3)
    for (B, T) in [(1, 1), (7, 5), (1024, 100), (4096, 20)]:
        rw = rng.normal(size=(B, T)).astype(np.float32)
        d = torch.as_tensor(rw).cuda()
        out = torch.empty(3, device="cuda")
        assert lib.tpc_episode_reward_stats(h, L.ptr(d), B, T, L.ptr(out),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
        ret = rw.astype(np.float64).sum(axis=1)
        np.testing.assert_allclose(out.cpu().numpy(), [ret.min(), ret.max(), ret.mean()], rtol=1e-5, atol=1e-5)
    lib.tpc_destroy(h)
