"""CPU reference path timing (oracle port of agents/trainer.py:49-156): per-step actor forward on the live batch,
then one batched forward/backward of critic and actor over the stored states and two Adam steps.

TEST / BASELINE INFRASTRUCTURE ONLY: used by bench.py's `cpu_baseline` leg and `--impl reference` arm.
The reference itself (TensorFlow 2.3.1) cannot be installed here, so this torch-CPU restatement with the same
execution structure stands in for it (`cpu_baseline.kind = "port"`).
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from . import model as om
from .env import ResourceV3Oracle


def make_cfgs(N, R, F=3):
    a = om.NetCfg(num_layers=1, dff=128, dec_features=F, num_bin_features=F, num_resource_features=F)
    c = om.NetCfg(num_layers=3, dff=512, dec_features=F, num_bin_features=F, num_resource_features=F, tensor_size=N + R)
    return a, c


def random_instances(rng, B, N, R, F=3):
    S = N + R
    batch = np.zeros((B, S, F), dtype=np.float32)
    batch[:, 1:N] = rng.integers(0, 100, size=(B, N - 1, F)).astype(np.float32) / 100
    batch[:, 0] = -2.0
    batch[:, N:] = rng.integers(1, 30, size=(B, R, F)).astype(np.float32) / 100
    return batch


def time_sample(env_cfg: dict, N: int, R: int, sample_states: int, rollout_steps: int = 1, threads: int | None = None,
                seed: int = 0):
    """Times (a) `rollout_steps` env+actor steps on `sample_states` episodes and (b) the critic + actor
    forward/backward/Adam on `sample_states` stored states; returns decisions/s extrapolated to a full iteration
    (every decision costs one rollout step + one update state) and the raw timings."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    rng = np.random.default_rng(seed)
    ca, cc = make_cfgs(N, R)
    fa = torch.tensor(om.init_params(om.actor_spec(ca), 1), requires_grad=True)
    fc = torch.tensor(om.init_params(om.critic_spec(cc), 2), requires_grad=True)
    cfg = dict(env_cfg)
    cfg.update(batch_size=sample_states, node_sample_size=N - 1, profiles_sample_size=R)
    env = ResourceV3Oracle(cfg)
    state, dec, bm, mh = env.load(random_instances(rng, sample_states, N, R))
    S = N + R
    with torch.no_grad():  # untimed warm-up (thread pool, allocator)
        om.actor_forward(fa, ca, torch.tensor(state), torch.tensor(dec), torch.tensor(bm), torch.tensor(mh), N)
    t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(rollout_steps):
            feas = env.build_feasible_mask(state, dec, bm)
            _l, probs, _i = om.actor_forward(fa, ca, torch.tensor(state), torch.tensor(dec), torch.tensor(feas),
                                             torch.tensor(mh), N)
            p = probs.numpy().astype(np.float64)
            p /= p.sum(-1, keepdims=True)
            acts = np.array([rng.choice(S, p=p[b]) for b in range(sample_states)])
            state, dec, rw, done, info = env.step(acts, feas)
            bm, mh = info["bin_net_mask"], info["mha_used_mask"]
    t_roll = (time.perf_counter() - t0) / rollout_steps
    # update on `sample_states` stored states
    st = torch.tensor(state)
    m4 = torch.tensor(mh)
    feas = env.build_feasible_mask(state, dec, bm) if not done else np.zeros((sample_states, S), np.float32)
    R_t = torch.randn(sample_states, 1)
    acts_t = torch.zeros(sample_states, dtype=torch.long)
    m_a, v_a = np.zeros(fa.numel(), np.float32), np.zeros(fa.numel(), np.float32)
    m_c, v_c = np.zeros(fc.numel(), np.float32), np.zeros(fc.numel(), np.float32)
    t0 = time.perf_counter()
    vals = om.critic_forward(fc, cc, st, m4, N)
    vl, adv = om.value_loss(vals, R_t, 1.0)
    vl.backward()
    logits, _p, _i = om.actor_forward(fa, ca, st, torch.tensor(np.ascontiguousarray(state[:, N:N + 1, :3])),
                                      torch.tensor(feas), m4, N)
    tot, _e, _pl, _pp = om.actor_loss(logits, acts_t, adv.detach(), 0.01)
    tot.backward()
    pa, pc = fa.detach().numpy().copy(), fc.detach().numpy().copy()
    om.keras_adam_step(pc, fc.grad.numpy(), m_c, v_c, 1, 5e-4, om.critic_spec(cc), clipnorm=1.0)
    om.keras_adam_step(pa, fa.grad.numpy(), m_a, v_a, 1, 1e-4, om.actor_spec(ca), clipnorm=1.0)
    t_upd = time.perf_counter() - t0
    per_decision = t_roll / sample_states + t_upd / sample_states
    return {"decisions_per_s": 1.0 / per_decision, "t_rollout_step_s": t_roll, "t_update_s": t_upd,
            "sample_states": sample_states, "threads": threads}


def time_full_iteration(env_cfg: dict, N: int, R: int, episodes: int = 32, threads: int | None = None, seed: int = 0):
    """One COMPLETE training iteration with the reference's execution structure (agents/trainer.py:37-175) at a reduced
    batch: a full T = R step rollout of `episodes` episodes (per step: feasible mask, actor forward on the live batch,
    categorical draw, env step, store), Monte-Carlo returns, then ONE critic and ONE actor forward/backward over the
    T * episodes stored states and both Adam steps. Nothing is extrapolated except the batch size.
    Returns decisions/s = T * episodes / wall time and the split."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    rng = np.random.default_rng(seed)
    ca, cc = make_cfgs(N, R)
    fa = torch.tensor(om.init_params(om.actor_spec(ca), 1), requires_grad=True)
    fc = torch.tensor(om.init_params(om.critic_spec(cc), 2), requires_grad=True)
    cfg = dict(env_cfg)
    cfg.update(batch_size=episodes, node_sample_size=N - 1, profiles_sample_size=R)
    env = ResourceV3Oracle(cfg)
    B, S, T = episodes, N + R, R
    t_start = time.perf_counter()
    state, dec, bm, mh = env.load(random_instances(rng, B, N, R))
    states, decs, masks, mhas, acts_l = [], [], [], [], []
    rewards = np.zeros((B, T), np.float32)
    with torch.no_grad():
        for t in range(T):
            feas = env.build_feasible_mask(state, dec, bm)
            _l, probs, _i = om.actor_forward(fa, ca, torch.tensor(state), torch.tensor(dec), torch.tensor(feas),
                                             torch.tensor(mh), N)
            cdf = np.cumsum(probs.numpy().astype(np.float64), axis=-1)
            u = rng.random((B, 1)) * cdf[:, -1:]
            acts = np.minimum((u > cdf).sum(-1), S - 1)
            acts = np.where(feas[np.arange(B), acts] == 0, acts, 0)
            states.append(state.copy()); decs.append(dec.copy()); masks.append(feas.copy()); mhas.append(mh.copy())
            acts_l.append(acts)
            state, dec, rw, done, info = env.step(acts, feas)
            rewards[:, t] = np.asarray(rw, dtype=np.float32).reshape(B)
            bm, mh = info["bin_net_mask"], info["mha_used_mask"]
    t_roll = time.perf_counter() - t_start
    st = torch.tensor(np.concatenate(states, 0))
    m4 = torch.tensor(np.concatenate(mhas, 0))
    returns = om.compute_discounted_rewards(rewards, 0.99)
    R_t = torch.tensor(np.ascontiguousarray(returns.T).reshape(-1, 1))
    vals = om.critic_forward(fc, cc, st, m4, N)
    vl, adv = om.value_loss(vals, R_t, 1.0)
    vl.backward()
    logits, _p, _i = om.actor_forward(fa, ca, st, torch.tensor(np.concatenate(decs, 0)),
                                      torch.tensor(np.concatenate(masks, 0)), m4, N)
    tot, _e, _pl, _pp = om.actor_loss(logits, torch.tensor(np.concatenate(acts_l, 0)), adv.detach(), 0.01)
    tot.backward()
    pa, pc = fa.detach().numpy().copy(), fc.detach().numpy().copy()
    m_a, v_a = np.zeros(fa.numel(), np.float32), np.zeros(fa.numel(), np.float32)
    m_c, v_c = np.zeros(fc.numel(), np.float32), np.zeros(fc.numel(), np.float32)
    om.keras_adam_step(pc, fc.grad.numpy(), m_c, v_c, 1, 5e-4, om.critic_spec(cc), clipnorm=1.0)
    om.keras_adam_step(pa, fa.grad.numpy(), m_a, v_a, 1, 1e-4, om.actor_spec(ca), clipnorm=1.0)
    wall = time.perf_counter() - t_start
    return {"decisions_per_s": B * T / wall, "t_rollout_s": t_roll, "t_update_s": wall - t_roll, "wall_s": wall,
            "episodes": B, "steps": T, "threads": threads}
