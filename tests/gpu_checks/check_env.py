"""GPU check: environment kernels vs reference-generated goldens (bit-exact), reset properties, sampling,
fused rollout vs oracle replay, Adam vs the Keras-Adam oracle."""
import ctypes as C
import glob
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import model as om  # noqa: E402
from oracle.env import KnapsackV2Oracle, ResourceV3Oracle  # noqa: E402
from tpc_b200 import lib as L  # noqa: E402
from tpc_b200.engine import Engine, RolloutTensors, env_config_struct  # noqa: E402
from tests.gpu_checks.check_model import make_cfg  # noqa: E402

dev = torch.device("cuda:0")


def bits_equal(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def check_goldens(lib, h):
    gd = os.path.join(ROOT, "tests", "golden")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for path in sorted(glob.glob(os.path.join(gd, "env_*.npz"))):
        case = os.path.basename(path)[4:-4]
        g = np.load(path)
        cfg = json.load(open(path[:-4] + ".json"))
        kind = 1 if case.startswith("kv2") else 0
        e = env_config_struct(cfg, kind)
        T, B, S, Fs = g["state"].shape
        F = cfg["num_features"]
        N = (cfg["bin_sample_size"] if kind else cfg["node_sample_size"]) + 1
        R = S - N
        cost = Fs > F
        batch = torch.tensor(g["state"][0][:, :, :F].copy(), device=dev)
        is_empty = torch.tensor(g["state"][0][:, :, F].copy(), device=dev) if cost else None
        bin_mask = torch.tensor(g["bin_net_mask"][0], device=dev)
        res_mask = torch.tensor(1 - g["bin_net_mask"][0], device=dev)
        mha = torch.tensor(g["mha_mask"][0].reshape(B, S), device=dev)
        feas = torch.empty(B, S, device=dev)
        reward = torch.empty(B, device=dev)
        for t in range(T):
            # API-exact feasible mask on the reference-format state (B,S,Fs)
            st_ref = torch.tensor(g["state"][t], device=dev)
            dec = torch.tensor(g["dec_input"][t].reshape(B, F), device=dev)
            L.check(lib.tpc_env_feasible_mask(h, C.byref(e), L.ptr(st_ref), Fs, L.ptr(dec), L.ptr(bin_mask), B, S,
                                              L.ptr(feas), stream))
            assert bits_equal(feas.cpu().numpy(), g["feasible_mask"][t]), (case, t, "feasible")
            assert bits_equal(batch.cpu().numpy(), g["state"][t][:, :, :F]), (case, t, "state")
            act = torch.tensor(g["action"][t], device=dev, dtype=torch.int32)
            L.check(lib.tpc_env_step(h, C.byref(e), L.ptr(batch), L.ptr(bin_mask), L.ptr(res_mask), L.ptr(mha),
                                     L.ptr(is_empty), L.ptr(act), N + t, B, N, R, L.ptr(reward), stream))
            torch.cuda.synchronize()
            assert bits_equal(reward.cpu().numpy(), g["reward"][t]), (case, t, "reward", reward.cpu().numpy(), g["reward"][t])
            assert bits_equal(batch.cpu().numpy(), g["next_state"][t][:, :, :F]), (case, t, "next_state")
            if cost:
                assert bits_equal(is_empty.cpu().numpy(), g["next_state"][t][:, :, F]), (case, t, "is_empty")
            assert bits_equal(bin_mask.cpu().numpy(), g["next_bin_net_mask"][t]), (case, t, "bin_mask")
            assert bits_equal(mha.cpu().numpy(), g["next_mha_mask"][t].reshape(B, S)), (case, t, "mha")
            assert bits_equal(res_mask.cpu().numpy(), g["next_resource_net_mask"][t]), (case, t, "res_mask")
        print(f"golden {case}: {T} steps bit-exact")


def check_reset(lib, h):
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "env_rv3_shipped_greedy.json")))
    e = env_config_struct(cfg, 0)
    P, F = cfg["num_profiles"], 3
    prof = torch.empty(P, F, device=dev)
    L.check(lib.tpc_env_generate_profiles(h, C.byref(e), 1235, L.ptr(prof), stream))
    B, N, R = 64, 11, 20
    S = N + R

    def reset(B, id0, epoch):
        batch = torch.empty(B, S, F, device=dev)
        bm, mm = torch.empty(B, S, device=dev), torch.empty(B, S, device=dev)
        L.check(lib.tpc_env_reset(h, C.byref(e), 1235, id0, epoch, L.ptr(prof), B, N, R, L.ptr(batch), L.ptr(bm),
                                  L.ptr(mm), None, stream))
        torch.cuda.synchronize()
        return batch.cpu().numpy(), bm.cpu().numpy(), mm.cpu().numpy()

    p = prof.cpu().numpy()
    k = np.rint(p * 100).astype(int)
    assert k.min() >= 1 and k.max() <= 29 and bits_equal((k / 100).astype(np.float32), p)
    b0, bm, mm = reset(B, 0, 0)
    assert np.all(b0[:, 0] == -2)
    kn = np.rint(b0[:, 1:N] * 100).astype(int)
    assert kn.min() >= 0 and kn.max() <= 99 and bits_equal((kn / 100).astype(np.float32), b0[:, 1:N])
    assert 40 < kn.mean() < 60
    for b in range(B):  # rules = R distinct rows of the profile table
        rows = [tuple(r) for r in b0[b, N:]]
        idx = [np.flatnonzero((p == np.array(r, dtype=np.float32)).all(1)) for r in rows]
        assert all(len(i) > 0 for i in idx)
    assert np.all(bm[:, :N] == 0) and np.all(bm[:, N:] == 1) and np.all(mm == 0)
    b1, _, _ = reset(B, 0, 0)
    assert bits_equal(b0, b1), "reset must be deterministic"
    b2, _, _ = reset(32, 32, 0)
    assert bits_equal(b0[32:], b2), "episode streams must not depend on the batch partition"
    b3, _, _ = reset(B, 0, 1)
    assert not np.array_equal(b0, b3)
    print("reset: ranges, distinct profiles, determinism, partition independence OK")


def check_sampling(lib, h):
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    B, S = 200000, 7
    p = np.array([0.5, 0.0, 0.25, 0.0, 0.125, 0.125, 0.0], dtype=np.float32)
    probs = torch.tensor(np.tile(p, (B, 1)), device=dev)
    out = torch.empty(B, dtype=torch.int32, device=dev)
    L.check(lib.tpc_sample_categorical(h, L.ptr(probs), B, S, 7, 0, 3, L.ptr(out), stream))
    torch.cuda.synchronize()
    freq = np.bincount(out.cpu().numpy(), minlength=S) / B
    assert np.abs(freq - p).max() < 5e-3, freq
    assert freq[1] == 0 and freq[3] == 0 and freq[6] == 0
    print("sampling: empirical frequencies", np.round(freq, 4))


def check_rollout_and_adam():
    rng = np.random.default_rng(3)
    for reward in ("greedy", "global_dominant", "reduced_node_usage", "single_node_dominant"):
        cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "env_rv3_shipped_greedy.json")))
        cfg["reward"]["type"] = reward
        N, R, B, F = 11, 20, 48, 3
        S = N + R
        cost = reward == "reduced_node_usage"
        mc, ca, cc = make_cfg(N, R, common=not cost, Fb=4 if cost else 3)
        eng = Engine(mc, dev)
        pa = om.init_params(om.actor_spec(ca), 5) + (rng.standard_normal(eng.n_params[0]) * 0.02).astype(np.float32)
        eng.set_params(0, pa)
        e = env_config_struct(cfg, 0)
        prof = torch.empty(cfg["num_profiles"], F, device=dev)
        L.check(eng.lib.tpc_env_generate_profiles(eng.h, C.byref(e), 1235, L.ptr(prof), eng.stream))
        for greedy in (True, False):
            bufs = RolloutTensors(B, N, R, F, dev, cost=cost, keep_logits=True)
            L.check(eng.lib.tpc_env_reset(eng.h, C.byref(e), 1235, 0, 7, L.ptr(prof), B, N, R, L.ptr(bufs.batch),
                                          L.ptr(bufs.bin_net_mask), L.ptr(bufs.mha_mask), L.ptr(bufs.is_empty), eng.stream))
            init = bufs.batch.cpu().numpy().copy()
            eng.rollout(e, bufs, greedy, 99, 0, 0)
            torch.cuda.synchronize()
            # replay the CUDA actions through the oracle env: every stored tensor must match bit for bit
            env = ResourceV3Oracle(cfg)
            state, dec, bm, mh = env.load(init)
            acts = bufs.actions.cpu().numpy()
            fa = torch.tensor(pa)
            agree, total, big_margin_bad = 0, 0, 0
            for t in range(R):
                feas = env.build_feasible_mask(state, dec, bm)
                assert bits_equal(bufs.states[t].cpu().numpy(), state[:, :, :F]), (reward, t, "states")
                if cost:
                    assert bits_equal(bufs.is_empties[t].cpu().numpy(), state[:, :, F]), (reward, t, "is_empty")
                assert bits_equal(bufs.dec_inputs[t].cpu().numpy(), dec.reshape(B, F)), (reward, t, "dec")
                assert bits_equal(bufs.ptr_masks[t].cpu().numpy(), feas), (reward, t, "feasible")
                assert bits_equal(bufs.mha_masks[t].cpu().numpy(), mh.reshape(B, S)), (reward, t, "mha")
                a = acts[t]
                assert np.all(feas[np.arange(B), a] == 0), "picked a masked position"
                ol, op, oi = om.actor_forward(fa, ca, torch.tensor(state), torch.tensor(dec), torch.tensor(feas),
                                              torch.tensor(mh), N)
                cl = bufs.logits[t].cpu().numpy()
                unm = feas == 0
                err = np.abs(cl[unm] - ol.numpy()[unm]).max() / np.abs(ol.numpy()[unm]).max()
                assert err < 5e-3, (reward, t, err)
                if greedy:
                    srt = np.sort(np.where(unm, ol.numpy(), -np.inf), axis=1)
                    margin = srt[:, -1] - srt[:, -2]
                    same = (a == oi.numpy())
                    agree += int(same.sum())
                    total += B
                    big_margin_bad += int(((~same) & (margin > 1e-3 * np.abs(srt[:, -1]).clip(1e-3))).sum())
                state, dec, rw, done, info = env.step(a, feas)
                assert bits_equal(bufs.rewards[:, t].cpu().numpy(), np.asarray(rw, dtype=np.float32).reshape(B)), (reward, t, "reward")
                bm, mh = info["bin_net_mask"], info["mha_used_mask"]
            assert done
            assert bits_equal(bufs.batch.cpu().numpy(), state[:, :, :F])
            if greedy:
                print(f"rollout {reward} greedy: trajectory bit-exact vs oracle env; argmax agree {agree}/{total}, "
                      f"disagreements with margin > tol: {big_margin_bad}")
                assert big_margin_bad == 0
            else:
                print(f"rollout {reward} sampled: trajectory bit-exact vs oracle env; mean reward {bufs.rewards.sum(1).mean().item():.3f}")
        eng.close()

    # Adam vs Keras-Adam oracle
    mc, ca, cc = make_cfg(6, 9)
    eng = Engine(mc, dev)
    for which, spec in ((0, om.actor_spec(ca)), (1, om.critic_spec(cc))):
        n = eng.n_params[which]
        p0 = om.init_params(spec, 11)
        p, m, v = p0.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
        eng.set_params(which, p0)
        for step in range(1, 4):
            g = (rng.standard_normal(n) * (10.0 if step == 2 else 0.01)).astype(np.float32)
            eng.grads[which].copy_(torch.tensor(g))
            eng.adam_apply(which, 1e-3, 1.0, grad_scale=0.5)
            om.keras_adam_step(p, g * np.float32(0.5), m, v, step, 1e-3, spec, clipnorm=1.0)
            torch.cuda.synchronize()
            err = np.abs(eng.params[which].cpu().numpy() - p).max()
            assert err < 2e-6, (which, step, err)
        print(f"adam which={which}: max abs param diff after 3 steps {err:.2e}")
    eng.close()


def check_config2_corner(nodes, rules, B=1024, n_check=16, reward="greedy"):
    """BASELINE config #2 at its own size: greedy decode of a batch-1024 tester cell on the device (fused rollout),
    then `n_check` of those episodes (spread over the batch, addressed by their global episode ids) are replayed
    through the oracle environment and the oracle actor: stored states / masks / rewards bit-exact, logits within
    1e-3, and every pick equal to the oracle's argmax wherever the oracle's top-2 logit margin exceeds 1e-3."""
    rng = np.random.default_rng(nodes * 1000 + rules)
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "env_rv3_shipped_greedy.json")))
    cfg["reward"]["type"] = reward
    cfg.update(batch_size=B, node_sample_size=nodes, profiles_sample_size=rules)
    N, R, F = nodes + 1, rules, 3
    S = N + R
    mc, ca, cc = make_cfg(N, R)
    eng = Engine(mc, dev)
    pa = om.init_params(om.actor_spec(ca), 5) + (rng.standard_normal(eng.n_params[0]) * 0.02).astype(np.float32)
    eng.set_params(0, pa)
    e = env_config_struct(cfg, 0)
    prof = torch.empty(cfg["num_profiles"], F, device=dev)
    L.check(eng.lib.tpc_env_generate_profiles(eng.h, C.byref(e), 1235, L.ptr(prof), eng.stream))
    bufs = RolloutTensors(B, N, R, F, dev, keep_logits=True)
    L.check(eng.lib.tpc_env_reset(eng.h, C.byref(e), 1235, 0, 3, L.ptr(prof), B, N, R, L.ptr(bufs.batch),
                                  L.ptr(bufs.bin_net_mask), L.ptr(bufs.mha_mask), L.ptr(bufs.is_empty), eng.stream))
    init = bufs.batch.cpu().numpy().copy()
    eng.rollout(e, bufs, True, 99, 0, 0)
    torch.cuda.synchronize()
    ids = np.unique(np.linspace(0, B - 1, n_check).astype(int))
    n = len(ids)
    sub = dict(cfg)
    sub["batch_size"] = n
    env = ResourceV3Oracle(sub)
    state, dec, bm, mh = env.load(init[ids])
    acts = bufs.actions.cpu().numpy()[:, ids]
    states, pms, mhs = (x.cpu().numpy()[:, ids] for x in (bufs.states, bufs.ptr_masks, bufs.mha_masks))
    logits, rewards = bufs.logits.cpu().numpy()[:, ids], bufs.rewards.cpu().numpy()[ids]
    fa = torch.tensor(pa)
    agree = bad = 0
    worst = 0.0
    for t in range(R):
        feas = env.build_feasible_mask(state, dec, bm)
        assert bits_equal(states[t], state[:, :, :F]), (t, "states")
        assert bits_equal(pms[t], feas), (t, "feasible")
        assert bits_equal(mhs[t], mh.reshape(n, S)), (t, "mha")
        ol, op, oi = om.actor_forward(fa, ca, torch.tensor(state), torch.tensor(dec), torch.tensor(feas),
                                      torch.tensor(mh), N)
        ol = ol.numpy()
        unm = feas == 0
        worst = max(worst, float(np.abs(logits[t][unm] - ol[unm]).max() / np.abs(ol[unm]).max()))
        srt = np.sort(np.where(unm, ol, -np.inf), axis=1)
        margin = srt[:, -1] - srt[:, -2]
        same = acts[t] == oi.numpy()
        agree += int(same.sum())
        bad += int(((~same) & (margin > 1e-3 * np.abs(srt[:, -1]).clip(1e-3))).sum())
        state, dec, rw, done, info = env.step(acts[t], feas)
        assert bits_equal(rewards[:, t], np.asarray(rw, dtype=np.float32).reshape(n)), (t, "reward")
        bm, mh = info["bin_net_mask"], info["mha_used_mask"]
    assert done and bits_equal(bufs.batch.cpu().numpy()[ids], state[:, :, :F])
    print(f"config #2 cell {nodes} x {rules}, batch {B}: {n} episodes replayed bit-exact; logits rel err {worst:.2e}; "
          f"picks equal {agree}/{n * R}, disagreements with margin > 1e-3: {bad}")
    assert worst < 1e-3 and bad == 0
    eng.close()
    return worst


def main():
    lib = L.load()
    h = C.c_void_p()
    L.check(lib.tpc_create(C.byref(h)))
    check_goldens(lib, h)
    check_reset(lib, h)
    check_sampling(lib, h)
    check_rollout_and_adam()
    print("OK")


if __name__ == "__main__":
    main()
