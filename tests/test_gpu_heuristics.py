"""GPU parity of the batched tester kernels (tpc_heuristic_dominant / _random, tpc_solution_stats,
tpc_compare_solutions, called through the Python mirror of heuristic/*.py) against the reference's own placements
(tests/golden/heuristic_resource_v3.npz) and against the pinned oracle on seeded batches: bit-exact everywhere."""
import json
import os

import numpy as np
import pytest

from oracle import heuristics as oh

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "heuristic_resource_v3.npz"))
N_CASES = int(G["num_cases"])
DEV = "cuda:0"


def _instances(rng, B, n_nodes, n_rules, lo, hi, F=3):
    s = np.zeros((B, n_nodes + n_rules, F), dtype=np.float32)
    s[:, 0] = -2
    s[:, 1:n_nodes, :3] = (rng.integers(lo, hi, size=(B, n_nodes - 1, 3)) / 100).astype(np.float32)
    s[:, n_nodes:, :3] = (rng.integers(1, 30, size=(B, n_rules, 3)) / 100).astype(np.float32)
    return s


def _golden_groups():
    """golden cases grouped by shape so that each group is solved as ONE batch on the device"""
    groups = {}
    for i in range(N_CASES):
        st = G[f"c{i}_state"]
        groups.setdefault((int(G[f"c{i}_nodes"]), st.shape[0]), []).append(i)
    return groups


def test_dominant_equals_reference_placements():
    from tpc_b200.heuristics import DominantResourceHeuristic
    for (n_nodes, _), idxs in _golden_groups().items():
        batch = np.stack([G[f"c{i}_state"] for i in idxs])
        for rd in (1, 0):
            for nd in (1, 0):
                sol = DominantResourceHeuristic(n_nodes, 100, {"resource_sort_descending": bool(rd),
                                                               "node_sort_descending": bool(nd)}, DEV)
                sol.solve(batch)
                a, r = sol.solution.assign.cpu().numpy(), sol.solution.remaining.cpu().numpy()
                dom, rej, emp = (t.cpu().numpy() for t in sol.stats())
                for k, i in enumerate(idxs):
                    np.testing.assert_array_equal(a[k], G[f"c{i}_dom_{rd}{nd}_assign"])
                    np.testing.assert_array_equal(r[k, 1:], G[f"c{i}_dom_{rd}{nd}_rem"][1:])
                    ref = G[f"c{i}_dom_{rd}{nd}_stats"]
                    assert (float(dom[k]), int(rej[k]), int(emp[k])) == (float(np.float32(ref[0])), int(ref[1]), int(ref[2]))
    assert sol.name == "dominant_resource_ASC_False_node_ASC_False"


def test_reference_unit_test_instance():
    """tests/unit/environment/resource_v3/heuristic_test.py:166-199 (un-normalised integers)."""
    from tpc_b200.heuristics import DominantResourceHeuristic
    sol = DominantResourceHeuristic(3, 100, {"resource_sort_descending": True, "node_sort_descending": True}, DEV)
    sol.solve(G["kat_state"][None])
    np.testing.assert_array_equal(sol.solution.assign.cpu().numpy()[0], [2, 0])
    np.testing.assert_array_equal(sol.solution.remaining.cpu().numpy()[0, 1:], G["kat_rem"][1:])


@pytest.mark.parametrize("shape", [(64, 6, 10, 0, 100, 3), (48, 11, 20, 0, 30, 3), (32, 51, 100, 0, 100, 4),
                                   (16, 40, 130, 0, 60, 3), (5, 2, 1, 0, 100, 3)])
def test_all_heuristics_equal_oracle(shape):
    from tpc_b200.heuristics import heuristic_factory
    B, n_nodes, n_rules, lo, hi, F = shape
    rng = np.random.default_rng(hash(shape) & 0xffff)
    batch = _instances(rng, B, n_nodes, n_rules, lo, hi, F)
    opts = {"dominant_resource": {"generate_params_combos": True, "resource_sort_descending": True,
                                  "node_sort_descending": True}, "random": {},
            "cplex_greedy_and_critical": {"use": False}, "cplex_node_reduction": {"use": False}}
    solvers = heuristic_factory(n_nodes, 100, opts, DEV, seed=99, instance_id0=1000)
    assert [s.name for s in solvers[:4]] == [f"dominant_resource_ASC_{r}_node_ASC_{n}" for r in (True, False)
                                             for n in (True, False)] and solvers[4].name == "random"
    for rep in range(2):          # the random heuristic moves to a new stream on every solve()
        for s in solvers:
            s.solve(batch)
        for s in solvers:
            a, r = s.solution.assign.cpu().numpy(), s.solution.remaining.cpu().numpy()
            dom, rej, emp = (t.cpu().numpy() for t in s.stats())
            for b in range(B):
                if s.name == "random":
                    oa, orem = oh.random_heuristic(batch[b, :, :3], n_nodes, oh.philox_randrange(99, 1000 + b, rep))
                else:
                    oa, orem = oh.dominant_heuristic(batch[b, :, :3], n_nodes, s.resource_sort_descending,
                                                     s.node_sort_descending)
                np.testing.assert_array_equal(a[b], oa)
                np.testing.assert_array_equal(r[b, 1:], orem[1:])
                od, orj, oe = oh.solution_stats(oa, orem)
                assert (dom[b], rej[b], emp[b]) == (od, orj, oe)


def test_compare_solutions_equals_oracle():
    import torch
    from tpc_b200.heuristics import compare_solutions
    rng = np.random.default_rng(5)
    B, H = 777, 5
    nd = (rng.integers(-5, 40, B) / 100 + rng.integers(0, 2, B) * 0.004999).astype(np.float32)
    nr = rng.integers(0, 6, B).astype(np.int32)
    hd = (rng.integers(-5, 40, (H, B)) / 100).astype(np.float32)
    hr = rng.integers(0, 6, (H, B)).astype(np.int32)
    t = lambda x: torch.as_tensor(x).to(DEV)
    rounded, res = compare_solutions(t(nd), t(nr), [t(x) for x in hd], [t(x) for x in hr])
    exp = np.zeros(6, dtype=np.int64)
    exp_round = np.zeros(B, dtype=np.float32)
    for b in range(B):
        exp_round[b], cd, cr = oh.compare_solutions(nd[b], int(nr[b]), hd[:, b], hr[:, b])
        exp[cd] += 1
        exp[3 + cr] += 1
    np.testing.assert_array_equal(rounded.cpu().numpy(), exp_round)
    np.testing.assert_array_equal(res.cpu().numpy(), exp)
    # no heuristics at all: the net wins against the reference's -1 / 9999 starting values
    _, res0 = compare_solutions(t(np.abs(nd)), t(nr), [], [])
    np.testing.assert_array_equal(res0.cpu().numpy(), [B, 0, 0, B, 0, 0])


def test_net_stats_and_tester_sweep_full_batch(tmp_path):
    """BASELINE config #2 shape (batch 1024, 50 nodes x 100 rules, greedy decode): statistics of the net's rollout
    against an independent host recount, heuristics are never worse than rejecting everything, CSV layout."""
    import torch
    from tpc_b200.agent import Agent
    from tpc_b200.configs import get_configs
    from tpc_b200.env import ResourceEnvironmentV3
    from tpc_b200.heuristics import net_solution_stats
    from tpc_b200.tester import test as tester
    agent_cfg, _, env_cfg, tester_cfg, _, _ = get_configs("ResourceV3", "tpc")
    env_cfg = json.loads(json.dumps(env_cfg))
    env_cfg.update(batch_size=1024, node_sample_size=50, profiles_sample_size=100)
    env = ResourceEnvironmentV3("ResourceV3", env_cfg, DEV)
    env.add_stats_to_agent_config(agent_cfg)
    agent = Agent("transformer", agent_cfg, DEV, num_features=3, seed=4)
    bufs = env.device_buffers(store=False)
    env.reset_into(bufs)
    init = bufs.batch.cpu().numpy().copy()
    agent.rollout(env, bufs, greedy=True)
    dom, rej, emp = (t.cpu().numpy() for t in net_solution_stats(bufs, 51))
    acts = bufs.actions.cpu().numpy().T                    # (B,T)
    final = bufs.batch.cpu().numpy()
    for b in range(0, 1024, 37):
        od, orj, oe = oh.solution_stats(acts[b], final[b, :51])
        assert (dom[b], rej[b], emp[b]) == (od, orj, oe)
    # replaying the net's picks through the heuristic oracle's Node arithmetic gives the env's final state (the
    # reference builds its history this way, env.py:354-384)
    for b in (0, 511, 1023):
        rem = init[b, :51].copy()
        for t in range(100):
            n = acts[b, t]
            if n:
                rem[n] = oh.rhu2(rem[n] - init[b, 51 + t])
        np.testing.assert_array_equal(rem[1:], final[b, 1:51])

    tcfg = json.loads(json.dumps(tester_cfg))
    tcfg.update(batch_size=1024, show_per_test_stats=False, show_inference_progress=False)
    tcfg["testbed"].update(num_tests=1, node_sample_configs={"min": 10, "max": 50, "step": 40},
                           request_sample_configs={"min": 20, "max": 100, "step": 80})
    dres, rres = tester(env, agent, tcfg, str(tmp_path))
    assert dres.shape == rres.shape == (3,) and dres.sum() == rres.sum() == 4 * 1024
    lines = open(os.path.join(str(tmp_path), "tests", "test.csv")).read().splitlines()
    assert lines[0].startswith("test_instance;node_sample_size;node_min_value;node_max_value;resource_sample_size;"
                               "net dominant;net rejected;net empty nodes;net optimal;"
                               "dominant resource ASC True node ASC True dominant;")
    assert lines[0].endswith("random dominant;random rejected;random empty nodes;random optimal;")
    assert len(lines) == 1 + 4 * 1024 and all(len(l.split(";")) == 5 + 6 * 4 + 1 for l in lines)
    # an untrained net loses to the best-fit heuristics on most instances; verdict sums are per instance
    stats = tester.global_stats[-1]["instance"]
    best_rej = np.min([list(s.values())[1] for s in stats[1:]], axis=0)
    net_rej = stats[0]["net_rejected"]
    assert rres[2] >= 0 and (best_rej <= 100).all() and (net_rej <= 100).all()


# ---------------------------------------------------------------------------------------------------------------
# KnapsackV2
# ---------------------------------------------------------------------------------------------------------------
KG = np.load(os.path.join(os.path.dirname(__file__), "golden", "heuristic_knapsack_v2.npz"))


def _kn_instances(rng, B, n_bins, n_items, lo, hi):
    s = np.zeros((B, n_bins + n_items, 2), dtype=np.float32)
    s[:, 0] = -2
    s[:, 1:n_bins, 0] = (rng.integers(lo, hi, size=(B, n_bins - 1)) / 100).astype(np.float32)
    s[:, n_bins:, 0] = (rng.integers(1, 30, size=(B, n_items)) / 100).astype(np.float32)
    s[:, n_bins:, 1] = (rng.integers(1, 100, size=(B, n_items)) / 100).astype(np.float32)
    return s


def test_knapsack_waste_reduction_equals_reference_placements():
    from tpc_b200.heuristics import WasteReductionHeuristic
    groups = {}
    for i in range(int(KG["num_cases"])):
        groups.setdefault((int(KG[f"c{i}_bins"]), KG[f"c{i}_state"].shape[0]), []).append(i)
    for (n_bins, _), idxs in groups.items():
        batch = np.stack([KG[f"c{i}_state"] for i in idxs])
        for i_desc in (1, 0):
            for b_desc in (1, 0):
                sol = WasteReductionHeuristic(n_bins, {"item_sort_descending": bool(i_desc),
                                                       "bin_sort_descending": bool(b_desc)}, DEV)
                sol.solve(batch)
                s = sol.solution
                rw, emp, rej, rv = (t.cpu().numpy() for t in sol.stats())
                for k, i in enumerate(idxs):
                    tag = f"c{i}_wr_{i_desc}{b_desc}"
                    np.testing.assert_array_equal(s.assign.cpu().numpy()[k], KG[tag + "_assign"])
                    np.testing.assert_array_equal(s.bins.cpu().numpy()[k, 1:], KG[tag + "_bins"][1:])
                    np.testing.assert_array_equal(s.values.cpu().numpy()[k], KG[tag + "_values"])
                    ref = KG[tag + "_stats"]
                    assert (float(rw[k]), int(emp[k]), int(rej[k]), float(rv[k])) == \
                        (float(np.float32(ref[0])), int(ref[1]), int(ref[2]), float(np.float32(ref[3])))
    assert sol.name == "item_ASC_False_bin_ASC_False"


@pytest.mark.parametrize("shape", [(40, 6, 10, 10, 100), (24, 11, 20, 10, 40), (16, 51, 100, 10, 100), (7, 2, 3, 5, 60)])
def test_knapsack_heuristics_equal_oracle(shape):
    import torch
    from tpc_b200.heuristics import knapsack_heuristic_factory, knapsack_stats
    from tpc_b200 import lib as L
    from tpc_b200.heuristics import _handle, _stream
    B, n_bins, n_items, lo, hi = shape
    batch = _kn_instances(np.random.default_rng(sum(shape)), B, n_bins, n_items, lo, hi)
    opts = {"waste_reduction": {"generate_params_combos": True, "item_sort_descending": True,
                                "bin_sort_descending": False}, "random": {}, "or_tools": {}}
    solvers = knapsack_heuristic_factory(n_bins, opts, DEV, seed=7, instance_id0=500)
    assert [s.name for s in solvers] == [f"item_ASC_{i}_bin_ASC_{b}" for i in (True, False) for b in (True, False)] + ["random"]
    for rep in range(2):
        for s in solvers:
            s.solve(batch)
        for s in solvers:
            sol = s.solution
            a, o, bn, vl = (t.cpu().numpy() for t in (sol.assign, sol.order, sol.bins, sol.values))
            rw, emp, rej, rv = (t.cpu().numpy() for t in s.stats())
            for b in range(B):
                if s.name == "random":
                    oa, oo, ob, ov = oh.knapsack_random(batch[b], n_bins, oh.philox_randrange(7, 500 + b, rep))
                else:
                    oa, oo, ob, ov = oh.knapsack_waste_reduction(batch[b], n_bins, s.item_sort_descending,
                                                                 s.bin_sort_descending)
                np.testing.assert_array_equal(a[b], oa)
                np.testing.assert_array_equal(o[b], oo)
                np.testing.assert_array_equal(bn[b, 1:], ob[1:])
                np.testing.assert_array_equal(vl[b], ov)
                assert (rw[b], emp[b], rej[b], rv[b]) == oh.knapsack_solution_stats(oa, ov)
    # per-bin values of a placement made in item order (the net's history): step-major actions (T, B)
    acts = torch.as_tensor(np.random.default_rng(1).integers(0, n_bins, size=(n_items, B)).astype(np.int32)).to(DEV)
    st = torch.as_tensor(batch).to(DEV)
    lib, h = _handle(torch.device(DEV))
    values = torch.empty(B, n_bins, dtype=torch.float32, device=DEV)
    L.check(lib.tpc_knapsack_bin_values(h, L.ptr(acts), 1, acts.stride(0), L.ptr(st), st.stride(0), st.stride(1), B,
                                        n_bins, n_items, L.ptr(values), _stream(torch.device(DEV))))
    rw, emp, rej, rv = (t.cpu().numpy() for t in knapsack_stats(acts, 1, acts.stride(0), values))
    for b in range(B):
        ov = oh.knapsack_bin_values(acts[:, b].cpu().numpy(), batch[b], n_bins)
        np.testing.assert_array_equal(values[b].cpu().numpy(), ov)
        assert (rw[b], emp[b], rej[b], rv[b]) == oh.knapsack_solution_stats(acts[:, b].cpu().numpy(), ov)


def test_knapsack_tester_sweep(tmp_path):
    """configs/KnapsackV2.json test bed (reduced), batch 64: verdict counts, CSV layout, reward == episode return."""
    import torch
    from tpc_b200.agent import Agent
    from tpc_b200.configs import get_configs
    from tpc_b200.env import env_factory
    agent_cfg, _, env_cfg, tester_cfg, _, _ = get_configs("KnapsackV2", "tpc")
    env_cfg = json.loads(json.dumps(env_cfg))
    env_cfg.update(batch_size=64)
    env, tester = env_factory("custom", "KnapsackV2", env_cfg, DEV)
    env.add_stats_to_agent_config(agent_cfg)
    agent = Agent("transformer", agent_cfg, DEV, num_features=2, seed=9)
    tcfg = json.loads(json.dumps(tester_cfg))
    tcfg.update(batch_size=64)
    tcfg["testbed"].update(num_tests=1, bin_sample_configs={"min": 5, "max": 10, "step": 5},
                           item_sample_configs={"min": 10, "max": 20, "step": 10})
    res = tester(env, agent, tcfg, str(tmp_path))
    assert res.shape == (3,) and res.sum() == 4 * 64
    lines = open(os.path.join(str(tmp_path), "tests", "test.csv")).read().splitlines()
    assert lines[0].startswith("test_instance;bin_sample_size;bin_min_value;bin_max_value;item_sample_size;"
                               "net reward;net empty bins;net num rejected items;net rejected value;"
                               "item ASC True bin ASC True reward;")
    assert len(lines) == 1 + 4 * 64 and all(len(l.split(";")) == 5 + 6 * 4 + 1 for l in lines)
    # the net's reward statistic equals the sum of the environment's step rewards (value of every placed item)
    env.set_testing_mode(64, 10, 20, 0, 100)
    agent.set_testing_mode(64, env.node_sample_size, env.profiles_sample_size)
    from tpc_b200.heuristics import knapsack_net_stats
    bufs = env.device_buffers(store=False)
    env.reset_into(bufs)
    init = bufs.batch.clone()
    agent.rollout(env, bufs, greedy=True)
    (rw, _, rej, rv), values = knapsack_net_stats(bufs, init, env.node_sample_size)
    total = bufs.rewards.sum(dim=1)
    assert torch.allclose(rw, total, atol=1e-5) and torch.all((rej == 0) == (rv == 0))


def test_reference_heuristic_unit_test_instances():
    """The solver known-answer tests the reference holds: resource_v3/heuristic_test.py:196-236 (RandomHeuristic on
    the un-normalised instance: whatever the draws, delta 1, one rejected rule, one empty node) and
    knapsack_v2/heuristic_test.py:236-300, 303-410 (WasteReduction and Random on the 2-bin / 4-item instance)."""
    from tpc_b200.heuristics import KnapsackRandomHeuristic, RandomHeuristic, WasteReductionHeuristic
    for seed in range(8):
        sol = RandomHeuristic(3, 100, {}, DEV, seed=seed)
        sol.solve(G["kat_state"][None])
        dom, rej, emp = (int(t.cpu().numpy()[0]) for t in sol.stats())
        assert (dom, rej, emp) == (1, 1, 1)
    kat = np.array([[-2.0, -2.0], [0.1, 0.0], [0.5, 0.0], [0.2, 0.1], [0.3, 0.5], [0.1, 0.4], [0.9, 0.4]], dtype=np.float32)
    sol = WasteReductionHeuristic(3, {"item_sort_descending": True, "bin_sort_descending": False}, DEV)
    sol.solve(kat[None])
    rw, emp, rej, rv = (t.cpu().numpy()[0] for t in sol.stats())
    a = sol.solution.assign.cpu().numpy()[0]
    assert round(float(rw), 2) == 1.0 and (int(emp), int(rej)) == (0, 1) and rv == np.float32(0.4)
    assert [int((a == n).sum()) for n in range(3)] == [1, 1, 2]
    for seed in range(8):      # random: item 3 (weight 0.9) never fits, every other item fits some bin or is rejected
        sol = KnapsackRandomHeuristic(3, {}, DEV, seed=seed)
        sol.solve(kat[None])
        a = sol.solution.assign.cpu().numpy()[0]
        bins = sol.solution.bins.cpu().numpy()[0]
        assert a[3] == 0 and np.all(bins[1:, 1] <= bins[1:, 0])
