"""Error behaviour and edge cases of the C-ABI (include/tpc.h): every entry point returns 0 or a negative TPC_ERR_*
code and never throws; empty batches are no-ops; shapes outside a kernel's limits are refused, not mis-computed."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
OK, ERR_ARG, ERR_SHAPE = 0, -2, -3


@pytest.fixture(scope="module")
def ctx():
    import torch
    from tpc_b200 import lib as L
    lib = L.load()
    h = C.c_void_p()
    assert lib.tpc_create(C.byref(h)) == OK
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    yield lib, h, st, L, torch
    lib.tpc_destroy(h)


def test_null_arguments_are_refused(ctx):
    lib, h, st, L, torch = ctx
    x = torch.zeros(4, 8, 3, device=DEV)
    a = torch.zeros(4, 5, dtype=torch.int32, device=DEV)
    r = torch.zeros(4, 3, 3, device=DEV)
    assert lib.tpc_heuristic_dominant(None, L.ptr(x), 24, 3, 4, 3, 5, 1, 1, L.ptr(a), L.ptr(r), st) == ERR_ARG
    assert lib.tpc_heuristic_dominant(h, None, 24, 3, 4, 3, 5, 1, 1, L.ptr(a), L.ptr(r), st) == ERR_ARG
    assert lib.tpc_heuristic_random(h, L.ptr(x), 24, 3, 4, 3, 5, 1, 0, 0, None, L.ptr(r), st) == ERR_ARG
    assert lib.tpc_solution_stats(h, L.ptr(a), 5, 1, L.ptr(r), 9, 3, 4, 3, 5, None, None, None, st) == ERR_ARG
    assert lib.tpc_compare_solutions(h, None, None, None, None, 0, 4, None, None, st) == ERR_ARG
    assert lib.tpc_gemm(None, None, 0, None, 0, None, 0, 1, 128, 128, None, None, 0, 0, None, None, None, 0, 1, None, st) == ERR_ARG
    assert lib.tpc_attention_fwd(h, None, None, 8, 0, 1, 8, None, None, st) == ERR_ARG
    assert lib.tpc_destroy(None) == OK          # destroying nothing is fine


def test_empty_batches_are_no_ops(ctx):
    lib, h, st, L, torch = ctx
    x = torch.zeros(1, 8, 3, device=DEV)
    a = torch.full((1, 5), 7, dtype=torch.int32, device=DEV)
    r = torch.full((1, 3, 3), 7.0, device=DEV)
    assert lib.tpc_heuristic_dominant(h, L.ptr(x), 24, 3, 0, 3, 5, 1, 1, L.ptr(a), L.ptr(r), st) == OK
    assert lib.tpc_heuristic_random(h, L.ptr(x), 24, 3, 0, 3, 5, 1, 0, 0, L.ptr(a), L.ptr(r), st) == OK
    d = torch.zeros(1, device=DEV)
    i = torch.zeros(1, dtype=torch.int32, device=DEV)
    assert lib.tpc_solution_stats(h, L.ptr(a), 5, 1, L.ptr(r), 9, 3, 0, 3, 5, L.ptr(d), L.ptr(i), L.ptr(i), st) == OK
    res = torch.zeros(6, dtype=torch.int32, device=DEV)
    assert lib.tpc_compare_solutions(h, L.ptr(d), L.ptr(i), None, None, 0, 0, None, L.ptr(res), st) == OK
    torch.cuda.synchronize()
    assert int(a.min()) == 7 and float(r.min()) == 7.0 and int(res.sum()) == 0     # nothing was touched
    # GEMM with no rows
    A = torch.zeros(1, 128, device=DEV)
    assert lib.tpc_gemm(h, L.ptr(A), 128, L.ptr(A), 128, L.ptr(A), 128, 0, 128, 128, None, None, 0, 0, None, None, None,
                        0, 1, None, st) == OK


def test_unsupported_shapes_are_refused(ctx):
    lib, h, st, L, torch = ctx
    A = torch.zeros(256, 160, device=DEV)
    W = torch.zeros(128, 160, device=DEV)
    D = torch.zeros(256, 128, device=DEV)
    # K must be a multiple of 32, N a multiple of 128; LayerNorm epilogue needs N == 128; split-K needs the atomic epilogue
    assert lib.tpc_gemm(h, L.ptr(A), 160, L.ptr(W), 160, L.ptr(D), 128, 256, 128, 144, None, None, 0, 0, None, None, None,
                        0, 1, None, st) == ERR_SHAPE
    assert lib.tpc_gemm(h, L.ptr(A), 160, L.ptr(W), 160, L.ptr(D), 128, 256, 96, 160, None, None, 0, 0, None, None, None,
                        0, 1, None, st) == ERR_SHAPE
    assert lib.tpc_gemm(h, L.ptr(A), 160, L.ptr(W), 160, L.ptr(D), 128, 256, 128, 160, None, None, 0, 0, None, None, None,
                        0, 4, None, st) == ERR_ARG
    # heuristics: at least one real node and one rule; shared-memory budget bounds the instance size
    x = torch.zeros(2, 6, 3, device=DEV)
    a = torch.zeros(2, 5, dtype=torch.int32, device=DEV)
    r = torch.zeros(2, 1, 3, device=DEV)
    assert lib.tpc_heuristic_dominant(h, L.ptr(x), 18, 3, 2, 1, 5, 1, 1, L.ptr(a), L.ptr(r), st) == ERR_SHAPE
    assert lib.tpc_heuristic_dominant(h, L.ptr(x), 18, 3, 2, 3, 0, 1, 1, L.ptr(a), L.ptr(r), st) == ERR_SHAPE
    assert lib.tpc_heuristic_dominant(h, L.ptr(x), 18, 3, 2, 3000, 9000, 1, 1, L.ptr(a), L.ptr(r), st) == ERR_SHAPE


def test_python_mirror_raises_on_bad_input():
    import torch
    from tpc_b200.heuristics import DominantResourceHeuristic, heuristic_factory
    from tpc_b200.lib import TpcError
    sol = DominantResourceHeuristic(3, 100, {"resource_sort_descending": True, "node_sort_descending": True}, DEV)
    with pytest.raises(ValueError):
        sol.solve(np.zeros((2, 6, 2), dtype=np.float32))            # needs 3 features
    with pytest.raises(ValueError):
        sol.solve(np.zeros((2, 3, 3), dtype=np.float32))            # no rules behind the 3 nodes
    with pytest.raises(NotImplementedError):
        heuristic_factory(3, 100, {"dominant_resource": {"generate_params_combos": False, "resource_sort_descending": True,
                                                         "node_sort_descending": True}, "random": {},
                                   "cplex_greedy_and_critical": {"use": True}, "cplex_node_reduction": {"use": False}}, DEV)
    big = DominantResourceHeuristic(3000, 100, {"resource_sort_descending": True, "node_sort_descending": True}, DEV)
    with pytest.raises(TpcError):
        big.solve(torch.zeros(1, 12000, 3))                          # TPC_ERR_SHAPE surfaces as an exception


def test_ragged_instances_single_node_and_all_rejected(ctx):
    """Smallest legal instance (one real node, one rule) and an instance where nothing fits: every rule goes to the
    EOS node, the node stays untouched, statistics follow misc/utils.py:56-81."""
    from tpc_b200.heuristics import heuristic_factory
    from oracle import heuristics as oh
    opts = {"dominant_resource": {"generate_params_combos": True, "resource_sort_descending": True,
                                  "node_sort_descending": True}, "random": {},
            "cplex_greedy_and_critical": {"use": False}, "cplex_node_reduction": {"use": False}}
    tiny = np.array([[[-2, -2, -2], [0.5, 0.5, 0.5], [0.5, 0.5, 0.5]],
                     [[-2, -2, -2], [0.1, 0.1, 0.1], [0.2, 0.05, 0.05]]], dtype=np.float32)
    for s in heuristic_factory(2, 100, opts, DEV):
        s.solve(tiny)
        a = s.solution.assign.cpu().numpy()
        dom, rej, emp = (t.cpu().numpy() for t in s.stats())
        assert a.tolist() == [[1], [0]] and rej.tolist() == [0, 1] and emp.tolist() == [0, 1]
        assert dom[0] == np.float32(0.0) and dom[1] == np.float32(0.1)
    full = np.zeros((1, 3 + 4, 3), dtype=np.float32)
    full[0, 0] = -2
    full[0, 1:3] = 0.05
    full[0, 3:] = 0.2
    for s in heuristic_factory(3, 100, opts, DEV):
        s.solve(full)
        assert s.solution.assign.cpu().numpy().tolist() == [[0, 0, 0, 0]]
        np.testing.assert_array_equal(s.solution.remaining.cpu().numpy()[0, 1:], full[0, 1:3])
        oa, orem = oh.dominant_heuristic(full[0], 3, True, True)
        assert oa.tolist() == [0, 0, 0, 0]


# ---------------------------------------------------------------------------------------------------------------------
# model path: smallest / odd / largest shapes of the actor, the fused rollout and the update
# ---------------------------------------------------------------------------------------------------------------------
def _agent(N_nodes, R, B, reward="greedy"):
    import json
    from tpc_b200.agent import Agent
    from tpc_b200.configs import get_configs
    from tpc_b200.env import ResourceEnvironmentV3
    agent_cfg, _, env_cfg, _, _, _ = get_configs("ResourceV3", "tpc")
    env_cfg = json.loads(json.dumps(env_cfg))
    env_cfg.update(batch_size=B, node_sample_size=N_nodes, profiles_sample_size=R)
    env_cfg["reward"]["type"] = reward
    env = ResourceEnvironmentV3("ResourceV3", env_cfg, DEV)
    agent_cfg = env.add_stats_to_agent_config(json.loads(json.dumps(agent_cfg)))
    return env, Agent("transformer", agent_cfg, DEV, num_features=3, seed=7)


@pytest.mark.parametrize("nodes,rules,B", [(1, 1, 1), (1, 3, 3), (3, 2, 5), (10, 20, 7)])
def test_smallest_and_odd_shapes_train_one_iteration(nodes, rules, B):
    """One real node / one rule / batches that fill no CTA of the decoder chain (2 or 8 states per CTA) and no GEMM tile:
    the rollout replays bit-exactly through the oracle env, logits match the oracle model, the update stays finite."""
    import torch
    from oracle import model as om
    from oracle.env import ResourceV3Oracle
    env, agent = _agent(nodes, rules, B)
    N, R, S = nodes + 1, rules, nodes + 1 + rules
    bufs = env.device_buffers(store=True, keep_logits=True)
    env.reset_into(bufs)
    init = bufs.batch.cpu().numpy().copy()
    agent.rollout(env, bufs, greedy=True)
    torch.cuda.synchronize()
    cfg = dict(env.opts)
    ora = ResourceV3Oracle(cfg)
    state, dec, bm, mh = ora.load(init)
    fa = torch.tensor(agent.bin_actor.get_flat())
    ca = om.NetCfg(num_layers=1, dff=128)
    for t in range(R):
        feas = ora.build_feasible_mask(state, dec, bm)
        assert np.array_equal(bufs.states[t].cpu().numpy(), state) and np.array_equal(bufs.ptr_masks[t].cpu().numpy(), feas)
        ol, _, oi = om.actor_forward(fa, ca, torch.tensor(state), torch.tensor(dec), torch.tensor(feas), torch.tensor(mh), N)
        unm = feas == 0
        cl = bufs.logits[t].cpu().numpy()
        assert np.abs(cl[unm] - ol.numpy()[unm]).max() <= 1e-3 * max(1.0, np.abs(ol.numpy()[unm]).max())
        assert np.all(cl[~unm] == -1e9)
        a = bufs.actions[t].cpu().numpy()
        state, dec, rw, done, info = ora.step(a, feas)
        assert np.array_equal(bufs.rewards[:, t].cpu().numpy(), np.asarray(rw, np.float32).reshape(B))
        bm, mh = info["bin_net_mask"], info["mha_used_mask"]
    assert done
    losses = agent.update(bufs).cpu().numpy()
    assert np.all(np.isfinite(losses))
    assert np.all(np.isfinite(agent.bin_actor.get_flat())) and np.all(np.isfinite(agent.critic.get_flat()))


def test_largest_sequence_and_refusal_beyond_it():
    """S = 512 tokens is the limit of the one-query kernels (cross-attention, pointer head): it runs and matches the
    oracle; one token more is refused with TPC_ERR_SHAPE (surfaced as TpcError), not mis-computed."""
    import torch
    from oracle import model as om
    from tpc_b200.lib import TpcError
    rng = np.random.default_rng(3)
    for N, R, ok in ((112, 400, True), (113, 400, False)):
        env, agent = _agent(N - 1, R, 2)
        S = N + R
        state = rng.integers(0, 100, size=(2, S, 3)).astype(np.float32) / 100
        state[:, 0] = -2
        dec = state[:, N:N + 1].copy()
        pm = np.zeros((2, S), np.float32)
        pm[:, N:] = 1
        mh = np.zeros((2, 1, 1, S), np.float32)
        mh[:, :, :, N:N + 7] = 1
        if not ok:
            with pytest.raises(TpcError):
                agent.bin_actor(state, dec, pm, True, N, enc_padding_mask=mh, dec_padding_mask=mh)
            continue
        lg, pr, idx, _ = agent.bin_actor(state, dec, pm, True, N, enc_padding_mask=mh, dec_padding_mask=mh)
        ol, op, oi = om.actor_forward(torch.tensor(agent.bin_actor.get_flat()), om.NetCfg(num_layers=1, dff=128),
                                      torch.tensor(state), torch.tensor(dec), torch.tensor(pm), torch.tensor(mh), N)
        unm = pm == 0
        assert np.abs(np.asarray(lg)[unm] - ol.numpy()[unm]).max() < 1e-3 * np.abs(ol.numpy()[unm]).max()
        assert np.abs(np.asarray(pr) - op.numpy()).max() < 1e-3


def test_cvrp_smallest_and_empty(ctx):
    lib, h, st, L, torch = ctx
    from tpc_b200.env import CVRP
    env = CVRP("CVRP", {"batch_size": 3, "node_sample_size": 2, "vehicle_sample_size": 1, "num_features": 3})
    state, bp, item, mha = env.reset()
    assert state.shape == (3, 3, 3) and env.steps == 2
    feas = env.build_feasible_mask(state, state[:, 1], bp)
    assert feas.tolist() == [[1.0, 1.0, 0.0]] * 3
    state, r, done, info = env.step([2, 2, 2], [1, 1, 1])
    assert not done and np.all(np.asarray(r) <= 0) and np.all(info["item_net_mask"][:, 0] == 0)     # the depot opens
    state, r, done, info = env.step([2, 2, 2], [0, 0, 0])
    assert done and np.all(info["backpack_net_mask"] == 1)
    x = torch.zeros(1, 3, 3, device=DEV)
    m = torch.zeros(1, 3, device=DEV)
    i = torch.zeros(1, dtype=torch.int32, device=DEV)
    assert lib.tpc_cvrp_step(h, L.ptr(x), L.ptr(m), L.ptr(m), L.ptr(m), L.ptr(i), L.ptr(i), 0, 2, 1, 0, L.ptr(m), st) == OK
    assert lib.tpc_cvrp_reset(h, 1, 0, 0, 1, 1, 1, 24, 100, L.ptr(x), L.ptr(m), L.ptr(m), L.ptr(m), st) == ERR_SHAPE
    assert lib.tpc_cvrp_feasible_mask(h, None, L.ptr(x), L.ptr(m), 1, 3, L.ptr(m), st) == ERR_ARG
