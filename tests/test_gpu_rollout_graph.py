"""The rollout's step loop replayed as a CUDA graph (tpc.h: tpc_rollout_graph_replays; net.cu rollout_impl) must be
indistinguishable from plain stream launches: same actions, rewards, masks and stored states, bit for bit, for sampled
rollouts too (the Philox draw index is the only argument that changes from call to call and is read from device memory)."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu


def _rollouts(graph: bool, n_calls: int, batch: int, nodes: int, rules: int, greedy: bool, fork: int = 1):
    import torch
    from tpc_b200.agent import Agent
    from tpc_b200.configs import get_configs
    from tpc_b200.env import ResourceEnvironmentV3
    os.environ["TPC_ROLLOUT_GRAPH"] = "1" if graph else "0"
    os.environ["TPC_ROLLOUT_FORK"] = str(fork)
    try:
        agent_cfg, _, env_cfg, _, _, _ = get_configs("ResourceV3", "tpc")
        env_cfg = json.loads(json.dumps(env_cfg))
        env_cfg.update(batch_size=batch, node_sample_size=nodes, profiles_sample_size=rules)
        env = ResourceEnvironmentV3("ResourceV3", env_cfg, "cuda:0")
        env.add_stats_to_agent_config(agent_cfg)
        agent = Agent("transformer", agent_cfg, "cuda:0", num_features=3, seed=3)
        bufs = env.device_buffers(store=True)
        out = []
        for _ in range(n_calls):
            env.reset_into(bufs)          # a new instance batch per call (env.epoch advances), same buffers
            agent.rollout(env, bufs, greedy=greedy)
            torch.cuda.synchronize()
            out.append({k: getattr(bufs, k).clone() for k in
                        ("actions", "rewards", "states", "ptr_masks", "mha_masks", "dec_inputs", "batch", "bin_net_mask")})
        eng = agent.engine
        replays = int(eng.lib.tpc_rollout_graph_replays(eng.m))
        return out, replays
    finally:
        os.environ.pop("TPC_ROLLOUT_GRAPH", None)
        os.environ.pop("TPC_ROLLOUT_FORK", None)


@pytest.mark.parametrize("batch,nodes,rules,greedy", [(16, 5, 9, False), (128, 10, 20, False), (64, 50, 100, True)])
def test_graph_replay_equals_stream_launches_bit_for_bit(batch, nodes, rules, greedy):
    import torch
    n_calls = 4
    ref, r0 = _rollouts(False, n_calls, batch, nodes, rules, greedy, fork=0)   # one stream, plain launches
    got, r1 = _rollouts(True, n_calls, batch, nodes, rules, greedy, fork=2)    # graph replay with the two-stream fork
    assert r0 == 0
    # call 1 runs on the stream, call 2 captures and replays, 3.. replay -- unless a CUDA injection tool (profiler,
    # tracer, sanitizer) is attached: the library then keeps stream launches (net.cu rollout_graph_wanted)
    traced = any(k in os.environ for k in ("CUDA_INJECTION64_PATH", "NV_COMPUTE_PROFILER_PERFWORKS_DIR"))
    assert r1 == (0 if traced else n_calls - 1), r1
    for i, (a, b) in enumerate(zip(ref, got)):
        for k in a:
            assert torch.equal(a[k], b[k]), (i, k)
    if not greedy:   # sampled rollouts of different epochs must differ (the draw index really advances under replay)
        assert not torch.equal(got[1]["actions"], got[2]["actions"]) or not torch.equal(got[1]["batch"], got[2]["batch"])


def test_two_stream_encoder_fork_equals_one_stream_bit_for_bit():
    """encoder_forward's fork / join (rules stack on a second stream beside the nodes stack) without a graph."""
    import torch
    ref, _ = _rollouts(False, 2, 32, 10, 20, False, fork=0)
    got, _ = _rollouts(False, 2, 32, 10, 20, False, fork=2)
    for a, b in zip(ref, got):
        for k in a:
            assert torch.equal(a[k], b[k]), k
