"""Measurement of the tester rows (BASELINE config #2): greedy decode of batch 1024 at the
largest test-bed cell (50 nodes x 100 rules), the five batched heuristics and the statistics / verdict kernels,
timed with CUDA events; the CPU oracle of the heuristics (the reference's per-instance Python algorithm restated
in NumPy) is timed beside them on a bounded sample. Prints one JSON line.

    python tests/gpu_checks/bench_tester.py [--batch 1024] [--reps 5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import heuristics as oh  # noqa: E402
from tpc_b200.agent import Agent  # noqa: E402
from tpc_b200.configs import get_configs  # noqa: E402
from tpc_b200.env import ResourceEnvironmentV3  # noqa: E402
from tpc_b200.heuristics import compare_solutions, heuristic_factory, net_solution_stats  # noqa: E402
from tpc_b200.lib import load  # noqa: E402


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu-instances", type=int, default=8)
    args = ap.parse_args()
    dev = "cuda:0"
    agent_cfg, _, env_cfg, tester_cfg, _, _ = get_configs("ResourceV3", "tpc")
    env_cfg = json.loads(json.dumps(env_cfg))
    B, nodes, rules = args.batch, 50, 100
    env_cfg.update(batch_size=B, node_sample_size=nodes, profiles_sample_size=rules)
    env = ResourceEnvironmentV3("ResourceV3", env_cfg, dev)
    env.add_stats_to_agent_config(agent_cfg)
    agent = Agent("transformer", agent_cfg, dev, num_features=3, seed=1)
    env.set_testing_mode(B, nodes, rules, 0, 100)
    agent.set_testing_mode(B, env.node_sample_size, env.profiles_sample_size)
    N = env.node_sample_size
    bufs = env.device_buffers(store=False)
    lib = load()
    env.reset_into(bufs)
    init = bufs.batch.clone()

    def rollout():
        env.reset_into(bufs)
        agent.rollout(env, bufs, greedy=True)
    n0 = lib.tpc_launch_count()
    ms_rollout = timed(rollout, args.reps)
    launches = (lib.tpc_launch_count() - n0) // (args.reps + 1)
    solvers = heuristic_factory(N, 100, tester_cfg["heuristic"], dev, seed=1235, instance_id0=0)

    def heur():
        for s in solvers:
            s.solve(init)
    ms_heur = timed(heur, args.reps)

    def stats():
        nd, nr, _ = net_solution_stats(bufs, N)
        hs = [s.stats() for s in solvers]
        compare_solutions(nd, nr, [h[0] for h in hs], [h[1] for h in hs])
    ms_stats = timed(stats, args.reps)

    # CPU: the oracle's per-instance algorithm on a few instances of the same batch
    host = init.cpu().numpy()
    t0 = time.perf_counter()
    for b in range(args.cpu_instances):
        for rd in (True, False):
            for nd_ in (True, False):
                oh.dominant_heuristic(host[b], N, rd, nd_)
        oh.random_heuristic(host[b], N, oh.philox_randrange(1235, b, 0))
    cpu_s_per_instance = (time.perf_counter() - t0) / args.cpu_instances
    out = {
        "workload": f"ResourceV3 tester cell {nodes} nodes x {rules} rules, batch {B}, greedy decode + 4 dominant-resource "
                    f"variants + random heuristic + statistics / verdicts",
        "rollout_ms": ms_rollout, "rollout_decisions_per_s": B * rules / ms_rollout * 1e3, "rollout_launches": int(launches),
        "heuristics_ms": ms_heur, "heuristic_instances_per_s": 5 * B / ms_heur * 1e3,
        "stats_ms": ms_stats,
        "cpu_oracle_heuristics_s_per_instance": cpu_s_per_instance,
        "cpu_oracle_heuristic_instances_per_s": 5 / cpu_s_per_instance,
        "cpu": "oracle/heuristics.py (NumPy restatement of the reference's per-instance Python heuristics), 1 thread, "
               f"{args.cpu_instances} instances",
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
