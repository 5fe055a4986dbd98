"""Test sweep with the reference's signature (environment/custom/resource_v3/tester.py:19-115): for every
(rules, nodes) cell of the test bed run `num_tests` greedy rollouts on device and gather the per-instance
statistics the reference derives from its Python history objects (misc/utils.py:56-81): rejected rules,
dominant (minimum remaining) resource, empty nodes. Statistics are computed for EVERY instance of the batch
(the reference only scores instance 0). Heuristic / CPLEX baselines are out of scope.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch


def rollout_stats(bufs, num_bins: int):
    """Device statistics of a finished greedy rollout -> dict of host arrays, one entry per instance."""
    acts = bufs.actions                      # (T,B)
    rejected = (acts == 0).sum(dim=0)        # rules sent to the EOS node (misc/utils.py:58-59)
    nodes = bufs.batch[:, 1:num_bins, :]     # remaining resources of the real nodes
    dominant = nodes.reshape(nodes.shape[0], -1).min(dim=1).values  # misc/utils.py:69-81
    B = acts.shape[1]
    used = torch.zeros(B, num_bins, dtype=torch.bool, device=acts.device)
    used.scatter_(1, acts.t().long(), True)
    empty = (num_bins - 1) - used[:, 1:].sum(dim=1)                 # nodes without any rule (misc/utils.py:66-67)
    reward = bufs.rewards.sum(dim=1)
    return {"rejected": rejected.cpu().numpy(), "dominant": np.round(dominant.cpu().numpy().astype(np.float64), 2),
            "empty_nodes": empty.cpu().numpy(), "reward": reward.cpu().numpy()}


def test_single_instance(env, agent, batch_size, node_size, resource_size, node_min, node_max):
    """tester.py:117-255 (rollout part): greedy decode of one batch of fresh instances."""
    env.set_testing_mode(batch_size, node_size, resource_size, node_min, node_max)
    agent.set_testing_mode(batch_size, env.node_sample_size, env.profiles_sample_size)
    bufs = env.device_buffers(store=False)
    env.reset_into(bufs)
    start = time.time()
    agent.rollout(env, bufs, greedy=True)
    torch.cuda.synchronize()
    return rollout_stats(bufs, env.node_sample_size), time.time() - start


def test(env, agent, opts: dict, log_dir: str):
    tb = opts["testbed"]
    batch_size = opts.get("batch_size", 1)
    num_tests = tb["num_tests"]
    nodes_cfg, req_cfg, res_cfg = tb["node_sample_configs"], tb["request_sample_configs"], tb["node_available_resources"]
    rows = []
    for resource_size in range(req_cfg["min"], req_cfg["max"] + 1, req_cfg["step"]):
        for node_size in range(nodes_cfg["min"], nodes_cfg["max"] + 1, nodes_cfg["step"]):
            for node_min in range(res_cfg["min"], res_cfg["max"], res_cfg["step"]):
                node_max = node_min + res_cfg["step"]
                for _ in range(num_tests):
                    stats, took = test_single_instance(env, agent, batch_size, node_size, resource_size, node_min, node_max)
                    for b in range(batch_size):
                        rows.append((node_size, resource_size, node_min, node_max, stats["dominant"][b],
                                     int(stats["rejected"][b]), int(stats["empty_nodes"][b]), took))
                if opts.get("show_per_test_stats"):
                    last = rows[-batch_size * num_tests:]
                    print(f"nodes {node_size:3d} rules {resource_size:3d}: mean rejected "
                          f"{np.mean([r[5] for r in last]):.3f} mean empty {np.mean([r[6] for r in last]):.3f}")
    gs = opts["export_stats"]["global_stats"]
    if gs["export_stats"] and log_dir is not None:
        folder = os.path.join(log_dir, gs["folder"])
        os.makedirs(folder, exist_ok=True)
        with open(os.path.join(folder, gs["filename"] + ".csv"), "w") as f:
            # ';'-separated like environment/custom/resource_v3/misc/csv_writer.py:3-35 (net columns only)
            f.write("node_sample_size;resource_sample_size;node_min_value;node_max_value;net_dominant;net_rejected;"
                    "net_empty_nodes;time\n")
            for r in rows:
                f.write(";".join(str(x) for x in r) + "\n")
    dominant = np.array([r[4] for r in rows])
    rejected = np.array([r[5] for r in rows])
    return dominant, rejected
