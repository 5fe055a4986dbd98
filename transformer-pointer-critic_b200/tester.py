"""Test sweep with the reference's signature, return value and CSV layout
(environment/custom/resource_v3/tester.py:19-255, misc/utils.py:97-159, misc/csv_writer.py:3-35).

For every (rules, nodes, node-resource range) cell of the test bed: `num_tests` greedy rollouts of `batch_size`
fresh instances on the device, the dominant-resource (4 sort combinations) and random heuristics on the same
instances, per-instance statistics (dominant = minimum remaining resource, rejected rules, empty nodes) and the
net-vs-best-heuristic won / draw / lost verdicts — all computed by libtpc kernels for EVERY instance of the batch.
The reference scores instance 0 only (`env.history[0]`, misc/utils.py:107) and its heuristics assert batch == 1,
so a batch-1024 run fails there; with batch_size == 1 the two produce the same rows.
Per-problem CSV export, attention plots and solution printing (tester.py:228-253) need the reference's per-step
Python history objects and are not provided.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from .heuristics import compare_solutions, heuristic_factory, net_solution_stats


def generate_file_name(agent_config: dict):
    """misc/utils.py:161-170 (the reference builds the name but returns None; we return the string)."""
    a, c = agent_config["actor"], agent_config["critic"]
    return (f"num_l:{a['num_layers']}|e:{agent_config['entropy_coefficient']}|ac_lr:{a['learning_rate']}"
            f"|cr_lr:{c['learning_rate']}")


def test_single_instance(instance_id, env, agent, opts: dict, batch_size: int, node_sample_size: int,
                         node_min_val: int, node_max_val: int, req_sample_size: int, log_dir: str):
    """tester.py:117-255. Returns (stats, dominant_result[3], rejected_result[3]); `stats` is the reference's list of
    dicts (net first, then one per heuristic) whose values are host arrays with one entry per instance."""
    env.set_testing_mode(batch_size, node_sample_size, req_sample_size, node_min_val, node_max_val)
    agent.set_testing_mode(batch_size, env.node_sample_size, env.profiles_sample_size)
    N = env.node_sample_size
    bufs = env.device_buffers(store=False)
    env.reset_into(bufs)
    heuristic_solvers = heuristic_factory(N, env.normalization_factor, opts["heuristic"], env.device,
                                          seed=env.seed_value, instance_id0=env.episode_id0)
    heuristic_input_state = bufs.batch.clone()   # tester.py:163: the initial state, before the net consumes it

    start = time.time()
    agent.rollout(env, bufs, greedy=True)
    if opts.get("show_inference_progress"):
        torch.cuda.synchronize()
        print(f"Net solution found in {time.time() - start:.2f} seconds", end="\r")

    for solver in heuristic_solvers:
        solver.solve(heuristic_input_state)

    # gather_stats_from_solutions (misc/utils.py:97-159), every instance
    net_dom, net_rej, net_empty = net_solution_stats(bufs, N)
    h_stats = [s.stats() for s in heuristic_solvers]
    net_dom_rounded, results = compare_solutions(net_dom, net_rej, [h[0] for h in h_stats], [h[1] for h in h_stats])
    host = lambda t: t.cpu().numpy()
    stats = [{"net_dominant": host(net_dom_rounded), "net_rejected": host(net_rej),
              "net_empty_nodes": host(net_empty), "net_optimal": 0}]
    for solver, (d, r, e) in zip(heuristic_solvers, h_stats):
        stats.append({f"{solver.name}_dominant": host(d), f"{solver.name}_rejected": host(r),
                      f"{solver.name}_empty_nodes": host(e), f"{solver.name}_optimal": solver.is_optimal})
    res = host(results).astype(np.int64)
    return stats, res[0:3], res[3:6]


def log_testing_stats(global_stats, folder, file_name):
    """misc/csv_writer.py:3-35: same header and columns; one row per instance of the batch."""
    with open(f"{folder}/{file_name}.csv", "w") as fp:
        header = ""
        for entry in global_stats[0]:
            if entry != "instance":
                header += f"{entry};"
        for entry in global_stats[0]["instance"]:
            keys = list(entry.keys())
            header += "".join(f"{' '.join(k.split('_'))};" for k in keys[:4])
        fp.write(f"{header}\n")
        for st in global_stats:
            batch = len(st["instance"][0]["net_dominant"])
            for b in range(batch):
                data = (f"{st['test_instance'] * batch + b};{st['node_sample_size']};{st['node_min_value']};"
                        f"{st['node_max_value']};{st['resource_sample_size']};")
                for entry in st["instance"]:
                    dominant, rejected, empty_nodes, is_optimal = list(entry.values())
                    data += f"{dominant[b]:.3f};{rejected[b]};{empty_nodes[b]};{is_optimal};"
                fp.write(f"{data}\n")


def test(env, agent, opts: dict, log_dir: str):
    """tester.py:19-115 -> (dominant_results, rejected_results): [won, draw, lost] counts of the net against the
    best heuristic over every tested instance. The per-test statistics stay available as `test.global_stats`."""
    add_brakes = opts.get("add_brakes", False)
    tb = opts["testbed"]
    num_tests = tb["num_tests"]
    nodes_cfg, res_cfg, req_cfg = tb["node_sample_configs"], tb["node_available_resources"], tb["request_sample_configs"]
    node_size_min, node_size_max, node_size_step = nodes_cfg["min"], nodes_cfg["max"], nodes_cfg["step"]
    batch_size = opts.get("batch_size", 1)
    gs = opts["export_stats"]["global_stats"]
    filename = gs["filename"] if gs["filename"] is not None else generate_file_name(agent.agent_config)

    global_stats = []
    dominant_results = np.zeros(3, dtype=np.int64)   # won, draw, lost
    rejected_results = np.zeros(3, dtype=np.int64)
    for resource_sample_size in range(req_cfg["min"], req_cfg["max"] + 1, req_cfg["step"]):
        for node_sample_size in range(node_size_min, node_size_max + 1, node_size_step):
            for node_min_value in range(res_cfg["min"], res_cfg["max"], res_cfg["step"]):
                for index in range(num_tests):
                    instance_stats, dom, rej = test_single_instance(
                        index, env, agent, opts, batch_size, node_sample_size, node_min_value,
                        node_min_value + res_cfg["step"], resource_sample_size, log_dir)
                    dominant_results += dom
                    rejected_results += rej
                    global_stats.append({"test_instance": index, "node_sample_size": node_sample_size,
                                         "node_min_value": node_min_value,
                                         "node_max_value": node_min_value + res_cfg["step"],
                                         "resource_sample_size": resource_sample_size, "instance": instance_stats})
            if opts.get("show_per_test_stats"):
                last = global_stats[-1]["instance"][0]
                print(f"nodes {node_sample_size:3d} rules {resource_sample_size:3d}: net mean rejected "
                      f"{np.mean(last['net_rejected']):.3f} mean empty {np.mean(last['net_empty_nodes']):.3f}")
            if add_brakes:
                break
        if add_brakes:
            node_size_min += node_size_step
    if gs["export_stats"] and log_dir is not None:
        folder = os.path.join(log_dir, gs["folder"])
        os.makedirs(folder, exist_ok=True)
        log_testing_stats(global_stats, folder, filename)
    test.global_stats = global_stats
    return dominant_results, rejected_results


def knapsack_test(env, agent, opts: dict, log_dir: str):
    """KnapsackV2: greedy rollouts over the test bed's (items, bins) cells; returns the per-instance collected value
    (episode reward) and the number of items sent to the EOS bin. The reference's KnapsackV2 heuristics (OR-Tools,
    waste reduction; environment/custom/knapsack_v2/heuristic/*) are a 'next' row."""
    tb = opts["testbed"]
    batch_size = opts.get("batch_size", 1)
    bins_cfg, items_cfg = tb["bin_sample_configs"], tb["item_sample_configs"]
    values, rejected = [], []
    for item_size in range(items_cfg["min"], items_cfg["max"] + 1, items_cfg["step"]):
        for bin_size in range(bins_cfg["min"], bins_cfg["max"] + 1, bins_cfg["step"]):
            for _ in range(tb["num_tests"]):
                env.set_testing_mode(batch_size, bin_size, item_size)
                agent.set_testing_mode(batch_size, env.node_sample_size, env.profiles_sample_size)
                bufs = env.device_buffers(store=False)
                env.reset_into(bufs)
                agent.rollout(env, bufs, greedy=True)
                values.append(bufs.rewards.sum(dim=1).cpu().numpy())
                rejected.append((bufs.actions == 0).sum(dim=0).cpu().numpy())
    return np.concatenate(values), np.concatenate(rejected)
