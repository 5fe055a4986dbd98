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


def knapsack_test_single_instance(instance_id, env, agent, opts: dict, batch_size: int, bin_sample_size: int,
                                  bin_min_val: int, bin_max_val: int, item_sample_size: int, log_dir: str):
    """knapsack_v2/tester.py:101-233 -> (stats, reward_result[3]); statistics for every instance of the batch."""
    from .heuristics import knapsack_heuristic_factory, knapsack_net_stats
    env.set_testing_mode(batch_size, bin_sample_size, item_sample_size, bin_min_val, bin_max_val)
    agent.set_testing_mode(batch_size, env.node_sample_size, env.profiles_sample_size)
    N = env.node_sample_size
    bufs = env.device_buffers(store=False)
    env.reset_into(bufs)
    solvers = knapsack_heuristic_factory(N, opts["heuristic"], env.device, seed=env.seed_value,
                                         instance_id0=env.episode_id0)
    heuristic_input_state = bufs.batch.clone()
    agent.rollout(env, bufs, greedy=True)
    for solver in solvers:
        solver.solve(heuristic_input_state)
    (net_reward, net_empty, net_rej, net_rej_value), _ = knapsack_net_stats(bufs, heuristic_input_state, N)
    h_stats = [s.stats() for s in solvers]
    # gather_stats_from_solutions (knapsack_v2/misc/utils.py:60-95): rounded rewards, heuristics start from -1
    dummy = torch.zeros_like(net_rej)
    net_rounded, results = compare_solutions(net_reward, dummy, [h[0] for h in h_stats], [dummy for _ in h_stats])
    host = lambda t: t.cpu().numpy()
    stats = [{"net_reward": host(net_rounded), "net_empty_bins": host(net_empty),
              "net_num_rejected_items": host(net_rej), "net_rejected_value": host(net_rej_value)}]
    for solver, (r, e, n, v) in zip(solvers, h_stats):
        stats.append({f"{solver.name}_reward": host(r), f"{solver.name}_empty_bins": host(e),
                      f"{solver.name}_num_rejected_items": host(n), f"{solver.name}_rejected_value": host(v)})
    return stats, host(results).astype(np.int64)[0:3]


def knapsack_log_testing_stats(global_stats, folder, file_name):
    """knapsack_v2/misc/csv_writer.py:5-36: same header and columns; one row per instance of the batch."""
    with open(f"{folder}/{file_name}.csv", "w") as fp:
        header = "".join(f"{k};" for k in global_stats[0] if k != "instance")
        for entry in global_stats[0]["instance"]:
            header += "".join(f"{' '.join(k.split('_'))};" for k in list(entry.keys())[:4])
        fp.write(f"{header}\n")
        for st in global_stats:
            batch = len(st["instance"][0]["net_reward"])
            for b in range(batch):
                data = (f"{st['test_instance'] * batch + b};{st['bin_sample_size']};{st['bin_min_value']};"
                        f"{st['bin_max_value']};{st['item_sample_size']};")
                for entry in st["instance"]:
                    reward, empty, rejected, rej_value = list(entry.values())
                    data += f"{reward[b]:.3f};{empty[b]};{rejected[b]};{rej_value[b]:.3f};"
                fp.write(f"{data}\n")


def knapsack_test(env, agent, opts: dict, log_dir: str):
    """knapsack_v2/tester.py:14-99 -> global_reward_results [won, draw, lost] of the net against the best heuristic
    (waste-reduction combos and random; the reference's ORTools solver needs Google OR-Tools and is left out)."""
    tb = opts["testbed"]
    batch_size = opts.get("batch_size", 1)
    bins_cfg, cap_cfg, items_cfg = tb["bin_sample_configs"], tb["bin_available_capacities"], tb["item_sample_configs"]
    gs = opts["export_stats"]["global_stats"]
    filename = gs["filename"] if gs["filename"] is not None else generate_file_name(agent.agent_config)
    global_stats = []
    global_reward_results = np.zeros(3, dtype=np.int64)
    for item_sample_size in range(items_cfg["min"], items_cfg["max"] + 1, items_cfg["step"]):
        for bin_sample_size in range(bins_cfg["min"], bins_cfg["max"] + 1, bins_cfg["step"]):
            for bin_min_value in range(cap_cfg["min"], cap_cfg["max"], cap_cfg["step"]):
                for index in range(tb["num_tests"]):
                    instance_stats, reward_result = knapsack_test_single_instance(
                        index, env, agent, opts, batch_size, bin_sample_size, bin_min_value,
                        bin_min_value + cap_cfg["step"], item_sample_size, log_dir)
                    global_reward_results += reward_result
                    global_stats.append({"test_instance": index, "bin_sample_size": bin_sample_size,
                                         "bin_min_value": bin_min_value,
                                         "bin_max_value": bin_min_value + cap_cfg["step"],
                                         "item_sample_size": item_sample_size, "instance": instance_stats})
    if gs["export_stats"] and log_dir is not None:
        folder = os.path.join(log_dir, gs["folder"])
        os.makedirs(folder, exist_ok=True)
        knapsack_log_testing_stats(global_stats, folder, filename)
    knapsack_test.global_stats = global_stats
    return global_reward_results
