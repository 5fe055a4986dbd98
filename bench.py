#!/usr/bin/env python
"""Headline benchmark: ResourceV3 placement decisions / second for one training iteration
(env reset + T-step rollout + A2C update) on N B200s, beside the CPU reference path.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 2 --warmup 1      # CPU reference arm (oracle port, host cores)

Workload (BASELINE.json configs[1]/headline): 50 nodes x 100 rules, batch 1024 episodes PER GPU (weak scaling),
greedy reward, stochastic action selection, actor 1 layer / critic 3 layers as shipped in configs/ResourceV3.json.
One step = B*T = 102 400 decisions per GPU. Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = {"nodes": 50, "rules": 100, "batch_per_gpu": 1024}
FLOP_PER_DECISION = 1618.7e6     # rollout + A2C update at 50x100


def load_cfg(batch, nodes, rules, reward="greedy"):
    import copy
    from tpc_b200.configs import get_configs
    agent_cfg, trainer_cfg, env_cfg, tester_cfg, tuner_cfg, _ = get_configs("ResourceV3", "tpc")
    env_cfg = copy.deepcopy(env_cfg)
    env_cfg.update(batch_size=batch, node_sample_size=nodes, profiles_sample_size=rules)
    env_cfg["reward"]["type"] = reward
    return agent_cfg, env_cfg


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs (recipe's clocks line)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def dataflow_bytes_per_state(N, R, B=1024, chunk=10240, actor=(1, 128), critic=(3, 512), d=128):
    """HBM bytes one stored state moves through the UPDATE in the current kernel decomposition,
    counted in units of one [token, 128] fp32 row (512 B) per encoder token-layer with dff = m * 128:
      forward  : QKV 1+3, attention 3+1, O-proj + LN 1+1+1, FFN1 1+m, FFN2 + LN m+1+1             = 14 + 2m
      backward : 2 x LN 3, dW2 m+1, dH 1+m, dW1 1+m, d(out1) m+1+1, dWo 2, d(att) 2, attention 3+1+1+3, dWqkv 1+3,
                 dx 3+1+1                                                                          = 32 + 4m
                 LayerNorm 1 backward rides in the epilogue of the d(out1) GEMM (-3 for ln_bwd, +1 for the saved
                 LayerNorm output the epilogue reads) and LayerNorm 2 backward in the dx GEMM of the layer above it,
                 for every layer but the top one of a stack                              = 30 + 4m - 2 (layers - 1) / layers
    times the tokens of the three stacks (bins N, rules R, joint N + R) x layers. The critic keeps every token; the
    actor leaves out the rule tokens placed before the first step of its chunk (token windows): chunk k of `chunk`
    states drops floor(k * chunk / B) rules, the rollout step t drops t. The actor's decoder adds the K2|V2|W2e
    projection of its tokens (1 + 3 forward, 3 + 1 + 1 + 3 backward rows per token)."""
    u = d * 4
    T = R
    n_chunks = max(1, -(-T * B // chunk))
    kept_update = sum(R - min(R - 1, (k * chunk) // B) for k in range(n_chunks)) / n_chunks
    kept_rollout = sum(R - t for t in range(T)) / T

    def encoder(layers, dff, rules):
        m = dff // d
        tokens = N + rules + (N + rules)                      # bins stack + rules stack + joint stack
        return tokens * layers, (14 + 2 * m), (30 + 4 * m - 2 * (layers - 1) / max(layers, 1))
    tl, f, b = encoder(*critic, R)
    total = tl * (f + b) * u
    tl, f, b = encoder(*actor, kept_update)
    total += tl * (f + b) * u + (N + kept_update) * 12 * u
    tl, f, b = encoder(*actor, kept_rollout)
    rollout = tl * f * u + (N + kept_rollout) * 4 * u
    return int(total), int(rollout)


CPU_EPISODES = 32


def cpu_sample_text(N, R, episodes):
    return (f"one COMPLETE training iteration at {N - 1} nodes x {R} rules with batch {episodes} instead of 1024: "
            f"{R}-step rollout of {episodes} episodes (actor forward per step), returns, one critic and one actor "
            f"forward/backward over the {R * episodes} stored states, both Adam steps (torch-CPU fp32 oracle port, all "
            f"host threads); decisions/s = {R * episodes} / wall time")


def cpu_baseline(env_cfg, N, R, episodes=CPU_EPISODES):
    from oracle.cpu_path import time_full_iteration
    r = time_full_iteration(env_cfg, N, R, episodes)
    return {"value": r["decisions_per_s"], "unit": "decisions/s", "cores": r["threads"], "kind": "port",
            "sample": cpu_sample_text(N, R, episodes), "rollout_s": r["t_rollout_s"], "update_s": r["t_update_s"]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _, env_cfg = load_cfg(WORKLOAD["batch_per_gpu"], args.nodes, args.rules, args.reward)
    N, R = args.nodes + 1, args.rules
    from oracle.cpu_path import time_full_iteration
    for _ in range(args.warmup):
        time_full_iteration(env_cfg, N, R, 2)
    vals, t0 = [], time.perf_counter()
    for _ in range(args.steps):
        vals.append(time_full_iteration(env_cfg, N, R, CPU_EPISODES))
    wall = time.perf_counter() - t0
    v = statistics.mean(x["decisions_per_s"] for x in vals)
    cores = vals[0]["threads"]
    sample = "per step: " + cpu_sample_text(N, R, CPU_EPISODES)
    print(json.dumps({
        "impl": "reference", "metric": "ResourceV3 placement decisions/sec (rollout+A2C update)", "value": v,
        "unit": "decisions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall / max(args.steps, 1) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"ResourceV3 {args.nodes} nodes x {args.rules} rules, batch "
                               f"{WORKLOAD['batch_per_gpu']}/GPU, {args.reward} reward, rollout + A2C update (actor 1 layer, "
                               f"critic 3 layers)", "timing": sample},
        "cpu_baseline": {"value": v, "unit": "decisions/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "decisions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def roofline_probe(lib, h, st, M, dev):
    """The launch of the dominant kernel that the roofline is quoted on: gemm_tc_kernel as the critic's second FFN
    layer (K = 512 -> 128, fp32-accurate FP16-split product) with the residual add + LayerNorm epilogue, M tokens of one update chunk.
    Returns (callable, algorithmic bytes per launch, label). tests/gpu_checks/roofline_probe.py runs the same
    launch under ncu for `profiles/r02f_ncu_full_gemm_tc_ffn2.json`."""
    import torch
    from tpc_b200 import lib as L
    K = 512
    A = torch.randn(M, K, device=dev)
    W = torch.randn(128, K, device=dev) / K ** 0.5
    Wlo = torch.empty_like(W)   # fp16 split image of the weights (tpc.h: tpc_gemm_split_image)
    L.check(lib.tpc_gemm_split_image(h, L.ptr(W), K, 128, K, L.ptr(Wlo), st))
    res = torch.randn(M, 128, device=dev)
    D = torch.empty(M, 128, device=dev)
    g = torch.ones(128, device=dev)
    b = torch.zeros(128, device=dev)
    rstd = torch.empty(M, device=dev)
    keep = (A, W, Wlo, res, D, g, b, rstd)

    def call(_keep=keep):
        L.check(lib.tpc_gemm(h, L.ptr(A), K, L.ptr(W), K, L.ptr(D), 128, M, 128, K, L.ptr(b), L.ptr(res), 128, 1,
                             L.ptr(g), L.ptr(b), L.ptr(rstd), 2, 1, L.ptr(Wlo), st))
    # read the activations and the residual once, write the normalised output and 1/sigma once
    alg_bytes = (M * K + M * 128 + M * 128 + M) * 4
    return call, alg_bytes, "FFN2 512->128 fp32-accurate (FP16x2 split) + residual + LayerNorm"


def run_check(args, world, rank, dev, B, nodes, rules):
    """Multi-GPU correctness: every rank rolls out its shard of the GLOBAL batch (episode ids rank*B ...),
    runs the overlapped update (critic all-reduce behind the actor tape) and keeps the all-reduced gradients; rank 0
    then repeats the iteration alone on the full global batch from the same parameters. Asserted: actions and rewards
    bit-identical for any rank count, all-reduced gradients equal to the full-batch gradients (per tensor, norm-wise,
    2e-3: two runs of the same single-GPU call already differ by 1e-4 -- fp32 atomics + TF32 backward), losses to 1e-4.
    Prints one JSON line with the measured differences."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from tpc_b200 import lib as L
    from tpc_b200.agent import Agent
    from tpc_b200.env import ResourceEnvironmentV3

    def one_pass(batch, id0, ws, allreduce):
        agent_cfg, env_cfg = load_cfg(batch, nodes, rules, args.reward)
        env = ResourceEnvironmentV3("ResourceV3", env_cfg, dev)
        env.add_stats_to_agent_config(agent_cfg)
        agent = Agent("transformer", agent_cfg, dev, num_features=3, seed=0)
        env.episode_id0 = agent.episode_id0 = id0
        bufs = env.device_buffers(store=True)
        env.reset_into(bufs)
        agent.rollout(env, bufs)
        eng = agent.engine
        hyper = (agent.gamma, agent.values_loss_coefficient, agent.entropy_coefficient)
        total = batch * ws * bufs.R
        if allreduce is None:
            eng.a2c_grads(bufs, *hyper, total_states=total)
        else:
            eng.a2c_grads(bufs, *hyper, total_states=total, phases=L.A2C_CRITIC)
            w1 = allreduce(eng.grads[1])
            eng.a2c_grads(bufs, *hyper, total_states=total, phases=L.A2C_ACTOR)
            for w in (w1, allreduce(eng.grads[0]), allreduce(eng.losses)):
                w.wait()
        torch.cuda.synchronize()
        names[0], names[1] = eng.tensor_names(0), eng.tensor_names(1)
        return (bufs.actions.clone(), bufs.rewards.clone(), eng.grads[0].clone(), eng.grads[1].clone(),
                (eng.losses / float(total)).clone(), eng.layout(0), eng.layout(1))

    names = [None, None]

    ar = (lambda t: dist.all_reduce(t, async_op=True)) if world > 1 else None
    acts, rews, ga, gc, losses, _, _ = one_pass(B, rank * B, world, ar)
    if world > 1:
        all_a = [torch.empty_like(acts) for _ in range(world)]
        all_r = [torch.empty_like(rews) for _ in range(world)]
        dist.all_gather(all_a, acts)
        dist.all_gather(all_r, rews)
        acts, rews = torch.cat(all_a, dim=1), torch.cat(all_r, dim=0)
    if rank == 0:
        fa, fr, fga, fgc, flosses, la, lc = one_pass(B * world, 0, 1, None)

        def per_tensor(g, ref, layout, names):
            """Worst per-tensor relative error (norm-wise) over the tensors that carry a gradient. Tensors whose true
            gradient is zero (every wk.bias: softmax rows are shift invariant; the decoder's mha1.wq / wk: one key) hold
            rounding noise only -- it changes with the summation order, i.e. with the token windows of the chunks, and
            Adam's clipnorm never sees it as signal -- so tensors below 1e-3 of the largest tensor norm are left out."""
            norms = [float(ref[o:o + n].norm()) for o, n in zip(*layout)]
            worst, where = 0.0, ""
            for (o, n), nr, name in zip(zip(*layout), norms, names):
                if nr > 1e-3 * max(norms):
                    e = float((g[o:o + n] - ref[o:o + n]).norm()) / nr
                    if e > worst:
                        worst, where = e, name
            return worst, where
        ea, wa = per_tensor(ga, fga, la, names[0])
        ec, wc = per_tensor(gc, fgc, lc, names[1])
        res = {"check": "multi_gpu", "n_gpus": world, "global_batch": B * world, "nodes": nodes, "rules": rules,
               "actions_equal": bool(torch.equal(acts, fa)), "rewards_equal": bool(torch.equal(rews, fr)),
               "actor_grad_rel": ea, "actor_worst_tensor": wa, "critic_grad_rel": ec, "critic_worst_tensor": wc,
               "actor_grad_rel_whole": float((ga - fga).norm() / fga.norm()),
               "critic_grad_rel_whole": float((gc - fgc).norm() / fgc.norm()),
               "loss_rel": float(((losses - flosses).abs() / flosses.abs().clamp_min(1e-12)).max())}
        res["ok"] = bool(res["actions_equal"] and res["rewards_equal"] and res["actor_grad_rel"] < 2e-3
                         and res["critic_grad_rel"] < 2e-3 and res["loss_rel"] < 1e-4)
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and not res["ok"]:
        raise SystemExit(1)


def run_ours(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from tpc_b200 import lib as L
    from tpc_b200.agent import Agent
    from tpc_b200.env import ResourceEnvironmentV3
    from tpc_b200.trainer import train_iteration

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, nodes, rules = args.batch or WORKLOAD["batch_per_gpu"], args.nodes, args.rules
    if args.scaling == "strong":
        # total work fixed: the GLOBAL batch (default 1024) is split over the ranks
        gb = args.global_batch or WORKLOAD["batch_per_gpu"]
        if gb % world:
            raise SystemExit(f"--global-batch {gb} is not divisible by {world} ranks")
        B = gb // world
    if args.check:
        return run_check(args, world, rank, dev, B, nodes, rules)
    agent_cfg, env_cfg = load_cfg(B, nodes, rules, args.reward)
    env = ResourceEnvironmentV3("ResourceV3", env_cfg, dev)
    env.add_stats_to_agent_config(agent_cfg)
    agent = Agent("transformer", agent_cfg, dev, num_features=3, seed=0)
    N, R, F = env.node_sample_size, env.profiles_sample_size, 3
    S, T = N + R, R
    lib = L.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput ("value")
    bufs = None
    for _ in range(args.warmup):
        bufs, losses = train_iteration(env, agent, bufs, chunk_states=args.chunk)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = lib.tpc_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        bufs, losses = train_iteration(env, agent, bufs, chunk_states=args.chunk)
    e1.record()
    barrier()
    launches = lib.tpc_launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop() if sampler else None
    value = world * B * T * args.steps / (ms_total / 1e3)
    loss_host = losses.cpu().numpy().tolist()

    # ---------------- end to end through the public API with HOST buffers ("e2e")
    # every step: pinned host instance batch -> device, rollout + update, losses + per-episode rewards -> host
    host_batches = []
    for _ in range(args.steps):
        env.reset_into(bufs)
        host_batches.append(bufs.batch.cpu().pin_memory())
    host_out = torch.empty(B + 4, dtype=torch.float32).pin_memory()
    barrier()
    e0.record()
    for hb in host_batches:
        bufs.batch.copy_(hb, non_blocking=True)
        bufs.bin_net_mask.zero_(); bufs.bin_net_mask[:, N:] = 1.0; bufs.mha_mask.zero_()
        agent.rollout(env, bufs)
        lo = agent.update(bufs, world_size=world, chunk_states=args.chunk,
                          allreduce=(lambda t: dist.all_reduce(t, async_op=True)) if world > 1 else None)
        host_out[:B].copy_(bufs.rewards.sum(dim=1), non_blocking=True)
        host_out[B:].copy_(lo, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e = world * B * T * args.steps / (float(ms2.item()) / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (gemm_tc_kernel), measured live with CUDA events
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    n_states = B * T
    pieces = (n_states + 10239) // 10240
    chunk = args.chunk or -(-n_states // pieces)
    M = chunk * S
    h = agent.engine.h
    st = agent.engine.stream
    call, alg_bytes, label = roofline_probe(lib, h, st, M, dev)
    for _ in range(3):
        call()
    e0.record()
    reps = 20
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    t_gemm = e0.elapsed_time(e1) / reps / 1e3
    traffic = None
    # dram__bytes_read.sum + dram__bytes_write.sum of this launch from the committed `ncu --set full` capture of the
    # current kernel (tests/gpu_checks/run_profiles_r02f.sh)
    tpath = os.path.join(ROOT, "profiles", "r02f_ncu_full_gemm_tc_ffn2.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath))["gemm_tc_kernel"].get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    # tensor-core view of the whole step: dense algorithmic FLOPs per decision x decisions/s per GPU against
    # a tcgen05 kind::tf32 peak measured right here (resident operands, every SM issuing: tpc_tf32_peak_probe)
    tf32 = C.c_float(0.0)
    L.check(lib.tpc_tf32_peak_probe(h, 20000, C.byref(tf32), st), "tpc_tf32_peak_probe")
    step_tflops = value / world * FLOP_PER_DECISION / 1e12 if (nodes, rules) == (50, 100) else None
    tensor = {"bound": "tensor", "achieved": step_tflops, "peak": float(tf32.value), "unit": "TFLOP/s",
              "frac": (step_tflops / float(tf32.value)) if step_tflops else None,
              "peak_source": "measured live: tcgen05.mma kind::tf32 M128 N256 K8, operands resident in shared memory, "
                             "148 CTAs (tpc_tf32_peak_probe); the forward GEMMs issue 3 kind::f16 MMAs (twice the TF32 rate) per "
                             "algorithmic product, the backward GEMMs one kind::tf32 MMA",
              "flop_per_decision": FLOP_PER_DECISION}
    roofline = {"bound": "hbm", "achieved": alg_bytes / t_gemm / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes / t_gemm / 1e9 / hbm_peak, "traffic": traffic,
                "kernel": f"gemm_tc_kernel ({label}, M={M}, {t_gemm * 1e6:.1f} us/launch)",
                "peak_source": peak_src, "step_tensor_tflops": step_tflops, "tensor": tensor}

    cpu = cpu_baseline(env_cfg, N, R) if world == 1 and not args.no_cpu else None
    out = {
        "metric": "ResourceV3 placement decisions/sec (rollout+A2C update)", "value": value, "unit": "decisions/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32 (forward: fp32-accurate two-term FP16 split products with fp32 accumulation; backward: TF32)",
        "data": "synthetic",
        "config": {"workload": f"ResourceV3 {nodes} nodes x {rules} rules, batch {B}/GPU, {args.reward} reward, rollout + A2C "
                               f"update (actor 1 layer, critic 3 layers)",
                   "decisions_per_step": world * B * T, "chunk_states": chunk,
                   "hbm_bytes_per_state": {"update": dataflow_bytes_per_state(N, R, B, chunk)[0],
                                           "rollout": dataflow_bytes_per_state(N, R, B, chunk)[1],
                                           "note": "modelled from the kernel decomposition (bench.py "
                                                   "dataflow_bytes_per_state), not a counter"},
                   "l2": "working set per step (>4 GB of activations) far exceeds the 126 MB L2; no flush needed",
                   "parallelism": f"dp{world}"},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": "decisions/s", "h2d_bytes_per_step": B * S * F * 4, "d2h_bytes_per_step": (B + 4) * 4},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "losses_last_step": loss_host,
    }
    if cpu is not None:
        out["cpu_baseline"] = cpu
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="episodes per GPU (default 1024)")
    ap.add_argument("--nodes", type=int, default=WORKLOAD["nodes"])
    ap.add_argument("--rules", type=int, default=WORKLOAD["rules"])
    ap.add_argument("--chunk", type=int, default=0, help="states per update chunk (default: ~10240, equal chunks)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch episodes per GPU; strong: --global-batch episodes split over the ranks")
    ap.add_argument("--global-batch", type=int, default=0, help="strong scaling: total episodes (default 1024)")
    ap.add_argument("--check", action="store_true",
                    help="multi-GPU correctness instead of timing: N-rank gradients + all-reduce vs the 1-rank "
                         "full-batch gradient, trajectories independent of the rank count")
    ap.add_argument("--reward", default="greedy", help="ResourceV3 reward type (BASELINE config #3: global_dominant = "
                    "fair, reduced_node_usage = cost); the headline metric is quoted on greedy")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: native libraries (the NCCL version banner, for instance) write to file
    # descriptor 1 directly, so fd 1 is pointed at stderr for the whole run and Python's sys.stdout keeps the real one.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
