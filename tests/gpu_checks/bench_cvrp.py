"""BASELINE config #5: CVRP (configs/CVRP.json: 32 nodes incl. depot, 5 vehicles) rollout throughput sweep over the
batch size, greedy decode through the fused single-pointer adapter (tpc_rollout, env kind TPC_ENV_CVRP).
One rank per GPU (torchrun) or a single process; the rollout has no collective, so ranks run independent shards of the
global batch (weak scaling) and the time is the max over ranks. Prints one JSON line per batch size (rank 0)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tpc_b200.agent import Agent  # noqa: E402
from tpc_b200.configs import get_configs  # noqa: E402
from tpc_b200.env import CVRP  # noqa: E402


def main():
    import torch.distributed as dist
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    batches = [int(b) for b in sys.argv[1:]] or [256, 1024, 4096, 16384]
    for B in batches:
        agent_cfg, _, env_cfg, _, _, _ = get_configs("CVRP", "tpc")
        env_cfg = dict(env_cfg)
        env_cfg.update(batch_size=B)
        env = CVRP("CVRP", env_cfg, dev)
        env.episode_id0 = rank * B
        agent = Agent("transformer", env.add_stats_to_agent_config(dict(agent_cfg)), dev, num_features=3, seed=0)
        bufs = env.device_buffers(store=False)
        for _ in range(3):
            env.reset_into(bufs)
            agent.rollout(env, bufs, greedy=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            env.reset_into(bufs)
            agent.rollout(env, bufs, greedy=True)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            T = env.steps
            print(json.dumps({"workload": "CVRP 32 nodes x 5 vehicles, greedy rollout (single-pointer adapter)",
                              "n_gpus": world, "batch_per_gpu": B, "steps_per_episode": T,
                              "ms_per_rollout": float(ms.item()),
                              "decisions_per_s": world * B * T / (float(ms.item()) / 1e3),
                              "mean_route_cost": float(-bufs.rewards.sum(1).mean().item())}))
        agent.engine.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
