"""Training loop with the reference's signature and return value (agents/trainer.py:8-192), running the
rollout and the A2C update on the device.

Per iteration (trainer.py:37-175): env.reset -> T = num_resources steps of {act, env.step, store} (tpc_rollout)
-> Monte-Carlo returns with bootstrap 0 (n_steps_to_update is dead code in the reference) ->
critic and actor losses + gradients (tpc_a2c_grads) -> [data-parallel: sum all-reduce of the flat gradient
buffers over NCCL] -> two Keras-Adam steps (tpc_adam_apply).
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from . import lib as L
from .agent import Agent


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def train_iteration(env, agent: Agent, bufs=None, chunk_states: int = 0):
    """One pass of the hot path. Returns (bufs, losses[4] device tensor: value, policy, total actor, entropy)."""
    dist, rank, world = _dist()
    env.episode_id0 = agent.episode_id0 = rank * env.batch_size
    if bufs is None or bufs.B != env.batch_size or bufs.N != env.node_sample_size or bufs.R != env.profiles_sample_size:
        bufs = env.device_buffers(store=True)
    env.reset_into(bufs)
    agent.rollout(env, bufs)
    allreduce = (lambda t: dist.all_reduce(t, async_op=True)) if world > 1 else None
    losses = agent.update(bufs, world_size=world, allreduce=allreduce, chunk_states=chunk_states)
    return bufs, losses


def trainer(env, agent: Agent, opts: dict, show_progress: bool, log_dir: str):
    export_weights: bool = opts["store_model_weights"]["export_weights"]
    n_iterations: int = opts["n_iterations"]
    _n_steps_to_update: int = opts["n_steps_to_update"]  # read but without effect, as in the reference

    average_rewards_buffer, min_rewards_buffer, max_rewards_buffer = [], [], []
    value_loss_buffer, bins_policy_loss_buffer, bins_total_loss_buffer, bins_entropy_buffer = [], [], [], []

    dist, rank, world = _dist()
    bufs = None
    episode_count = 0
    while episode_count < n_iterations:
        start = time.time()
        bufs, losses = train_iteration(env, agent, bufs)
        episode_count += 1
        # statistics (trainer.py:89-100,159-162): one small device->host read per iteration
        stats = torch.empty(3, dtype=torch.float32, device=bufs.rewards.device)
        eng = agent.engine
        L.check(eng.lib.tpc_episode_reward_stats(eng.h, L.ptr(bufs.rewards), bufs.B, bufs.R, L.ptr(stats), eng.stream),
                "tpc_episode_reward_stats")
        if world > 1:
            mn, mx, av = stats[0:1].clone(), stats[1:2].clone(), stats[2:3].clone()
            dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(av)
            stats = torch.cat([mn, mx, av / world])
        host = torch.cat([stats, losses]).cpu().numpy()
        min_in_batch, max_in_batch, episode_reward = host[0], host[1], host[2]
        value_loss, policy_loss, bin_loss, entropy = host[3], host[4], host[5], host[6]
        average_rewards_buffer.append(episode_reward)
        min_rewards_buffer.append(min_in_batch)
        max_rewards_buffer.append(max_in_batch)
        value_loss_buffer.append(value_loss)
        bins_policy_loss_buffer.append(policy_loss)
        bins_total_loss_buffer.append(bin_loss)
        bins_entropy_buffer.append(entropy)
        if show_progress and rank == 0:
            print(f"\rEp: {episode_count} took {time.time() - start:.2f} sec.\t"
                  f"Min@Batch: {min_in_batch:.3f}\tMax@Batch: {max_in_batch:.3f}\tAvg@Batch: {episode_reward:.3f}\t"
                  f"Critic_Loss: {value_loss:.3f} \tActor_Total_Loss: {bin_loss:.3f} \t"
                  f"Policy_Loss: {policy_loss:.3f} \tEntropy_Loss: {entropy:.3f}", end="\n")
        agent.clear_memory()

    if export_weights and rank == 0:
        weight_location = os.path.join(log_dir, opts["store_model_weights"]["folder"],
                                       opts["store_model_weights"]["filename"])
        agent.save_weights(weight_location)

    return (average_rewards_buffer, min_rewards_buffer, max_rewards_buffer, value_loss_buffer,
            bins_policy_loss_buffer, bins_total_loss_buffer, bins_entropy_buffer)
