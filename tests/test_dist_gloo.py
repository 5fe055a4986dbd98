"""World-size-2 (gloo, CPU) check of the data-parallel recipe used by tpc_b200.trainer / Agent.update:
each rank differentiates the MEAN-over-all-states losses on its own episode shard using the GLOBAL state count as
denominator (tpc_a2c_hyper.total_states), gradients are SUM all-reduced, and the result equals the single-process
gradient of the full batch. Uses the oracle model as the stand-in for the device kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import model as om


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    rng = np.random.default_rng(5)
    n, N, R = 8, 3, 3
    S = N + R
    st = rng.random((n, S, 3)).astype(np.float32)
    st[:, 0] = -2
    pm = np.zeros((n, S), np.float32)
    pm[:, N:] = 1
    mm = np.zeros((n, 1, 1, S), np.float32)
    acts = rng.integers(0, N, size=n)
    adv = rng.standard_normal((n, 1)).astype(np.float32)
    return st, pm, mm, acts, adv, N


def _grad(flat, c, sl, total):
    st, pm, mm, acts, adv, N = _problem()
    f = flat.clone().requires_grad_(True)
    logits, _, _ = om.actor_forward(f, c, torch.tensor(st[sl]), torch.tensor(st[sl][:, N:N + 1]), torch.tensor(pm[sl]),
                                    torch.tensor(mm[sl]), N)
    tot, ent, pl, _ = om.actor_loss(logits, torch.tensor(acts[sl]), torch.tensor(adv[sl]), 0.01)
    n_local = logits.shape[0]
    (tot * n_local / total).backward()       # sum over local states / GLOBAL count
    return f.grad


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    c = om.NetCfg(num_layers=1, dff=128)
    flat = torch.tensor(om.init_params(om.actor_spec(c), 9))
    g = _grad(flat, c, slice(rank * 4, rank * 4 + 4), 8)
    dist.all_reduce(g)
    if rank == 0:
        torch.save(g, out)
    dist.destroy_process_group()


def test_sum_allreduce_of_shard_gradients_equals_full_batch_gradient(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    g2 = torch.load(out)
    c = om.NetCfg(num_layers=1, dff=128)
    flat = torch.tensor(om.init_params(om.actor_spec(c), 9))
    g1 = _grad(flat, c, slice(0, 8), 8)
    assert float((g1 - g2).abs().max()) < 1e-6 * max(1.0, float(g1.abs().max()))
