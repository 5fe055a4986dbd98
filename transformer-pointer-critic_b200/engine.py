"""Device engine: owns the libtpc handle/model, the flat parameter / gradient / Adam buffers (torch-allocated
device memory) and the scratch workspace; exposes the hot-path calls as Python methods on torch tensors.

PyTorch is used for allocation, streams and torch.distributed only: every arithmetic op below is a libtpc call.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np
import torch

from . import lib as L

REWARD_TYPES = {"greedy": 0, "single_node_dominant": 1, "global_dominant": 2, "reduced_node_usage": 3}


@dataclass
class Shapes:
    num_bins: int        # N (nodes + dummy)
    num_resources: int   # R
    num_features: int    # env feature count (decoder-input width)

    @property
    def S(self):
        return self.num_bins + self.num_resources


def model_config_from_agent_config(agent_cfg: dict, num_features: int) -> L.ModelConfig:
    """agents/models/model_factory.py:4-36 + env-injected keys (environment/custom/resource_v3/env.py:314-334)."""
    a, c = agent_cfg["actor"], agent_cfg["critic"]
    for k in ("dim_model", "num_heads"):
        if a[k] != c[k]:
            raise ValueError(f"actor/critic {k} differ: {a[k]} vs {c[k]}")
    if (c.get("last_layer_activation") or "linear") != "linear":
        raise NotImplementedError("only last_layer_activation='linear' (every shipped config) is supported")
    emb = agent_cfg["encoder_embedding"]
    clipc = a.get("logit_clipping_C")
    mc = L.ModelConfig()
    mc.actor_num_layers, mc.actor_inner_layer_dim = a["num_layers"], a["inner_layer_dim"]
    mc.critic_num_layers, mc.critic_inner_layer_dim = c["num_layers"], c["inner_layer_dim"]
    mc.dim_model, mc.num_heads = a["dim_model"], a["num_heads"]
    mc.attention_dense_units, mc.last_layer_units = a["attention_dense_units"], c["last_layer_units"]
    mc.logit_clipping_C = float(clipc) if clipc is not None else 0.0
    mc.has_logit_clipping = 0 if clipc is None else 1
    mc.common_embedding = 1 if emb["common"] else 0
    mc.num_bin_features = emb["num_bin_features"] or num_features
    mc.num_resource_features = emb["num_resource_features"] or num_features
    mc.dec_features = num_features
    mc.num_bins, mc.num_resources = agent_cfg["num_bins"], agent_cfg["num_resources"]
    return mc


def env_config_struct(env_cfg: dict, kind: int = 0) -> L.EnvConfig:
    """configs/ResourceV3.json env_config / configs/KnapsackV2.json env_config -> tpc_env_config."""
    e = L.EnvConfig()
    e.kind = kind
    rw = env_cfg["reward"]
    e.reward_type = REWARD_TYPES[rw["type"]] if kind == 0 else 0
    sub = rw.get(rw["type"], {}) or {}
    e.rejection_penalty = float(sub.get("rejection_penalty", 0.0))
    e.use_new_node_penalty = float(sub.get("use_new_node_penalty", 0.0))
    e.mask_nodes_in_mha = 1 if env_cfg["mask_nodes_in_mha"] else 0
    e.num_features = env_cfg["num_features"]
    e.eos_code = float(env_cfg["EOS_CODE"])
    e.normalization_factor = env_cfg["normalization_factor"]
    if kind == 0:
        e.node_min_val, e.node_max_val = env_cfg["node_min_val"], env_cfg["node_max_val"]
        e.req_min_val, e.req_max_val = env_cfg["req_min_val"], env_cfg["req_max_val"]
        e.val_min_val = e.val_max_val = 0
        e.num_profiles = 0 if env_cfg["generate_request_on_the_fly"] else env_cfg["num_profiles"]
    else:
        e.node_min_val, e.node_max_val = env_cfg["bin_min_capacity"], env_cfg["bin_max_capacity"]
        e.req_min_val, e.req_max_val = env_cfg["item_min_weight"], env_cfg["item_max_weight"]
        e.val_min_val, e.val_max_val = env_cfg["item_min_value"], env_cfg["item_max_value"]
        e.num_profiles = 0 if env_cfg["generate_items_on_the_fly"] else env_cfg["item_set"]
    return e


class Engine:
    """One per process / GPU."""

    def __init__(self, model_cfg: L.ModelConfig, device: str | torch.device = "cuda:0"):
        if not torch.cuda.is_available():
            raise L.TpcError("tpc_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = L.load()
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.h = C.c_void_p()
        L.check(self.lib.tpc_create(C.byref(self.h)), "tpc_create")
        self.cfg = model_cfg
        self.m = C.c_void_p()
        L.check(self.lib.tpc_model_create(self.h, C.byref(model_cfg), C.byref(self.m)), "tpc_model_create")
        self.n_params = [self.lib.tpc_param_count(self.m, w) for w in (0, 1)]
        self.n_shadow = [self.lib.tpc_shadow_count(self.m, w) for w in (0, 1)]
        z = lambda n: torch.zeros(n, dtype=torch.float32, device=self.device)
        self.params = [z(n) for n in self.n_params]
        self.shadow = [z(n) for n in self.n_shadow]
        self.grads = [z(n) for n in self.n_params]
        self.adam_m = [z(n) for n in self.n_params]
        self.adam_v = [z(n) for n in self.n_params]
        self.adam_step = [0, 0]
        self.losses = z(4)
        self._ws = None

    # ------------------------------------------------------------------ plumbing
    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def workspace(self, nbytes: int) -> torch.Tensor:
        if nbytes == 0:
            raise L.TpcError("workspace query failed (unsupported shape?)")
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes * 1.02) + 4096, dtype=torch.uint8, device=self.device)
        return self._ws

    def layout(self, which: int):
        n = self.lib.tpc_num_tensors(self.m, which)
        offs = (C.c_int64 * n)()
        sizes = (C.c_int64 * n)()
        L.check(self.lib.tpc_param_layout(self.m, which, offs, sizes), "tpc_param_layout")
        return list(offs), list(sizes)

    def _tensor_info(self, which: int):
        out = []
        buf = C.create_string_buffer(256)
        rows, cols = C.c_int64(), C.c_int64()
        for i in range(self.lib.tpc_num_tensors(self.m, which)):
            L.check(self.lib.tpc_param_info(self.m, which, i, buf, 256, C.byref(rows), C.byref(cols)), "tpc_param_info")
            out.append((buf.value.decode(), (rows.value, cols.value) if cols.value else (rows.value,)))
        return out

    def tensor_names(self, which: int):
        return [n for n, _ in self._tensor_info(which)]

    def tensor_shapes(self, which: int):
        return [s for _, s in self._tensor_info(which)]

    def set_params(self, which: int, flat) -> None:
        t = torch.as_tensor(np.asarray(flat, dtype=np.float32)) if not torch.is_tensor(flat) else flat
        if t.numel() != self.n_params[which]:
            raise ValueError(f"expected {self.n_params[which]} parameters, got {t.numel()}")
        self.params[which].copy_(t.to(self.device, dtype=torch.float32).reshape(-1))
        self.prepare(which)

    def prepare(self, which: int) -> None:
        L.check(self.lib.tpc_prepare_weights(self.m, which, L.ptr(self.params[which]), L.ptr(self.shadow[which]),
                                             self.stream), "tpc_prepare_weights")

    # ------------------------------------------------------------------ model forward (API-exact)
    def actor_forward(self, state, dec_input, ptr_mask, mha_mask, num_bins: int):
        B, S = state.shape[0], state.shape[1]
        R = S - num_bins
        ws = self.workspace(self.lib.tpc_workspace_bytes(self.m, 0, B, num_bins, R))
        logits = torch.empty(B, S, dtype=torch.float32, device=self.device)
        probs = torch.empty_like(logits)
        idx = torch.empty(B, dtype=torch.int32, device=self.device)
        L.check(self.lib.tpc_actor_forward(self.m, L.ptr(self.shadow[0]), L.ptr(self.params[0]), L.ptr(state),
                                           L.ptr(dec_input), L.ptr(ptr_mask), L.ptr(mha_mask), B, num_bins, R,
                                           L.ptr(logits), L.ptr(probs), L.ptr(idx), L.ptr(ws), ws.numel(),
                                           self.stream), "tpc_actor_forward")
        return logits, probs, idx

    def critic_forward(self, state, mha_mask, num_bins: int):
        n, S = state.shape[0], state.shape[1]
        R = S - num_bins
        ws = self.workspace(self.lib.tpc_workspace_bytes(self.m, 1, n, num_bins, R))
        values = torch.empty(n, dtype=torch.float32, device=self.device)
        L.check(self.lib.tpc_critic_forward(self.m, L.ptr(self.shadow[1]), L.ptr(self.params[1]), L.ptr(state),
                                            L.ptr(mha_mask), n, num_bins, R, L.ptr(values), L.ptr(ws), ws.numel(),
                                            self.stream), "tpc_critic_forward")
        return values

    # ------------------------------------------------------------------ fused path
    def rollout(self, env_cfg: L.EnvConfig, bufs: "RolloutTensors", greedy: bool, seed: int, episode_id0: int,
                epoch: int):
        B, N, R = bufs.B, bufs.N, bufs.R
        ws = self.workspace(self.lib.tpc_workspace_bytes(self.m, 4, B, N, R))
        rb = bufs.struct()
        L.check(self.lib.tpc_rollout(self.m, C.byref(env_cfg), L.ptr(self.shadow[0]), L.ptr(self.params[0]),
                                     C.byref(rb), B, N, R, 1 if greedy else 0, seed, episode_id0, epoch, L.ptr(ws),
                                     ws.numel(), self.stream), "tpc_rollout")

    def a2c_grads(self, bufs: "RolloutTensors", gamma: float, c_v: float, c_e: float, chunk_states: int = 0,
                  total_states: int = 0, phases: int = 0):
        """Zeroes and fills self.grads[0/1]; returns (values, advantages) of the local T*B states.
        phases: 0 = both tapes; L.A2C_CRITIC, then L.A2C_ACTOR, run them as two calls (the second one reuses the
        values / advantages of the first) so that a gradient all-reduce can start in between."""
        B, N, R = bufs.B, bufs.N, bufs.R
        n = B * R
        hp = L.A2CHyper(gamma, c_v, c_e, chunk_states, total_states or n, phases)
        ws = self.workspace(self.lib.tpc_a2c_workspace_bytes(self.m, B, N, R, chunk_states))
        if phases != L.A2C_ACTOR:
            for t in (self.grads[0], self.grads[1], self.losses):
                L.check(self.lib.tpc_zero(self.h, L.ptr(t), t.numel() * 4, self.stream), "tpc_zero")
            self._values = torch.empty(n, dtype=torch.float32, device=self.device)
            self._adv = torch.empty(n, dtype=torch.float32, device=self.device)
        values, adv = self._values, self._adv
        rb = bufs.struct()
        L.check(self.lib.tpc_a2c_grads(self.m, L.ptr(self.shadow[0]), L.ptr(self.params[0]), L.ptr(self.shadow[1]),
                                       L.ptr(self.params[1]), C.byref(rb), B, N, R, C.byref(hp),
                                       L.ptr(self.grads[0]), L.ptr(self.grads[1]), L.ptr(values), L.ptr(adv),
                                       L.ptr(self.losses), L.ptr(ws), ws.numel(), self.stream), "tpc_a2c_grads")
        return values, adv

    def adam_apply(self, which: int, lr: float, clipnorm, grad_scale: float = 1.0):
        self.adam_step[which] += 1
        L.check(self.lib.tpc_adam_apply(self.m, which, L.ptr(self.params[which]), L.ptr(self.grads[which]),
                                        L.ptr(self.adam_m[which]), L.ptr(self.adam_v[which]), self.adam_step[which],
                                        float(lr), float(clipnorm) if clipnorm else 0.0, float(grad_scale),
                                        self.stream), "tpc_adam_apply")
        self.prepare(which)

    def close(self):
        if getattr(self, "m", None):
            self.lib.tpc_model_destroy(self.m)
            self.m = None
        if getattr(self, "h", None):
            self.lib.tpc_destroy(self.h)
            self.h = None


class RolloutTensors:
    """Device buffers of one rollout: the live env state + the trajectory memory (Agent.store, agent.py:75-95)."""

    def __init__(self, B: int, N: int, R: int, F: int, device, cost: bool = False, store: bool = True,
                 keep_logits: bool = False, T: int | None = None):
        self.B, self.N, self.R, self.F = B, N, R, F
        S = N + R
        T = R if T is None else T      # steps per episode (CVRP adapter: N - 1 + R)
        self.T = T
        f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=device)
        self.batch, self.bin_net_mask, self.mha_mask = f(B, S, F), f(B, S), f(B, S)
        self.is_empty = f(B, S) if cost else None
        self.states = f(T, B, S, F) if store else None
        self.is_empties = f(T, B, S) if (cost and store) else None
        self.dec_inputs, self.ptr_masks, self.mha_masks = f(T, B, F), f(T, B, S), f(T, B, S)
        self.actions = torch.zeros(T, B, dtype=torch.int32, device=device)
        self.rewards = f(B, T)
        self.logits = f(T, B, S) if keep_logits else None

    def struct(self) -> L.RolloutBuffers:
        rb = L.RolloutBuffers()
        for name, _ in L.RolloutBuffers._fields_:
            t = getattr(self, name)
            setattr(rb, name, t.data_ptr() if t is not None else None)
        return rb
