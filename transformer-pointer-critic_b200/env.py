"""Device-resident ResourceV3 / KnapsackV2 environments with the reference's Python interface
(environment/custom/resource_v3/env.py:21-447, environment/custom/knapsack_v2/env.py).

State lives in HBM (torch-allocated); every transition is a libtpc kernel (csrc/env.cu). The step-wise methods
(`reset/state/step/build_feasible_mask`) return host NumPy arrays exactly like the reference so that code
written against it keeps working; the fused trainer / tester use `device_buffers()` and never leave the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import lib as L
from . import vrp_instances
from .engine import REWARD_TYPES, RolloutTensors, env_config_struct


class HostTensor(np.ndarray):
    """ndarray with the `.numpy()` accessor the reference's trainer calls on TF tensors (trainer.py:73-74)."""

    def numpy(self):
        return np.asarray(self)


def _host(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


class _DeviceEnvBase:
    KIND = 0

    def __init__(self, name: str, opts: dict, device="cuda:0", seed_stream: int = 0):
        self.name = name
        self.opts = opts
        self.device = torch.device(device)
        self.lib = L.load()
        self.h = C.c_void_p()
        L.check(self.lib.tpc_create(C.byref(self.h)), "tpc_create")
        self.gather_stats = False
        self.mask_nodes_in_mha = opts["mask_nodes_in_mha"]
        self.normalization_factor = opts["normalization_factor"]
        self.decimal_precision = opts["decimal_precision"]
        if self.decimal_precision != 2:
            # the kernels round with floor(x * 100 + 0.5) / 100 (misc/utils.py:19-22 at decimals = 2, every shipped config)
            raise NotImplementedError(f"decimal_precision={self.decimal_precision}: only 2 is implemented on device")
        self.batch_size = opts["batch_size"]
        self.num_features = opts["num_features"]
        self.EOS_CODE = opts["EOS_CODE"]
        self.EOS_BIN = np.full((1, self.num_features), self.EOS_CODE, dtype="float32")
        self.seed_value = int(opts.get("seed_value", 1235))
        self.episode_id0 = 0          # first global episode id of this process (data-parallel ranks differ)
        self.epoch = 0                # reset counter: a new Philox stream per reset
        self.history = []
        self._cfg = None
        self.profiles = None

    # -- properties the subclasses define: node_sample_size (incl. EOS), profiles_sample_size --------------
    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def tensor_size(self):
        return self.node_sample_size + self.profiles_sample_size

    @property
    def is_cost(self):
        return False

    def _cfg_struct(self):
        return self._cfg

    def _alloc(self):
        B, S, F = self.batch_size, self.tensor_size, self.num_features
        f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
        self.batch_d, self.bin_net_mask_d, self.mha_mask_d = f(B, S, F), f(B, S), f(B, S)
        self.resource_net_mask_d = f(B, S)
        self.is_empty_d = f(B, S) if self.is_cost else None

    def generate_dataset(self):
        """env.py:205-215 on device (Philox stream keyed by seed_value)."""
        e = self._cfg_struct()
        if e.num_profiles <= 0:
            return None
        prof = torch.empty(e.num_profiles, self.num_features, dtype=torch.float32, device=self.device)
        L.check(self.lib.tpc_env_generate_profiles(self.h, C.byref(e), self.seed_value, L.ptr(prof), self.stream))
        return prof

    def reset(self):
        """env.py:83-99."""
        self.decoding_step = self.node_sample_size
        self._alloc()
        e = self._cfg_struct()
        L.check(self.lib.tpc_env_reset(self.h, C.byref(e), self.seed_value, self.episode_id0, self.epoch,
                                       L.ptr(self.profiles), self.batch_size, self.node_sample_size,
                                       self.profiles_sample_size, L.ptr(self.batch_d), L.ptr(self.bin_net_mask_d),
                                       L.ptr(self.mha_mask_d), L.ptr(self.is_empty_d), self.stream), "tpc_env_reset")
        # initial resource mask = complement of the initial bin mask (env.py:265-288): node positions 1, rule positions 0
        self.resource_net_mask_d[:, :self.node_sample_size].fill_(1.0)
        self.resource_net_mask_d[:, self.node_sample_size:].fill_(0.0)
        self.epoch += 1
        if self.gather_stats:
            self.history = None  # per-instance Python history objects are replaced by device statistics (tester.py)
        return self.state()

    def load_batch(self, batch) -> None:
        """Install a given (B,S,F) instance tensor (tests / reproducing a stored problem)."""
        batch = torch.as_tensor(np.asarray(batch, dtype=np.float32)[:, :, : self.num_features])
        self.batch_size = batch.shape[0]
        self.decoding_step = self.node_sample_size
        self._alloc()
        self.batch_d.copy_(batch.to(self.device))
        self.bin_net_mask_d[:, self.node_sample_size:] = 1
        self.resource_net_mask_d[:, : self.node_sample_size] = 1
        if self.is_cost:
            self.is_empty_d[:, 0] = self.EOS_CODE

    def _state_host(self):
        batch = _host(self.batch_d)
        if self.is_cost:  # add_is_empty_dim (env.py:424-426)
            batch = np.concatenate([batch, _host(self.is_empty_d)[:, :, None]], axis=-1)
        return batch

    def state(self):
        """env.py:101-113 -> (state, decoder_input, bin_net_mask, mha_used_mask) as host arrays."""
        batch = self._state_host()
        dec = np.expand_dims(batch[:, self.decoding_step, : self.num_features], axis=1)
        return batch, dec, _host(self.bin_net_mask_d), _host(self.mha_mask_d)[:, None, None, :]

    def build_feasible_mask(self, state, resources, bin_net_mask):
        """env.py:387-422 on device; accepts host or device arrays, returns a host array."""
        st = torch.as_tensor(np.ascontiguousarray(state, dtype=np.float32)).to(self.device)
        dec = torch.as_tensor(np.ascontiguousarray(resources, dtype=np.float32)).to(self.device).reshape(st.shape[0], -1)
        bm = torch.as_tensor(np.ascontiguousarray(bin_net_mask, dtype=np.float32)).to(self.device)
        out = torch.empty_like(bm)
        e = self._cfg_struct()
        L.check(self.lib.tpc_env_feasible_mask(self.h, C.byref(e), L.ptr(st), st.shape[2], L.ptr(dec), L.ptr(bm),
                                               st.shape[0], st.shape[1], L.ptr(out), self.stream))
        return _host(out)

    def step(self, bin_ids, feasible_bin_mask):
        """env.py:115-203."""
        ids_h = np.asarray(bin_ids, dtype=np.int32).reshape(-1)
        B = self.batch_size
        if ids_h.shape[0] != B:
            raise ValueError(f"step() expects one bin id per episode ({B}), got {ids_h.shape[0]}")
        if ids_h.size and (ids_h.min() < -self.tensor_size or ids_h.max() >= self.tensor_size):
            raise IndexError(f"bin id out of bounds for axis 1 with size {self.tensor_size}")   # env.py:129 fancy index
        ids = torch.as_tensor(np.where(ids_h < 0, ids_h + self.tensor_size, ids_h).astype(np.int32)).to(self.device)
        reward = torch.empty(B, dtype=torch.float32, device=self.device)
        e = self._cfg_struct()
        L.check(self.lib.tpc_env_step(self.h, C.byref(e), L.ptr(self.batch_d), L.ptr(self.bin_net_mask_d),
                                      L.ptr(self.resource_net_mask_d), L.ptr(self.mha_mask_d), L.ptr(self.is_empty_d),
                                      L.ptr(ids), self.decoding_step, B, self.node_sample_size,
                                      self.profiles_sample_size, L.ptr(reward), self.stream), "tpc_env_step")
        res_mask = _host(self.resource_net_mask_d)
        is_done = bool(np.all(res_mask == 1))
        info = {"bin_net_mask": _host(self.bin_net_mask_d), "resource_net_mask": res_mask,
                "mha_used_mask": _host(self.mha_mask_d)[:, None, None, :]}
        self.decoding_step += 1
        batch = self._state_host()
        if self.decoding_step < self.tensor_size:
            dec = np.expand_dims(batch[:, self.decoding_step, : self.num_features], axis=1)
        else:
            dec = np.array([None])
        rewards = _host(reward).reshape(B, 1).view(HostTensor)
        return batch, dec, rewards, is_done, info

    # -- fused path ----------------------------------------------------------------------------------------
    def device_buffers(self, store=True, keep_logits=False) -> RolloutTensors:
        """Fresh trajectory buffers whose live-state tensors alias this env's device state."""
        bufs = RolloutTensors(self.batch_size, self.node_sample_size, self.profiles_sample_size, self.num_features,
                              self.device, cost=self.is_cost, store=store, keep_logits=keep_logits)
        return bufs

    def reset_into(self, bufs: RolloutTensors) -> None:
        """reset() writing straight into rollout buffers (no host copies)."""
        e = self._cfg_struct()
        L.check(self.lib.tpc_env_reset(self.h, C.byref(e), self.seed_value, self.episode_id0, self.epoch,
                                       L.ptr(self.profiles), bufs.B, bufs.N, bufs.R, L.ptr(bufs.batch),
                                       L.ptr(bufs.bin_net_mask), L.ptr(bufs.mha_mask), L.ptr(bufs.is_empty),
                                       self.stream), "tpc_env_reset")
        self.epoch += 1

    def store_dataset(self, location) -> None:
        """env.py:443-444."""
        np.savetxt(location, _host(self.profiles) if self.profiles is not None else np.zeros((0, self.num_features)))

    def load_dataset(self, location):
        """env.py:446-447."""
        self.profiles = torch.as_tensor(np.loadtxt(location).astype(np.float32)).to(self.device).contiguous()
        self._cfg.num_profiles = self.profiles.shape[0]


class ResourceEnvironmentV3(_DeviceEnvBase):
    """environment/custom/resource_v3/env.py:20-447."""
    KIND = 0

    def __init__(self, name: str, opts: dict, device="cuda:0"):
        super().__init__(name, opts, device)
        self.generate_request_on_the_fly = opts["generate_request_on_the_fly"]
        self.num_profiles = opts["num_profiles"]
        self.profiles_sample_size = opts["profiles_sample_size"]
        assert self.num_profiles >= self.profiles_sample_size, \
            "Resource sample size should be less than total number of resources"
        self.node_sample_size = opts["node_sample_size"] + 1  # + EOS node (env.py:48)
        self.req_min_val, self.req_max_val = opts["req_min_val"], opts["req_max_val"]
        self.node_min_val, self.node_max_val = opts["node_min_val"], opts["node_max_val"]
        if opts["reward"]["type"] not in REWARD_TYPES:
            raise NameError(f"Unknown Reward Name! Select one of {list(REWARD_TYPES.keys())}")
        self.reward_type = opts["reward"]["type"]
        self._cfg = env_config_struct(opts, 0)
        self.decoding_step = self.node_sample_size
        self.profiles = self.total_profiles = self.generate_dataset()
        self.reset()

    @property
    def is_cost(self):
        return self.reward_type == "reduced_node_usage"

    def add_stats_to_agent_config(self, agent_config: dict):
        """env.py:314-334."""
        agent_config["num_resources"] = self.profiles_sample_size
        agent_config["num_bins"] = self.node_sample_size
        agent_config["tensor_size"] = self.node_sample_size + self.profiles_sample_size
        agent_config["batch_size"] = self.batch_size
        agent_config["encoder_embedding"] = {}
        if self.is_cost:
            agent_config["encoder_embedding"].update(common=False, num_bin_features=4, num_resource_features=3)
        else:
            agent_config["encoder_embedding"].update(common=True, num_bin_features=None, num_resource_features=None)
        return agent_config

    def set_testing_mode(self, batch_size, node_sample_size, profiles_sample_size, node_min_val, node_max_val) -> None:
        """env.py:336-352."""
        self.gather_stats = True
        self.batch_size = batch_size
        self.node_min_val, self.node_max_val = node_min_val, node_max_val
        self._cfg.node_min_val, self._cfg.node_max_val = node_min_val, node_max_val
        self.node_sample_size = node_sample_size + 1
        self.profiles_sample_size = profiles_sample_size


class KnapsackEnvironmentV2(_DeviceEnvBase):
    """environment/custom/knapsack_v2/env.py (same agent-facing interface, 2 features)."""
    KIND = 1

    def __init__(self, name: str, opts: dict, device="cuda:0"):
        super().__init__(name, opts, device)
        self.item_sample_size = self.profiles_sample_size = opts["item_sample_size"]
        self.bin_sample_size = self.node_sample_size = opts["bin_sample_size"] + 1
        self.reward_type = "greedy"
        self._cfg = env_config_struct(opts, 1)
        self.decoding_step = self.node_sample_size
        self.profiles = self.total_items = self.generate_dataset()
        self.reset()

    def add_stats_to_agent_config(self, agent_config: dict):
        """knapsack_v2/env.py:318-332: always separate embeddings with 2 / 2 features."""
        agent_config["num_resources"] = self.item_sample_size
        agent_config["num_bins"] = self.bin_sample_size
        agent_config["tensor_size"] = self.bin_sample_size + self.item_sample_size
        agent_config["batch_size"] = self.batch_size
        agent_config["encoder_embedding"] = {"common": False, "num_bin_features": 2, "num_resource_features": 2}
        return agent_config

    def set_testing_mode(self, batch_size, bin_sample_size, item_sample_size, bin_min_capacity=None,
                         bin_max_capacity=None):
        """knapsack_v2/env.py:334-352."""
        self.gather_stats = True
        self.batch_size = batch_size
        if bin_min_capacity is not None and bin_max_capacity is not None:
            self.bin_min_capacity, self.bin_max_capacity = bin_min_capacity, bin_max_capacity
            self._cfg.node_min_val, self._cfg.node_max_val = bin_min_capacity, bin_max_capacity
        self.bin_sample_size = self.node_sample_size = bin_sample_size + 1
        self.item_sample_size = self.profiles_sample_size = item_sample_size


def env_factory(type: str, name: str, opts: dict, device="cuda:0"):
    """environment/env_factory.py:27-33 -> (env, tester)."""
    from .tester import knapsack_test, test as resource_tester
    table = {"ResourceV3": (ResourceEnvironmentV3, resource_tester), "KnapsackV2": (KnapsackEnvironmentV2, knapsack_test),
             "CVRP": (CVRP, None)}                      # environment/env_factory.py:22: CVRP has no tester
    if type != "custom" or name not in table:
        raise NameError(f"Unknown environment {type}/{name}")
    env_cls, tester = table[name]
    return env_cls(name, opts, device), tester


class CVRP:
    """environment/custom/vrp/env.py:14-301 on the device, with the reference's own (double-index) interface --
    `reset() -> (batch, backpack_net_mask, item_net_mask, mha_used_mask)`, `step(vehicle_ids, node_ids) -> (batch,
    rewards, isDone, info)`, `build_feasible_mask(state, items, backpack_net_mask)` -- and the single-pointer adapter the
    fused rollout uses (`device_buffers` / `reset_into` / `Agent.rollout`): step t serves node t + 1 (then the depot once
    per vehicle) and the pointer chooses the vehicle; every adapter step is one reference `step(vehicle, node)`.

    The reference samples its batch from one Augerat instance file (`dir_path` / `problem_name`, :27-45). When
    `dir_path` holds that instance (`<name>.vrp.json` as the reference keeps it, or the bare `<name>.vrp`; the files are
    not part of this repository) `reset` does the same: the depot, `node_sample_size - 1` distinct customers of the
    instance and `vehicle_sample_size` vehicles of the instance's capacity per episode (`vrp_instances.py`; host NumPy,
    seeded by (seed_value, first episode id, reset count)), and `optimum_value` is the file's optimum. Otherwise `reset`
    draws synthetic instances of the same form on the device (integer coordinates 0..99, demands 1..`max_demand`,
    `vehicle_capacity`). `load_batch` installs a given batch."""
    KIND = 2

    def __init__(self, name: str, opts: dict, device="cuda:0"):
        self.name, self.opts = name, opts
        self.device = torch.device(device)
        self.lib = L.load()
        self.h = C.c_void_p()
        L.check(self.lib.tpc_create(C.byref(self.h)), "tpc_create")
        self.batch_size = opts["batch_size"]
        self.node_sample_size = opts["node_sample_size"]          # incl. the depot (row 0)
        self.vehicle_sample_size = opts["vehicle_sample_size"]
        self.num_features = opts["num_features"]
        if self.num_features != 3:
            raise ValueError("CVRP rows are [x, y, demand | capacity]: num_features must be 3")
        self.max_demand = int(opts.get("max_demand", 24))
        self.vehicle_capacity = int(opts.get("vehicle_capacity", 100))
        self.seed_value = int(opts.get("seed_value", 1235))
        self.episode_id0, self.epoch, self.visited_nodes = 0, 0, 0
        cfg = L.EnvConfig()
        cfg.kind, cfg.num_features = self.KIND, 3
        self._cfg = cfg
        # instance file (vrp/env.py:27-45) when it can be found, else synthetic instances
        self.instance_name, self.optimum_value, self._tables = None, None, None
        found = vrp_instances.load_problem(opts.get("dir_path"), opts.get("problem_name")) if opts.get("problem_name") else None
        if found is not None:
            problem, solution = found
            self.instance_name, self.optimum_value = problem["name"], solution.get("cost")
            self.vehicle_capacity = int(problem["capacity"])
            self.total_demand = int(problem["total_demand"])
            self._tables = vrp_instances.instance_tables(problem, 3)
        self.reset()

    def _sample_instance_batch(self, batch_size=None) -> np.ndarray:
        rng = np.random.default_rng([self.seed_value, self.episode_id0, self.epoch])
        self.epoch += 1
        return vrp_instances.sample_batch(self._tables, batch_size or self.batch_size, self.node_sample_size,
                                          self.vehicle_sample_size, rng)

    # the names Agent.rollout / trainer-side code use for the two token segments
    @property
    def profiles_sample_size(self):
        return self.vehicle_sample_size

    @property
    def tensor_size(self):
        return self.node_sample_size + self.vehicle_sample_size

    @property
    def steps(self):
        return self.node_sample_size - 1 + self.vehicle_sample_size

    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _cfg_struct(self):
        return self._cfg

    def _alloc(self):
        B, S = self.batch_size, self.tensor_size
        f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
        self.batch_d, self.backpack_net_mask_d, self.item_net_mask_d, self.mha_mask_d = f(B, S, 3), f(B, S), f(B, S), f(B, S)

    def reset(self):
        """:50-65."""
        if self._tables is not None:
            return self.load_batch(self._sample_instance_batch())
        self.visited_nodes = 0
        self._alloc()
        L.check(self.lib.tpc_cvrp_reset(self.h, self.seed_value, self.episode_id0, self.epoch, self.batch_size,
                                        self.node_sample_size, self.vehicle_sample_size, self.max_demand,
                                        self.vehicle_capacity, L.ptr(self.batch_d), L.ptr(self.backpack_net_mask_d),
                                        L.ptr(self.item_net_mask_d), L.ptr(self.mha_mask_d), self.stream), "tpc_cvrp_reset")
        self.epoch += 1
        return self.state()

    def load_batch(self, batch):
        batch = torch.as_tensor(np.asarray(batch, dtype=np.float32))
        self.batch_size, self.visited_nodes = batch.shape[0], 0
        self._alloc()
        Nn = self.node_sample_size
        self.batch_d.copy_(batch.to(self.device))
        self.backpack_net_mask_d[:, :Nn] = 1
        self.item_net_mask_d[:, Nn:] = 1
        self.item_net_mask_d[:, 0] = 1
        return self.state()

    def state(self):
        """:61-65."""
        return (_host(self.batch_d), _host(self.backpack_net_mask_d), _host(self.item_net_mask_d),
                _host(self.mha_mask_d)[:, None, None, :])

    def request_node(self, t: int) -> int:
        return t + 1 if t < self.node_sample_size - 1 else 0

    def build_feasible_mask(self, state, items, backpack_net_mask):
        """:270-301."""
        st = torch.as_tensor(np.ascontiguousarray(state, dtype=np.float32)).to(self.device)
        it = torch.as_tensor(np.ascontiguousarray(items, dtype=np.float32)).to(self.device).reshape(st.shape[0], -1)
        bm = torch.as_tensor(np.ascontiguousarray(backpack_net_mask, dtype=np.float32)).to(self.device)
        out = torch.empty_like(bm)
        L.check(self.lib.tpc_cvrp_feasible_mask(self.h, L.ptr(st), L.ptr(it), L.ptr(bm), st.shape[0], st.shape[1],
                                                L.ptr(out), self.stream), "tpc_cvrp_feasible_mask")
        return _host(out)

    def step(self, vehicle_ids, node_ids):
        """:67-146."""
        B, S = self.batch_size, self.tensor_size
        ids = []
        for a in (vehicle_ids, node_ids):
            a = np.asarray(a, dtype=np.int32).reshape(-1)
            if a.shape[0] != B:
                raise ValueError(f"step() expects one id per episode ({B})")
            if a.size and (a.min() < -S or a.max() >= S):
                raise IndexError(f"index out of bounds for axis 1 with size {S}")
            ids.append(torch.as_tensor(np.where(a < 0, a + S, a).astype(np.int32)).to(self.device))
        state, _, _, _ = self.state()
        v, n = ids[0].cpu().numpy(), ids[1].cpu().numpy()
        if np.any(state[np.arange(B), v, 2] - state[np.arange(B), n, 2] < 0):   # the reference's assert (:97-98)
            raise AssertionError("Vehicle is overloaded")
        reward = torch.empty(B, dtype=torch.float32, device=self.device)
        L.check(self.lib.tpc_cvrp_step(self.h, L.ptr(self.batch_d), L.ptr(self.backpack_net_mask_d),
                                       L.ptr(self.item_net_mask_d), L.ptr(self.mha_mask_d), L.ptr(ids[0]), L.ptr(ids[1]),
                                       B, self.node_sample_size, self.vehicle_sample_size, self.visited_nodes,
                                       L.ptr(reward), self.stream), "tpc_cvrp_step")
        self.visited_nodes += 1
        is_done = self.visited_nodes == self.steps
        batch, bp, item, mha = self.state()
        info = {"backpack_net_mask": bp, "item_net_mask": item, "mha_used_mask": mha,
                "num_items_to_place": self.node_sample_size}
        return batch, _host(reward).reshape(B, 1).view(HostTensor), is_done, info

    def add_stats_to_agent_config(self, agent_config: dict):
        """:234-247 (`num_items`, `tensor_size`, `num_backpacks`, `batch_size`) plus the keys the single-pointer Agent
        reads (agents/agent.py:25-29, resource_v3/env.py:314-334): nodes are the first token segment, vehicles the second."""
        agent_config["num_items"] = self.steps
        agent_config["tensor_size"] = self.tensor_size
        agent_config["num_backpacks"] = self.vehicle_sample_size
        agent_config["batch_size"] = self.batch_size
        agent_config["num_bins"] = self.node_sample_size
        agent_config["num_resources"] = self.vehicle_sample_size
        agent_config["encoder_embedding"] = {"common": True, "num_bin_features": None, "num_resource_features": None}
        return agent_config

    # -- fused path (single-pointer adapter) ---------------------------------------------------------------
    def device_buffers(self, store=True, keep_logits=False) -> RolloutTensors:
        return RolloutTensors(self.batch_size, self.node_sample_size, self.vehicle_sample_size, 3, self.device,
                              store=store, keep_logits=keep_logits, T=self.steps)

    def reset_into(self, bufs: RolloutTensors) -> None:
        if self._tables is not None:     # instance file: the batch comes from the host, the masks as cvrp_reset_kernel sets them
            bufs.batch.copy_(torch.as_tensor(self._sample_instance_batch(bufs.B)))
            bufs.bin_net_mask.zero_()
            bufs.bin_net_mask[:, : bufs.N] = 1
            bufs.mha_mask.zero_()
            return
        L.check(self.lib.tpc_cvrp_reset(self.h, self.seed_value, self.episode_id0, self.epoch, bufs.B, bufs.N, bufs.R,
                                        self.max_demand, self.vehicle_capacity, L.ptr(bufs.batch),
                                        L.ptr(bufs.bin_net_mask), None, L.ptr(bufs.mha_mask), self.stream),
                "tpc_cvrp_reset")
        self.epoch += 1
