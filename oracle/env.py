"""CPU oracle (NumPy restatement) of the reference ResourceV3 / KnapsackV2 environments.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs as the *checker*; never by the product package.

Pinned: tests/test_oracle_env.py replays tests/golden/env_*.npz, which were produced by executing the
unmodified reference code (tests/golden/make_env_golden.py), plus the reference's own known-answer tests
(tests/unit/environment/resource_v3/{env,reward,utils}_test.py) restated in tests/test_reference_kats.py.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def round_half_up(n, decimals=0):
    """environment/custom/resource_v3/misc/utils.py:19-22 -- floor(n * 10^d + 0.5) / 10^d in the array's dtype."""
    multiplier = 10 ** decimals
    return np.floor(n * multiplier + 0.5) / multiplier


def compute_remaining_resources(nodes, reqs, decimal_precision=2):
    """resource_v3/misc/utils.py:36-42."""
    return round_half_up(nodes - reqs, decimal_precision)


def bins_eos_checker(bins, eos_symbol, num_features, dtype=np.int32):
    """resource_v3/misc/utils.py:24-34 -- 1 where every feature of the selected row equals the EOS code."""
    is_eos = (bins == eos_symbol).astype(dtype).sum(axis=-1)
    return (is_eos == num_features).astype(dtype)


def reshape_into_vertical_format(data, batch_size):
    """resource_v3/misc/utils.py:13-17 -- (B,T) -> (T*B,1), step-major."""
    return np.reshape(np.transpose(data, [1, 0]), [batch_size, 1])


def reshape_into_horizontal_format(data, batch_size, decoding_steps):
    """resource_v3/misc/utils.py:7-11."""
    return np.transpose(np.reshape(data, [decoding_steps, batch_size]), [1, 0])


# ------------------------------------------------------------------------------------------------------
# rewards: environment/custom/resource_v3/reward.py
# ------------------------------------------------------------------------------------------------------
class GreedyReward:
    """reward.py:23-44 (int32 result)."""

    def __init__(self, opts, eos_node):
        self.eos = eos_node

    def compute_reward(self, updated_batch, original_batch, total_num_nodes, nodes, reqs, feasible_mask, node_ids,
                       is_empty=None):
        is_eos = bins_eos_checker(nodes, self.eos[0], updated_batch.shape[2])
        return 1 - is_eos


class SingleNodeDominantReward:
    """reward.py:46-85 -- min over features of the UNROUNDED fp32 difference."""

    def __init__(self, opts, eos_node):
        self.eos = eos_node
        self.rejection_penalty = opts["rejection_penalty"]

    def compute_reward(self, updated_batch, original_batch, total_num_nodes, nodes, reqs, feasible_mask, node_ids,
                       is_empty=None):
        is_eos = bins_eos_checker(nodes, self.eos[0], updated_batch.shape[2], F32)
        penalties = is_eos * self.rejection_penalty
        dominant = np.min(nodes - reqs, axis=-1)
        return (dominant * (1 - is_eos) + penalties).astype(F32)


class GlobalDominantReward:
    """reward.py:88-127 -- min over real nodes and features of the updated state ("fair"/"critical")."""

    def __init__(self, opts, eos_node):
        self.eos = eos_node
        self.rejection_penalty = opts["rejection_penalty"]

    def compute_reward(self, updated_batch, original_batch, total_num_nodes, nodes, reqs, feasible_mask, node_ids,
                       is_empty=None):
        dominant = np.min(np.min(updated_batch[:, 1:total_num_nodes], axis=-1), axis=-1)
        is_eos = bins_eos_checker(nodes, self.eos[0], updated_batch.shape[2], F32)
        penalties = is_eos * self.rejection_penalty
        return (dominant * (1 - is_eos) + penalties).astype(F32)


class ReducedNodeUsage:
    """reward.py:130-166 -- "cost": penalty for opening a new node; mutates is_empty."""

    def __init__(self, opts, eos_node):
        self.eos = eos_node
        self.rejection_penalty = opts["rejection_penalty"]
        self.use_new_node_penalty = opts["use_new_node_penalty"]

    def compute_reward(self, updated_batch, original_batch, total_num_nodes, nodes, reqs, feasible_mask, node_ids,
                       is_empty):
        idx = np.arange(updated_batch.shape[0])
        is_eos = bins_eos_checker(nodes, self.eos[0], updated_batch.shape[2], F32)
        penalties = is_eos * self.rejection_penalty
        already_used = is_empty[idx, node_ids, 0].copy()
        is_empty[idx, node_ids, 0] = 1
        is_empty[:, 0, 0] = self.eos[0][0]
        return (((1 - already_used) * self.use_new_node_penalty) * (1 - is_eos) + penalties).astype(F32)


def reward_factory(opts, eos_node):
    """reward.py:6-21 (gini is broken in the reference, SURVEY E9, and not restated)."""
    table = {"greedy": GreedyReward, "single_node_dominant": SingleNodeDominantReward,
             "global_dominant": GlobalDominantReward, "reduced_node_usage": ReducedNodeUsage}
    try:
        return table[opts["type"]](opts[opts["type"]], eos_node)
    except KeyError:
        raise NameError(f"Unknown Reward Name! Select one of {list(table.keys())}")


# ------------------------------------------------------------------------------------------------------
# environment: environment/custom/resource_v3/env.py
# ------------------------------------------------------------------------------------------------------
class ResourceV3Oracle:
    """State machine of ResourceEnvironmentV3 (env.py:21-447) with instance generation factored out:
    `load(batch)` installs a given (B,S,3) instance tensor (the reference samples it with TF's RNG,
    env.py:217-263, which cannot be reproduced without TF)."""

    def __init__(self, opts: dict):
        self.mask_nodes_in_mha = opts["mask_nodes_in_mha"]
        self.decimal_precision = opts["decimal_precision"]
        self.batch_size = opts["batch_size"]
        self.num_features = opts["num_features"]
        self.profiles_sample_size = opts["profiles_sample_size"]
        self.EOS_CODE = opts["EOS_CODE"]
        self.EOS_BIN = np.full((1, self.num_features), self.EOS_CODE, dtype=F32)
        self.node_sample_size = opts["node_sample_size"] + 1          # env.py:48
        self.rewarder = reward_factory(opts["reward"], self.EOS_BIN)  # env.py:59-62
        self.is_cost = isinstance(self.rewarder, ReducedNodeUsage)
        self.decoding_step = self.node_sample_size
        self.batch = None
        self.is_empty = None

    @property
    def tensor_size(self):
        return self.node_sample_size + self.profiles_sample_size

    def load(self, batch):
        """env.py:83-99 (reset) given the sampled batch."""
        batch = np.array(batch, dtype=F32)[:, :, : self.num_features].copy()
        self.batch_size = batch.shape[0]
        assert batch.shape[1] == self.tensor_size
        self.decoding_step = self.node_sample_size
        if self.is_cost:  # env.py:87-91
            self.is_empty = np.zeros((self.batch_size, self.tensor_size, 1), dtype=F32)
            self.is_empty[:, 0, 0] = self.EOS_BIN[0][0]
        self.batch = batch
        self.bin_net_mask, self.resource_net_mask, self.mha_used_mask = self.generate_masks()
        return self.state()

    def generate_masks(self):
        """env.py:265-288."""
        S = self.tensor_size
        profiles_net_mask = np.zeros((self.batch_size, S), dtype=F32)
        nodes_net_mask = np.ones((self.batch_size, S), dtype=F32)
        profiles_net_mask[:, : self.node_sample_size] = 1
        nodes_net_mask = nodes_net_mask - profiles_net_mask
        mha_used_mask = np.zeros_like(profiles_net_mask)[:, np.newaxis, np.newaxis, :]
        return nodes_net_mask, profiles_net_mask, mha_used_mask

    def add_is_empty_dim(self, batch, is_empty):
        """env.py:424-426."""
        return round_half_up(np.concatenate([batch, is_empty], axis=-1), 2)

    def remove_is_empty_dim(self, batch):
        """env.py:428-430."""
        return round_half_up(batch[:, :, : self.num_features], 2)

    def state(self):
        """env.py:101-113."""
        dec = np.expand_dims(self.batch[:, self.decoding_step], axis=1)
        batch = self.batch.copy()
        if self.is_cost:
            batch = self.add_is_empty_dim(batch, self.is_empty)
        return batch, dec, self.bin_net_mask.copy(), self.mha_used_mask.copy()

    def build_feasible_mask(self, state, resources, bin_net_mask):
        """env.py:387-422."""
        if self.is_cost:
            state = self.remove_is_empty_dim(state)
        num_elems = state.shape[1]
        demands = np.tile(resources, [1, num_elems, 1])
        remaining = compute_remaining_resources(state, demands, self.decimal_precision)
        dominant = np.min(remaining, axis=-1)
        after_place = (dominant < 0).astype(F32)
        feasible_mask = np.maximum(after_place, bin_net_mask).astype(F32)
        assert np.all(dominant * (1 - feasible_mask) >= 0), "Masking Scheme Is Wrong!"
        feasible_mask[:, 0] = 0
        return feasible_mask

    def step(self, bin_ids, feasible_bin_mask):
        """env.py:115-203."""
        bin_ids = np.asarray(bin_ids)
        B = self.batch.shape[0]
        idx = np.arange(B)
        req_ids = np.full(B, self.decoding_step)
        copy_batch = self.batch.copy()
        nodes = self.batch[idx, bin_ids]
        reqs = self.batch[idx, req_ids]
        remaining = compute_remaining_resources(nodes, reqs, self.decimal_precision)
        self.batch[idx, bin_ids] = remaining
        self.batch[idx, 0] = self.EOS_BIN                       # :139
        self.resource_net_mask[idx, req_ids] = 1                # :142
        is_full = (np.min(remaining, axis=-1) == 0).astype(F32)  # :145-146
        self.bin_net_mask[idx, bin_ids] = is_full               # :148
        self.bin_net_mask[:, 0] = 0
        self.mha_used_mask[idx, :, :, req_ids] = 1              # :152
        if self.mask_nodes_in_mha:
            self.mha_used_mask[idx, :, :, bin_ids] = is_full.reshape(B, 1, 1)
        self.mha_used_mask[idx, :, :, 0] = 0                    # :159
        done = bool(np.all(self.resource_net_mask == 1))
        rewards = self.rewarder.compute_reward(self.batch, copy_batch, self.node_sample_size, nodes, reqs,
                                               feasible_bin_mask, bin_ids, self.is_empty)
        rewards = np.reshape(rewards, (B, 1))
        info = {"bin_net_mask": self.bin_net_mask.copy(), "resource_net_mask": self.resource_net_mask.copy(),
                "mha_used_mask": self.mha_used_mask.copy()}
        self.decoding_step += 1
        if self.decoding_step < self.tensor_size:
            dec = np.expand_dims(self.batch[:, self.decoding_step], axis=1)
        else:
            dec = np.array([None])
        batch = self.batch.copy()
        if self.is_cost:
            batch = self.add_is_empty_dim(batch, self.is_empty)
        return batch, dec, rewards, done, info


# ------------------------------------------------------------------------------------------------------
# KnapsackV2 variant: environment/custom/knapsack_v2/{env.py, reward.py, misc/utils.py}
# ------------------------------------------------------------------------------------------------------
class KnapsackV2Oracle:
    """KnapsackEnvironmentV2 (knapsack_v2/env.py:21-408): features [capacity, load] for bins and
    [weight, value] for items; load += weight (rounded), full iff rhu(cap-load) == 0, feasible iff
    cap - (load + w) >= 0 evaluated UNROUNDED in fp32 (env.py:384-408); reward = value * (1 - eos)."""

    def __init__(self, opts: dict):
        self.mask_nodes_in_mha = opts["mask_nodes_in_mha"]
        self.decimal_precision = opts["decimal_precision"]
        self.batch_size = opts["batch_size"]
        self.num_features = opts["num_features"]
        self.item_sample_size = opts["item_sample_size"]
        self.EOS_CODE = opts["EOS_CODE"]
        self.EOS_BIN = np.full((1, self.num_features), self.EOS_CODE, dtype=F32)
        self.bin_sample_size = opts["bin_sample_size"] + 1
        self.decoding_step = self.bin_sample_size
        self.batch = None

    @property
    def tensor_size(self):
        return self.bin_sample_size + self.item_sample_size

    def load(self, batch):
        self.batch = np.array(batch, dtype=F32).copy()
        self.batch_size = self.batch.shape[0]
        self.decoding_step = self.bin_sample_size
        S = self.tensor_size
        item_mask = np.zeros((self.batch_size, S), dtype=F32)
        item_mask[:, : self.bin_sample_size] = 1
        self.bin_net_mask = np.ones((self.batch_size, S), dtype=F32) - item_mask
        self.item_net_mask = item_mask
        self.mha_used_mask = np.zeros((self.batch_size, 1, 1, S), dtype=F32)
        return self.state()

    def state(self):
        dec = np.expand_dims(self.batch[:, self.decoding_step], axis=1)
        return self.batch.copy(), dec, self.bin_net_mask.copy(), self.mha_used_mask.copy()

    def build_feasible_mask(self, state, items, bin_net_mask):
        """knapsack_v2/env.py:384-408."""
        num_elems = state.shape[1]
        item_weight = np.tile(items, [1, num_elems, 1])[:, :, 0]
        bin_capacity = state[:, :, 0]
        bin_load = state[:, :, 1]
        after = ((bin_capacity - (bin_load + item_weight)) < 0).astype(F32)
        feasible = np.maximum(after, bin_net_mask).astype(F32)
        feasible[:, 0] = 0
        return feasible

    def step(self, bin_ids, feasible_bin_mask):
        """knapsack_v2/env.py:102-185."""
        bin_ids = np.asarray(bin_ids)
        B = self.batch.shape[0]
        idx = np.arange(B)
        item_ids = np.full(B, self.decoding_step)
        bins = self.batch[idx, bin_ids]
        items = self.batch[idx, item_ids]
        # knapsack_v2/misc/utils.py:10-20: load += weight, then the whole row is rounded half-up
        updated = bins.copy()
        updated[:, 1] = updated[:, 1] + items[:, 0]
        updated = round_half_up(updated, self.decimal_precision)
        self.batch[idx, bin_ids] = updated
        self.batch[idx, 0] = self.EOS_BIN
        self.item_net_mask[idx, item_ids] = 1
        remaining = round_half_up(updated[:, 0] - updated[:, 1], self.decimal_precision)  # env.py:129-131
        is_full = (remaining == 0).astype(F32)
        self.bin_net_mask[idx, bin_ids] = is_full
        self.bin_net_mask[:, 0] = 0
        self.mha_used_mask[idx, :, :, item_ids] = 1
        if self.mask_nodes_in_mha:
            self.mha_used_mask[idx, :, :, bin_ids] = is_full.reshape(B, 1, 1)
        self.mha_used_mask[idx, :, :, 0] = 0
        done = bool(np.all(self.item_net_mask == 1))
        is_eos = bins_eos_checker(bins, self.EOS_BIN[0], self.num_features, F32)  # knapsack_v2/reward.py:17-41
        rewards = np.reshape((items[:, 1] * (1 - is_eos)).astype(F32), (B, 1))
        info = {"bin_net_mask": self.bin_net_mask.copy(), "resource_net_mask": self.item_net_mask.copy(),
                "mha_used_mask": self.mha_used_mask.copy()}
        self.decoding_step += 1
        if self.decoding_step < self.tensor_size:
            dec = np.expand_dims(self.batch[:, self.decoding_step], axis=1)
        else:
            dec = np.array([None])
        return self.batch.copy(), dec, rewards, done, info
