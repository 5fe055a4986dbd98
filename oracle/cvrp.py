"""CPU oracle (NumPy restatement) of the reference CVRP environment, environment/custom/vrp/env.py, plus the
single-pointer adapter this repository drives it with.

TEST INFRASTRUCTURE ONLY (see oracle/env.py header). PINNED: tests/golden/cvrp_*.npz are trajectories recorded by
executing the UNMODIFIED reference class `CVRP` (tests/golden/make_cvrp_golden.py; the Augerat instance files it parses
stay in /root/reference, the goldens carry the sampled batches).

Reference semantics restated (file:line = environment/custom/vrp/env.py):
  state          (B, Nn + V, 3): rows [0, Nn) = nodes (row 0 the depot) as [x, y, demand], rows [Nn, Nn + V) = vehicles as
                 [x, y, remaining capacity]; raw integers stored as float32 (:148-206, :243-268)
  masks          backpack_net_mask: 1 on node rows and on vehicles that returned to the depot; item_net_mask: 1 on vehicle
                 rows, on visited nodes and on the depot until every node is visited; mha_used_mask (B,1,1,S) (:208-232)
  step(v, n)     reward = -round(sqrt(dx^2 + dy^2)) between vehicle v and node n (Python round of a float64 sqrt of
                 float32 operands); the vehicle moves to the node and its capacity drops by the demand; demand != 0 marks
                 the node visited (item mask + mha mask), demand == 0 (the depot) retires the vehicle (backpack mask + mha
                 mask); after Nn - 1 steps the depot becomes selectable; done after Nn - 1 + V steps (:67-146)
  feasible mask  max((capacity - demand) * (1 - backpack_net_mask) < 0, backpack_net_mask) (:270-301)

The reference Agent / trainer cannot drive this class (double-pointer `step`, 4-tuple `reset`).
Single-pointer adapter (documented in DESIGN.md): the nodes are served in sequence order -- step t < Nn - 1 serves node
t + 1, the last V steps send a vehicle back to the depot (node 0) -- and the POINTER chooses the vehicle: decoder input =
the node row, pointer mask = the reference's feasible mask. Every adapter step is one reference `step(vehicle, node)`."""
from __future__ import annotations

import math

import numpy as np


class CvrpOracle:
    def __init__(self, num_nodes: int, num_vehicles: int, num_features: int = 3):
        self.node_sample_size, self.vehicle_sample_size, self.num_features = num_nodes, num_vehicles, num_features

    @property
    def steps(self) -> int:
        return self.node_sample_size - 1 + self.vehicle_sample_size

    def load(self, batch):
        """Install a (B, Nn + V, 3) instance batch; masks as generate_masks (:208-232)."""
        self.batch = np.array(batch, dtype=np.float32)
        B, S = self.batch.shape[0], self.batch.shape[1]
        Nn = self.node_sample_size
        assert S == Nn + self.vehicle_sample_size
        self.batch_size = B
        self.visited_nodes = 0
        self.backpack_net_mask = np.zeros((B, S), np.float32)
        self.backpack_net_mask[:, :Nn] = 1
        self.item_net_mask = np.ones((B, S), np.float32) - self.backpack_net_mask
        self.item_net_mask[:, 0] = 1
        self.mha_used_mask = np.zeros((B, 1, 1, S), np.float32)
        return self.state()

    def state(self):
        return (self.batch.copy(), self.backpack_net_mask.copy(), self.item_net_mask.copy(),
                self.mha_used_mask.copy())

    def request_node(self, t: int) -> int:
        """Adapter: the node served at step t."""
        return t + 1 if t < self.node_sample_size - 1 else 0

    def build_feasible_mask(self, state, items, backpack_net_mask):
        """:270-301. items (B, 3) = the node rows being served."""
        B = state.shape[0]
        item_net_mask = np.ones_like(backpack_net_mask) - backpack_net_mask
        weight = np.reshape(np.asarray(items, np.float32)[:, 2], (B, 1))
        totals = (state[:, :, 2] - weight) * item_net_mask
        return np.maximum((totals < 0).astype(np.float32), backpack_net_mask).astype(np.float32)

    def step(self, vehicle_ids, node_ids):
        """:67-146."""
        B = self.batch_size
        rewards = np.zeros((B, 1), np.float32)
        for b in range(B):
            v, n = int(vehicle_ids[b]), int(node_ids[b])
            veh, node = self.batch[b, v].copy(), self.batch[b, n].copy()
            xd, yd = veh[0] - node[0], veh[1] - node[1]                      # float32 scalars
            dist = round(math.sqrt(np.float32(xd * xd) + np.float32(yd * yd)))
            assert veh[2] - node[2] >= 0, f"Vehicle {v} is overloaded"
            self.batch[b, v, 0], self.batch[b, v, 1] = node[0], node[1]
            self.batch[b, v, 2] = veh[2] - node[2]
            if node[2] != 0:
                self.item_net_mask[b, n] = 1
                self.mha_used_mask[b, :, :, n] = 1
            else:
                self.backpack_net_mask[b, v] = 1
                self.mha_used_mask[b, :, :, v] = 1
            rewards[b, 0] = -1 * dist
        self.visited_nodes += 1
        if self.node_sample_size - 1 == self.visited_nodes:
            self.item_net_mask[:, 0] = 0
        done = self.visited_nodes == self.node_sample_size - 1 + self.vehicle_sample_size
        info = {"backpack_net_mask": self.backpack_net_mask.copy(), "item_net_mask": self.item_net_mask.copy(),
                "mha_used_mask": self.mha_used_mask.copy(), "num_items_to_place": self.node_sample_size}
        return self.batch.copy(), rewards, done, info
