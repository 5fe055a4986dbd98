"""Generate golden CVRP trajectories by executing the UNMODIFIED reference class environment/custom/vrp/env.py::CVRP
(its Augerat parser and instance files included) on the NumPy TensorFlow shim.

    python tests/golden/make_cvrp_golden.py        # build container only (needs /root/reference)

The trajectory follows the single-pointer adapter of oracle/cvrp.py: step t serves node t + 1 (then the depot V times)
and a seeded policy picks a feasible vehicle from the reference's own `build_feasible_mask`; every adapter step is one
reference `step(vehicle_ids, node_ids)` call. Recorded per step: state, the three masks, the served node rows, the
feasible mask, the action pair, the reward, and the post-step state / masks."""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
sys.path.insert(0, "/root/reference")

from environment.custom.vrp.env import CVRP  # noqa: E402

CASES = [("cvrp_n12_v5", dict(node_sample_size=12, vehicle_sample_size=5, batch_size=6, problem_name="A-n32-k5")),
         ("cvrp_n20_v5", dict(node_sample_size=20, vehicle_sample_size=5, batch_size=4, problem_name="A-n33-k5")),
         ("cvrp_n8_v3", dict(node_sample_size=8, vehicle_sample_size=3, batch_size=5, problem_name="A-n32-k5"))]


def run_case(name, over, seed):
    cfg = {"load_from_file": False, "dir_path": "/root/reference/environment/custom/vrp/datasets", "num_features": 3}
    cfg.update(over)
    for attempt in range(50):
        np.random.seed(seed + attempt)
        rng = np.random.default_rng(seed + attempt)
        env = CVRP("CVRP", cfg)
        Nn, V, B = env.node_sample_size, env.vehicle_sample_size, env.batch_size
        T = Nn - 1 + V
        state, bp_mask, item_mask, mha = env.reset()
        rec = {k: [] for k in ("state", "backpack_net_mask", "item_net_mask", "mha_mask", "items", "feasible_mask",
                               "vehicle_ids", "node_ids", "reward", "next_state", "next_backpack_net_mask",
                               "next_item_net_mask", "next_mha_mask", "done")}
        ok = True
        for t in range(T):
            node = t + 1 if t < Nn - 1 else 0
            items = state[:, node].copy()
            feas = np.asarray(env.build_feasible_mask(state, items, bp_mask), dtype=np.float32)
            vids = np.zeros(B, np.int32)
            for b in range(B):
                free = np.flatnonzero(feas[b] == 0)
                if len(free) == 0:
                    ok = False
                    break
                vids[b] = rng.choice(free)
            if not ok:
                break
            nids = np.full(B, node, np.int32)
            for k, v in (("state", state), ("backpack_net_mask", bp_mask), ("item_net_mask", item_mask),
                         ("mha_mask", mha), ("items", items), ("feasible_mask", feas), ("vehicle_ids", vids),
                         ("node_ids", nids)):
                rec[k].append(np.array(v))
            state, reward, done, info = env.step(vids, nids)
            bp_mask, item_mask, mha = info["backpack_net_mask"], info["item_net_mask"], info["mha_used_mask"]
            for k, v in (("reward", reward.reshape(B)), ("next_state", state), ("next_backpack_net_mask", bp_mask),
                         ("next_item_net_mask", item_mask), ("next_mha_mask", mha), ("done", done)):
                rec[k].append(np.array(v))
        if ok:
            assert done
            out = {k: np.stack(v) for k, v in rec.items()}
            out["num_nodes"], out["num_vehicles"] = Nn, V
            path = os.path.join(HERE, f"{name}.npz")
            np.savez_compressed(path, **out)
            print(f"{name}: {T} steps, reward sum {out['reward'].sum():.0f} (attempt {attempt}) -> {path}")
            return
    raise SystemExit(f"{name}: no dead-end-free trajectory found")


if __name__ == "__main__":
    for name, over in CASES:
        run_case(name, over, zlib.crc32(name.encode()) % 10000)
