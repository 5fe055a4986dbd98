"""Generate golden trajectories by executing the UNMODIFIED reference environment code.

Run in the build container only (needs /root/reference):

    PYTHONPATH=oracle/tf_shim:/root/reference python tests/golden/make_env_golden.py

The reference env (environment/custom/resource_v3/env.py, reward.py, misc/utils.py) is imported as is; its
TensorFlow calls resolve to the NumPy shim in oracle/tf_shim (TensorFlow is not installable here). Each case
replays a seeded random feasible policy and records every tensor that crosses the env API so that the oracle
(oracle/env.py) and the CUDA kernels can be checked step by step, bit for bit.
Outputs: tests/golden/env_<case>.npz
"""
import json
import zlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
sys.path.insert(0, "/root/reference")

import tensorflow as tf  # noqa: E402  (the shim)
from environment.custom.resource_v3.env import ResourceEnvironmentV3  # noqa: E402
from environment.custom.knapsack_v2.env import KnapsackEnvironmentV2  # noqa: E402


def run_case(env, rng, name, eos_bias=0.15):
    state, dec_input, bin_net_mask, mha_mask = env.reset()
    rec = {k: [] for k in ("state", "dec_input", "bin_net_mask", "feasible_mask", "mha_mask", "action", "reward",
                           "next_state", "next_bin_net_mask", "next_mha_mask", "next_resource_net_mask", "done")}
    done = False
    steps = 0
    while not done:
        feas = env.build_feasible_mask(state, dec_input, bin_net_mask)
        feas = np.asarray(feas)
        B, S = feas.shape
        actions = np.zeros(B, dtype=np.int32)
        for b in range(B):
            free = np.flatnonzero(feas[b] == 0)
            real = free[free != 0]
            if len(real) == 0 or rng.random() < eos_bias:
                actions[b] = 0
            else:
                actions[b] = rng.choice(real)
        rec["state"].append(np.array(state, dtype=np.float32))
        rec["dec_input"].append(np.array(dec_input, dtype=np.float32))
        rec["bin_net_mask"].append(np.array(bin_net_mask, dtype=np.float32))
        rec["feasible_mask"].append(np.array(feas, dtype=np.float32))
        rec["mha_mask"].append(np.array(mha_mask, dtype=np.float32))
        rec["action"].append(actions.copy())
        next_state, next_dec, reward, done, info = env.step(actions, feas)
        rec["reward"].append(np.asarray(reward, dtype=np.float32).reshape(B))
        rec["next_state"].append(np.array(next_state, dtype=np.float32))
        rec["next_bin_net_mask"].append(np.array(info["bin_net_mask"], dtype=np.float32))
        rec["next_mha_mask"].append(np.array(info["mha_used_mask"], dtype=np.float32))
        rec["next_resource_net_mask"].append(np.array(info["resource_net_mask"], dtype=np.float32))
        rec["done"].append(bool(done))
        state, bin_net_mask, mha_mask = next_state, info["bin_net_mask"], info["mha_used_mask"]
        if not done:
            dec_input = next_dec.copy()
        steps += 1
        assert steps <= 1000
    out = {k: np.stack(v) for k, v in rec.items()}
    path = os.path.join(HERE, f"env_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {steps} steps, state {out['state'].shape}, reward sum {out['reward'].sum():.4f} -> {path}")


def resource_v3_cases():
    base = json.load(open("/root/reference/configs/ResourceV3.json"))["env_config"]
    cases = []
    for reward in ("greedy", "single_node_dominant", "global_dominant", "reduced_node_usage"):
        for mask_nodes in (True, False):
            cfg = json.loads(json.dumps(base))
            cfg.update(batch_size=6, node_sample_size=5, profiles_sample_size=9, mask_nodes_in_mha=mask_nodes,
                       node_min_val=0, node_max_val=60)
            cfg["reward"]["type"] = reward
            cases.append((f"rv3_{reward}_{'mha' if mask_nodes else 'nomha'}", cfg))
    cfg = json.loads(json.dumps(base))
    cfg.update(batch_size=16)  # shipped shape: 10 nodes x 20 rules
    cases.append(("rv3_shipped_greedy", cfg))
    cfg = json.loads(json.dumps(base))
    cfg.update(batch_size=4, generate_request_on_the_fly=True, node_sample_size=7, profiles_sample_size=12)
    cfg["reward"]["type"] = "reduced_node_usage"
    cases.append(("rv3_onthefly_cost", cfg))
    return cases


def knapsack_v2_cases():
    base = json.load(open("/root/reference/configs/KnapsackV2.json"))["env_config"]
    cases = []
    for mask_nodes in (True, False):
        cfg = json.loads(json.dumps(base))
        cfg.update(batch_size=6, bin_sample_size=4, item_sample_size=10, mask_nodes_in_mha=mask_nodes)
        cases.append((f"kv2_{'mha' if mask_nodes else 'nomha'}", cfg))
    return cases


def main():
    for name, cfg in resource_v3_cases():
        tf.random.set_seed(zlib.crc32(name.encode()) % 10000 + 7)
        rng = np.random.default_rng(zlib.crc32(name.encode()))
        env = ResourceEnvironmentV3("ResourceV3", cfg)
        run_case(env, rng, name)
        with open(os.path.join(HERE, f"env_{name}.json"), "w") as f:
            json.dump(cfg, f, indent=1)
    try:
        cases = knapsack_v2_cases()
    except Exception as e:  # pragma: no cover
        print("knapsack cases skipped:", e)
        cases = []
    for name, cfg in cases:
        tf.random.set_seed(11)
        rng = np.random.default_rng(5)
        env = KnapsackEnvironmentV2("KnapsackV2", cfg)
        run_case(env, rng, name, eos_bias=0.1)
        with open(os.path.join(HERE, f"env_{name}.json"), "w") as f:
            json.dump(cfg, f, indent=1)


if __name__ == "__main__":
    main()
