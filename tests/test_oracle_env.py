"""Pin the NumPy env oracle to trajectories recorded from the unmodified reference code
(tests/golden/make_env_golden.py): every state / mask / reward / next-input tensor, bit for bit."""
import glob
import json
import os

import numpy as np
import pytest

from oracle.env import KnapsackV2Oracle, ResourceV3Oracle

CASES = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "env_*.npz")))


def _same(a, b, what):
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) or np.array_equal(a, b), what


@pytest.mark.parametrize("case", CASES)
def test_env_oracle_matches_reference_trajectory(case, golden_dir):
    g = np.load(os.path.join(golden_dir, f"env_{case}.npz"))
    cfg = json.load(open(os.path.join(golden_dir, f"env_{case}.json")))
    env = KnapsackV2Oracle(cfg) if case.startswith("kv2") else ResourceV3Oracle(cfg)
    T = g["state"].shape[0]
    state, dec, bin_mask, mha = env.load(g["state"][0])
    for t in range(T):
        _same(state, g["state"][t], f"state[{t}]")
        _same(dec, g["dec_input"][t], f"dec_input[{t}]")
        _same(bin_mask, g["bin_net_mask"][t], f"bin_net_mask[{t}]")
        _same(mha, g["mha_mask"][t], f"mha_mask[{t}]")
        feas = env.build_feasible_mask(state, dec, bin_mask)
        _same(feas, g["feasible_mask"][t], f"feasible[{t}]")
        nstate, ndec, reward, done, info = env.step(g["action"][t], feas)
        _same(np.asarray(reward, dtype=np.float32).reshape(-1), g["reward"][t], f"reward[{t}]")
        _same(nstate, g["next_state"][t], f"next_state[{t}]")
        _same(info["bin_net_mask"], g["next_bin_net_mask"][t], f"next_bin[{t}]")
        _same(info["mha_used_mask"], g["next_mha_mask"][t], f"next_mha[{t}]")
        _same(info["resource_net_mask"], g["next_resource_net_mask"][t], f"next_res[{t}]")
        assert done == bool(g["done"][t])
        state, bin_mask, mha = nstate, info["bin_net_mask"], info["mha_used_mask"]
        if not done:
            dec = ndec
    assert done
