"""CPU-side checks of the product package: the C-ABI library loads and exports every symbol include/tpc.h
declares (no compute calls without a GPU), config plumbing, and that the product never imports the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tpc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tpc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from tpc_b200 import build, lib as L
    path = build.build()
    so = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 20
    missing = [s for s in declared if not hasattr(so, s)]
    assert not missing, missing
    assert sorted(L.SIGNATURES) == declared, set(L.SIGNATURES) ^ set(declared)


def test_struct_layouts_match_header():
    from tpc_b200 import lib as L
    assert ctypes.sizeof(L.ModelConfig) == 16 * 4
    assert ctypes.sizeof(L.EnvConfig) == 15 * 4
    assert ctypes.sizeof(L.RolloutBuffers) == 12 * 8
    assert ctypes.sizeof(L.A2CHyper) == 6 * 4


def test_get_configs_matches_reference_signature():
    from tpc_b200.configs import get_configs
    agent, trainer, env, tester, tuner, everything = get_configs("ResourceV3", "tpc")
    assert agent["actor"]["num_layers"] == 1 and agent["critic"]["inner_layer_dim"] == 512
    assert env["node_sample_size"] == 10 and env["profiles_sample_size"] == 20 and env["batch_size"] == 128
    assert trainer["n_steps_to_update"] == 30 and tuner == {}
    assert get_configs("ResourceV3", "no_such_agent") is None      # configs/configs.py:13-15 prints and returns None
    agent2, *_ = get_configs("KnapsackV2", "tpc")
    assert agent2["actor"]["clipnorm"] is None


def test_model_and_env_config_translation():
    from tpc_b200.configs import get_configs
    from tpc_b200.engine import env_config_struct, model_config_from_agent_config
    agent, _, env, _, _, _ = get_configs("ResourceV3", "tpc")
    agent.update(num_bins=11, num_resources=20, tensor_size=31, batch_size=128,
                 encoder_embedding={"common": True, "num_bin_features": None, "num_resource_features": None})
    mc = model_config_from_agent_config(agent, 3)
    assert (mc.actor_num_layers, mc.critic_num_layers, mc.dim_model, mc.num_heads) == (1, 3, 128, 8)
    assert mc.has_logit_clipping == 1 and mc.logit_clipping_C == 10.0 and mc.common_embedding == 1
    e = env_config_struct(env, 0)
    assert (e.reward_type, e.num_profiles, e.node_max_val, e.req_max_val, e.mask_nodes_in_mha) == (0, 1000, 100, 30, 1)
    env["reward"]["type"] = "reduced_node_usage"
    e = env_config_struct(env, 0)
    assert e.reward_type == 3 and e.rejection_penalty == -2.0 and e.use_new_node_penalty == -1.0


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "transformer-pointer-critic_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tpc_b200 import lib as L
    from tpc_b200.engine import Engine
    with pytest.raises(L.TpcError):
        Engine(L.ModelConfig(), "cuda:0")


def test_training_log_csv_matches_reference_layout(tmp_path):
    """agents/plotter.py:99-126."""
    from tpc_b200.main import log_training_stats
    data = ([1.0, 2.5], [0.0, 1.0], [3.0, 4.0], [10.12345, 9.0], [0.5, 0.25], [0.75, 0.5], [1.6, 1.5])
    log_training_stats(data, str(tmp_path), "logs")
    lines = open(tmp_path / "logs.csv").read().splitlines()
    assert lines[0] == "Step;Avg Reward;Max Reward;Min Reward;Value Loss;Bin Entropy;Total Bin Loss;Bin Policy Loss"
    assert lines[1] == "0;1.000;3.000;0.000;10.123;1.600;0.750;0.500"
    assert lines[2] == "1;2.500;4.000;1.000;9.000;1.500;0.500;0.250"


def test_tester_csv_writers_match_reference_layout(tmp_path):
    """resource_v3/misc/csv_writer.py:3-35 and knapsack_v2/misc/csv_writer.py:5-36: header built from the stats keys
    ('_' -> ' '), one ';'-terminated row per instance, dominant / reward / rejected value with three decimals."""
    import numpy as np
    from tpc_b200.tester import knapsack_log_testing_stats, log_testing_stats
    inst = [{"net_dominant": np.float32([0.1, 0.25]), "net_rejected": np.int32([1, 0]),
             "net_empty_nodes": np.int32([2, 3]), "net_optimal": 0},
            {"random_dominant": np.float32([0.0, 0.5]), "random_rejected": np.int32([4, 5]),
             "random_empty_nodes": np.int32([0, 1]), "random_optimal": 0}]
    stats = [{"test_instance": 3, "node_sample_size": 5, "node_min_value": 0, "node_max_value": 100,
              "resource_sample_size": 10, "instance": inst}]
    log_testing_stats(stats, str(tmp_path), "t")
    lines = open(tmp_path / "t.csv").read().splitlines()
    assert lines[0] == ("test_instance;node_sample_size;node_min_value;node_max_value;resource_sample_size;"
                        "net dominant;net rejected;net empty nodes;net optimal;"
                        "random dominant;random rejected;random empty nodes;random optimal;")
    assert lines[1] == "6;5;0;100;10;0.100;1;2;0;0.000;4;0;0;"
    assert lines[2] == "7;5;0;100;10;0.250;0;3;0;0.500;5;1;0;"
    kinst = [{"net_reward": np.float32([1.5]), "net_empty_bins": np.int32([0]), "net_num_rejected_items": np.int32([2]),
              "net_rejected_value": np.float32([0.4])}]
    kstats = [{"test_instance": 0, "bin_sample_size": 5, "bin_min_value": 0, "bin_max_value": 100,
               "item_sample_size": 10, "instance": kinst}]
    knapsack_log_testing_stats(kstats, str(tmp_path), "k")
    klines = open(tmp_path / "k.csv").read().splitlines()
    assert klines[0] == ("test_instance;bin_sample_size;bin_min_value;bin_max_value;item_sample_size;"
                         "net reward;net empty bins;net num rejected items;net rejected value;")
    assert klines[1] == "0;5;0;100;10;1.500;0;2;0.400;"


def test_heuristic_factories_follow_reference_order_and_names():
    """heuristic/factory.py:16-71 (ResourceV3) and knapsack_v2/heuristic/factory.py:5-50: construction needs no GPU."""
    from tpc_b200.heuristics import heuristic_factory, knapsack_heuristic_factory
    opts = {"dominant_resource": {"generate_params_combos": True, "resource_sort_descending": True,
                                  "node_sort_descending": True}, "random": {},
            "cplex_greedy_and_critical": {"use": False}, "cplex_node_reduction": {"use": False}}
    names = [s.name for s in heuristic_factory(6, 100, opts, "cpu")]
    assert names == ["dominant_resource_ASC_True_node_ASC_True", "dominant_resource_ASC_True_node_ASC_False",
                     "dominant_resource_ASC_False_node_ASC_True", "dominant_resource_ASC_False_node_ASC_False", "random"]
    opts["dominant_resource"]["generate_params_combos"] = False
    assert [s.name for s in heuristic_factory(6, 100, opts, "cpu")] == ["dominant_resource_ASC_True_node_ASC_True", "random"]
    kopts = {"waste_reduction": {"generate_params_combos": True, "item_sort_descending": True,
                                 "bin_sort_descending": False}, "random": {}, "or_tools": {}}
    assert [s.name for s in knapsack_heuristic_factory(4, kopts, "cpu")] == \
        ["item_ASC_True_bin_ASC_True", "item_ASC_True_bin_ASC_False", "item_ASC_False_bin_ASC_True",
         "item_ASC_False_bin_ASC_False", "random"]
    # without a CUDA device the solvers refuse to run instead of falling back to the CPU
    import numpy as np
    import pytest
    import torch
    if not torch.cuda.is_available():
        from tpc_b200.lib import TpcError
        with pytest.raises((TpcError, RuntimeError, AssertionError)):
            heuristic_factory(3, 100, opts, "cpu")[0].solve(np.zeros((1, 5, 3), dtype=np.float32))
