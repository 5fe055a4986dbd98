"""ctypes binding of libtpc.so (include/tpc.h). No compute fallback: if the library or a CUDA device is
missing every call fails loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TPC_LIB") or os.path.join(_HERE, "libtpc.so")   # TPC_LIB: A/B runs of two builds

c_int, c_float, c_void_p, c_size_t = C.c_int, C.c_float, C.c_void_p, C.c_size_t
c_u64, c_i64 = C.c_uint64, C.c_int64


class ModelConfig(C.Structure):
    _fields_ = [(n, c_int) for n in ("actor_num_layers", "actor_inner_layer_dim", "critic_num_layers",
                                     "critic_inner_layer_dim", "dim_model", "num_heads", "attention_dense_units",
                                     "last_layer_units")] + \
               [("logit_clipping_C", c_float), ("has_logit_clipping", c_int), ("common_embedding", c_int),
                ("num_bin_features", c_int), ("num_resource_features", c_int), ("dec_features", c_int),
                ("num_bins", c_int), ("num_resources", c_int)]


class EnvConfig(C.Structure):
    _fields_ = [("kind", c_int), ("reward_type", c_int), ("rejection_penalty", c_float),
                ("use_new_node_penalty", c_float), ("mask_nodes_in_mha", c_int), ("num_features", c_int),
                ("eos_code", c_float), ("normalization_factor", c_int), ("node_min_val", c_int),
                ("node_max_val", c_int), ("req_min_val", c_int), ("req_max_val", c_int), ("val_min_val", c_int),
                ("val_max_val", c_int), ("num_profiles", c_int)]


class RolloutBuffers(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("batch", "bin_net_mask", "mha_mask", "is_empty", "states", "is_empties",
                                        "dec_inputs", "ptr_masks", "mha_masks", "actions", "rewards", "logits")]


class A2CHyper(C.Structure):
    _fields_ = [("gamma", c_float), ("values_loss_coefficient", c_float), ("entropy_coefficient", c_float),
                ("chunk_states", c_int), ("total_states", c_int), ("phases", c_int)]


A2C_CRITIC, A2C_ACTOR = 1, 2


ERRORS = {0: "TPC_OK", -1: "TPC_ERR_CUDA", -2: "TPC_ERR_ARG", -3: "TPC_ERR_SHAPE", -4: "TPC_ERR_WORKSPACE",
          -5: "TPC_ERR_DRIVER"}

# name -> (restype, argtypes): every symbol include/tpc.h declares
SIGNATURES = {
    "tpc_create": (c_int, [C.POINTER(c_void_p)]),
    "tpc_destroy": (c_int, [c_void_p]),
    "tpc_num_sms": (c_int, [c_void_p]),
    "tpc_launch_count": (C.c_ulonglong, []),
    "tpc_rollout_graph_replays": (C.c_ulonglong, [C.c_void_p]),
    "tpc_model_create": (c_int, [c_void_p, C.POINTER(ModelConfig), C.POINTER(c_void_p)]),
    "tpc_model_destroy": (c_int, [c_void_p]),
    "tpc_param_count": (c_size_t, [c_void_p, c_int]),
    "tpc_num_tensors": (c_int, [c_void_p, c_int]),
    "tpc_param_layout": (c_int, [c_void_p, c_int, C.POINTER(c_i64), C.POINTER(c_i64)]),
    "tpc_param_info": (c_int, [c_void_p, c_int, c_int, C.c_char_p, c_int, C.POINTER(c_i64), C.POINTER(c_i64)]),
    "tpc_shadow_count": (c_size_t, [c_void_p, c_int]),
    "tpc_prepare_weights": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "tpc_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int, c_int]),
    "tpc_a2c_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int, c_int]),
    "tpc_gemm": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                         c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "tpc_gemm_ln_bwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tpc_gemm_split_image": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tpc_wgrad": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                          c_void_p]),
    "tpc_attention_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "tpc_attention_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                  c_void_p, c_void_p]),
    "tpc_actor_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                  c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tpc_critic_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                   c_void_p, c_size_t, c_void_p]),
    "tpc_env_generate_profiles": (c_int, [c_void_p, C.POINTER(EnvConfig), c_u64, c_void_p, c_void_p]),
    "tpc_env_reset": (c_int, [c_void_p, C.POINTER(EnvConfig), c_u64, c_i64, c_i64, c_void_p, c_int, c_int, c_int,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tpc_env_feasible_mask": (c_int, [c_void_p, C.POINTER(EnvConfig), c_void_p, c_int, c_void_p, c_void_p, c_int,
                                      c_int, c_void_p, c_void_p]),
    "tpc_env_step": (c_int, [c_void_p, C.POINTER(EnvConfig), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tpc_cvrp_reset": (c_int, [c_void_p, c_u64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p]),
    "tpc_cvrp_feasible_mask": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "tpc_cvrp_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                              c_int, c_void_p, c_void_p]),
    "tpc_sample_categorical": (c_int, [c_void_p, c_void_p, c_int, c_int, c_u64, c_i64, c_i64, c_void_p, c_void_p]),
    "tpc_rollout": (c_int, [c_void_p, C.POINTER(EnvConfig), c_void_p, c_void_p, C.POINTER(RolloutBuffers), c_int,
                            c_int, c_int, c_int, c_u64, c_i64, c_i64, c_void_p, c_size_t, c_void_p]),
    "tpc_a2c_grads": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, C.POINTER(RolloutBuffers), c_int,
                              c_int, c_int, C.POINTER(A2CHyper), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_size_t, c_void_p]),
    "tpc_zero": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "tpc_scale": (c_int, [c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p]),
    "tpc_tf32_peak_probe": (c_int, [c_void_p, c_int, C.POINTER(c_float), c_void_p]),
    "tpc_adam_apply": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float,
                               c_float, c_void_p]),
    "tpc_heuristic_dominant": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                       c_void_p, c_void_p]),
    "tpc_heuristic_random": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_u64, c_i64, c_i64,
                                     c_void_p, c_void_p, c_void_p]),
    "tpc_solution_stats": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "tpc_compare_solutions": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                      c_void_p, c_void_p]),
    "tpc_knapsack_waste_reduction": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tpc_knapsack_random": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_u64, c_i64, c_i64, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "tpc_knapsack_bin_values": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                        c_void_p, c_void_p]),
    "tpc_knapsack_stats": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p]),
    "tpc_discounted_returns": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    "tpc_value_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "tpc_actor_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    "tpc_episode_reward_stats": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
}

_lib = None


class TpcError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libtpc.so and attach the prototypes. Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TpcError(f"{LIB_PATH} is missing: run `python -m tpc_b200.build` (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise TpcError(f"libtpc call {what} failed: {ERRORS.get(rc, rc)}")


def ptr(t) -> c_void_p:
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return c_void_p(None)
    return c_void_p(t.data_ptr())
