"""Helpers shared by the CPU and GPU tests that check against tests/golden/model_*.npz (vectors recorded by running
the UNMODIFIED reference model / agent / trainer code, see tests/golden/make_model_golden.py)."""
import glob
import json
import os
import re

import numpy as np

from oracle import model as om

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[len("model_"):-len(".npz")] for p in glob.glob(os.path.join(GOLDEN, "model_*.npz")))


def load(case):
    z = np.load(os.path.join(GOLDEN, f"model_{case}.npz"))
    meta = json.load(open(os.path.join(GOLDEN, f"model_{case}.json")))
    return z, meta


def normalise_path(p: str) -> str:
    """Reference attribute path -> the oracle's spec name: TimeDistributed(`.layer`) and Sequential (`.layers`)
    wrappers carry no parameters of their own; `last_decoder_layer` is the pointer."""
    p = re.sub(r"\.layers\.(\d)\.", r".\1.", p)
    p = p.replace(".layer.", ".")
    return p.replace("decoder.last_decoder_layer.", "decoder.pointer.")


def net_cfgs(meta):
    """(actor NetCfg, critic NetCfg) from the recorded agent_config (model_factory.py:4-36)."""
    ac = meta["agent_config"]
    emb = ac["encoder_embedding"]
    F = meta["env_config"]["num_features"]
    common = dict(d_model=ac["actor"]["dim_model"], num_heads=ac["actor"]["num_heads"],
                  common_embedding=bool(emb["common"]), num_bin_features=emb["num_bin_features"] or F,
                  num_resource_features=emb["num_resource_features"] or F, dec_features=F,
                  tensor_size=ac["tensor_size"])
    a = om.NetCfg(num_layers=ac["actor"]["num_layers"], dff=ac["actor"]["inner_layer_dim"],
                  logit_clipping_C=ac["actor"]["logit_clipping_C"],
                  attention_dense_units=ac["actor"]["attention_dense_units"], **common)
    c = om.NetCfg(num_layers=ac["critic"]["num_layers"], dff=ac["critic"]["inner_layer_dim"],
                  last_layer_units=ac["critic"]["last_layer_units"],
                  last_layer_activation=ac["critic"]["last_layer_activation"], **common)
    return a, c


def iteration(z, k):
    pre = f"it{k}_"
    return {key[len(pre):]: z[key] for key in z.files if key.startswith(pre)}


def flat_inputs(it):
    """Step-major concatenation the agent performs (agent.py:131-137,176-179): row t*B+b."""
    T, B, S, F = it["states"].shape
    return dict(states=it["states"].reshape(T * B, S, F), dec=it["dec_inputs"].reshape(T * B, 1, -1),
                ptr_mask=it["bin_masks"].reshape(T * B, S), mha=it["mha_masks"].reshape(T * B, 1, 1, S),
                actions=it["actions"].reshape(T * B), T=T, B=B, S=S, F=F)


def sample_flat(flat, spec):
    """Entries of a flat parameter-shaped vector at the golden's sample positions + per-tensor norms."""
    offs, _ = om.spec_offsets(spec)
    vals, norms = [], []
    for name, shape in spec:
        o, _s = offs[name]
        n = int(np.prod(shape))
        a = np.asarray(flat[o:o + n], dtype=np.float32)
        vals.append(a[om.sample_indices(n)])
        norms.append(np.sqrt(np.sum(a.astype(np.float64) ** 2)))
    return np.asarray(norms), np.concatenate(vals)


def sample_owner(spec):
    """Tensor index of every sampled entry (to report which tensor a mismatch belongs to)."""
    return np.concatenate([np.full(len(om.sample_indices(int(np.prod(s)))), i) for i, (_, s) in enumerate(spec)])
