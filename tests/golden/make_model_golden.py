"""Generate golden MODEL vectors by executing the UNMODIFIED reference model / agent / trainer code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_model_golden.py

`agents/models/transformer/**`, `agents/agent.py`, `agents/trainer.py` and the ResourceV3 / KnapsackV2 environments
are imported as they are; their TensorFlow / Keras / TFP calls resolve to `oracle/tf_shim` (torch-CPU fp32 for the
model arithmetic, NumPy for the environment), because TensorFlow cannot be installed here. The reference's own
`trainer()` drives everything: instance-level recorders (no reference code is edited) capture what crosses
`Agent.act`, `Agent.compute_value_loss`, `Agent.compute_actor_loss` and `Adam.apply_gradients`.

Per case `tests/golden/model_<case>.npz` holds, for each training iteration k:
  inputs   : states, dec_inputs, bin_masks (the feasible masks `act` returned), mha_masks, actions, rewards
  outputs  : returns, values, advantages, value_loss; logits, probs, argmax index (one actor call on the T*B
             concatenation, as compute_actor_loss does), act-time probs; total / policy / entropy losses
  gradients: per-tensor L2 norm + entries at `oracle.model.sample_indices` for both networks
  weights  : the same sample of both parameter sets after the two Adam steps
plus the parameter inventory (attribute path + shape of every `trainable_weights` entry, in order) and the 7
buffers `trainer()` returns. Parameters are `oracle.model.golden_params(seed)` written in the reference's order.
"""
import contextlib
import io
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
sys.path.insert(0, "/root/reference")

import types  # noqa: E402

# /root/reference/agents has no __init__.py (a namespace package); a regular `agents` package in site-packages
# would shadow it, so bind the name to the reference directory explicitly.
_agents = types.ModuleType("agents")
_agents.__path__ = ["/root/reference/agents"]
sys.modules["agents"] = _agents

import tensorflow as tf  # noqa: E402  (the shim)
from tensorflow.keras.layers import Layer  # noqa: E402
from agents.agent import Agent  # noqa: E402
from agents.trainer import trainer  # noqa: E402
from environment.custom.resource_v3.env import ResourceEnvironmentV3  # noqa: E402
from environment.custom.knapsack_v2.env import KnapsackEnvironmentV2  # noqa: E402
from oracle import model as om  # noqa: E402


def weight_paths(model):
    """Attribute path of every variable, walking `__dict__` in assignment order (the order Keras tracks)."""
    found = {}

    def walk(layer, path):
        for key, val in vars(layer).items():
            if key.startswith("_"):
                continue
            if isinstance(val, tf.Variable):
                found.setdefault(id(val), f"{path}.{key}" if path else key)
            elif isinstance(val, Layer):
                walk(val, f"{path}.{key}" if path else key)
            elif isinstance(val, (list, tuple)):
                for i, v in enumerate(val):
                    if isinstance(v, Layer):
                        walk(v, f"{path}.{key}.{i}" if path else f"{key}.{i}")
    walk(model, "")
    return [found[id(w)] for w in model.trainable_weights]


def set_params(model, seed):
    ws = model.trainable_weights
    paths = weight_paths(model)
    spec = [(p, tuple(w.shape)) for p, w in zip(paths, ws)]
    flat = om.golden_params(spec, seed)
    o = 0
    for w in ws:
        n = int(np.prod(w.shape))
        w.assign(flat[o:o + n].reshape(w.shape))
        o += n
    return spec, flat


def flat_of(tensors):
    return np.concatenate([np.asarray(t, dtype=np.float32).reshape(-1) for t in tensors])


def sampled(tensors):
    """-> (per-tensor L2 norms, concatenated samples at om.sample_indices)."""
    norms, vals = [], []
    for t in tensors:
        a = np.asarray(t, dtype=np.float32).reshape(-1)
        norms.append(np.sqrt(np.sum(a.astype(np.float64) ** 2)))
        vals.append(a[om.sample_indices(a.size)])
    return np.asarray(norms, np.float64), np.concatenate(vals)


def run_case(name, env_kind, env_over, agent_over, iterations, seed):
    cfg_file = {"rv3": "ResourceV3.json", "kv2": "KnapsackV2.json"}[env_kind]
    allcfg = json.load(open(os.path.join("/root/reference/configs", cfg_file)))
    env_cfg = allcfg["env_config"]
    for k, v in env_over.items():
        if k == "reward_type":
            env_cfg["reward"]["type"] = v
        else:
            env_cfg[k] = v
    agent_cfg = allcfg["tpc"]["agent_config"]
    for k, v in agent_over.items():
        tgt = agent_cfg
        *head, leaf = k.split(".")
        for h in head:
            tgt = tgt[h]
        tgt[leaf] = v
    trainer_cfg = allcfg["trainer_config"]
    trainer_cfg["n_iterations"] = iterations
    trainer_cfg["store_model_weights"]["export_weights"] = False

    tf.random.set_seed(seed)
    np.random.seed(seed)
    import tensorflow_probability as tfp
    tfp.distributions.Categorical._rng = np.random.default_rng(seed + 1)

    env = (ResourceEnvironmentV3 if env_kind == "rv3" else KnapsackEnvironmentV2)(cfg_file[:-5], env_cfg)
    agent_cfg = env.add_stats_to_agent_config(agent_cfg)
    agent = Agent("transformer", agent_cfg)

    # variables are created at the first call: build both networks, then write the golden parameters
    st, dec, bmask, mmask = env.reset()
    agent.bin_actor(st, dec, env.build_feasible_mask(st, dec, bmask), True, agent.num_bins,
                    enc_padding_mask=mmask, dec_padding_mask=mmask)
    agent.critic(st, True, agent.num_bins, enc_padding_mask=mmask)
    actor_spec, actor_flat = set_params(agent.bin_actor, seed + 10)
    critic_spec, critic_flat = set_params(agent.critic, seed + 20)

    rec = {"iters": []}
    cur = {}

    def new_iter():
        cur.clear()
        cur.update(act_probs=[])

    new_iter()
    orig_act = agent.act

    def act(state, dec_input, bins_mask, mha_used_mask, build_feasible_mask):
        out = orig_act(state, dec_input, bins_mask, mha_used_mask, build_feasible_mask)
        cur["act_probs"].append(np.asarray(out[2], dtype=np.float32))
        return out
    agent.act = act

    orig_value = agent.compute_value_loss

    def compute_value_loss(discounted_rewards):
        T = len(agent.states)
        cur.update(
            states=np.stack(agent.states).astype(np.float32),
            dec_inputs=np.stack(agent.actor_decoder_input).astype(np.float32),
            bin_masks=np.stack(agent.bin_masks).astype(np.float32),
            mha_masks=np.stack(agent.mha_masks).astype(np.float32),
            actions=np.stack([np.asarray(b) for b in agent.bins]).astype(np.int32),
            rewards=agent.rewards.copy(),
            returns=np.asarray(discounted_rewards, dtype=np.float32).copy())
        assert cur["states"].shape[0] == T
        out = orig_value(discounted_rewards)
        cur.update(value_loss=np.float32(out[0].numpy()), values=out[1].numpy(), advantages=out[2].numpy())
        return out
    agent.compute_value_loss = compute_value_loss

    orig_actor = agent.compute_actor_loss

    def compute_actor_loss(model, masks, actions, decoder_inputs, advantages):
        out = orig_actor(model, masks, actions, decoder_inputs, advantages)
        cur.update(actor_loss=np.float32(out[0].numpy()), entropy=out[2].numpy(), policy_loss=out[3].numpy(),
                   probs=np.asarray(out[4], dtype=np.float32))
        # the logits themselves: one more call of the reference model on the same concatenation (agent.py:176-189)
        lg, pr, idx, _ = model(tf.concat(agent.states, axis=0), tf.concat(decoder_inputs, axis=0),
                               tf.concat(masks, axis=0), agent.training, agent.num_bins,
                               enc_padding_mask=tf.concat(agent.mha_masks, axis=0),
                               dec_padding_mask=tf.concat(agent.mha_masks, axis=0))
        cur.update(logits=lg.numpy(), argmax=np.asarray(idx, dtype=np.int32))
        return out
    agent.compute_actor_loss = compute_actor_loss

    def wrap_opt(opt, key, model):
        orig = opt.apply_gradients

        def apply_gradients(grads_and_vars):
            gv = list(grads_and_vars)
            assert [id(v) for _, v in gv] == [id(v) for v in model.trainable_weights]
            gn, gs = sampled([np.zeros(v.shape, np.float32) if g is None else g for g, v in gv])
            cur[f"{key}_grad_norms"], cur[f"{key}_grad_sample"] = gn, gs
            orig(gv)
            _, cur[f"{key}_weights_after"] = sampled([v.numpy() for v in model.trainable_weights])
            if key == "actor":         # second optimizer step of the iteration (trainer.py:144-156)
                cur["act_probs"] = np.stack(cur["act_probs"])
                rec["iters"].append(dict(cur))
                new_iter()
        opt.apply_gradients = apply_gradients
    wrap_opt(agent.critic_opt, "critic", agent.critic)
    wrap_opt(agent.pointer_opt, "actor", agent.bin_actor)

    with contextlib.redirect_stdout(io.StringIO()) as out:
        buffers = trainer(env, agent, trainer_cfg, True, "/tmp")
    stdout_lines = out.getvalue().strip().splitlines()

    save = {
        "actor_paths": np.array([p for p, _ in actor_spec]), "critic_paths": np.array([p for p, _ in critic_spec]),
        "actor_shapes": np.array([json.dumps(s) for _, s in actor_spec]),
        "critic_shapes": np.array([json.dumps(s) for _, s in critic_spec]),
        "actor_seed": seed + 10, "critic_seed": seed + 20,
        "actor_param_checksum": np.float64(actor_flat.astype(np.float64).sum()),
        "critic_param_checksum": np.float64(critic_flat.astype(np.float64).sum()),
        "trainer_buffers": np.asarray(buffers, dtype=np.float32),
        "stdout": np.array(stdout_lines),
        "n_iters": len(rec["iters"]),
    }
    for k, it in enumerate(rec["iters"]):
        for key, val in it.items():
            save[f"it{k}_{key}"] = val
    path = os.path.join(HERE, f"model_{name}.npz")
    np.savez_compressed(path, **save)
    meta = {"env_kind": env_kind, "env_config": env_cfg, "agent_config": agent_cfg, "iterations": iterations}
    with open(os.path.join(HERE, f"model_{name}.json"), "w") as f:
        json.dump(meta, f, indent=1, default=lambda o: None)
    it0 = rec["iters"][0]
    print(f"{name}: states {it0['states'].shape} value_loss {it0['value_loss']:.6f} actor_loss {it0['actor_loss']:.6f} "
          f"-> {os.path.getsize(path) / 1e3:.0f} KB")


CASES = [
    # name, env, env overrides, agent overrides, iterations
    ("rv3_6x10_greedy", "rv3", dict(batch_size=4, node_sample_size=5, profiles_sample_size=10, node_max_val=60), {}, 5),
    ("rv3_11x20_greedy", "rv3", dict(batch_size=3), {}, 2),
    ("rv3_51x100_greedy", "rv3", dict(batch_size=2, node_sample_size=50, profiles_sample_size=100), {}, 1),
    ("rv3_6x10_cost", "rv3", dict(batch_size=4, node_sample_size=5, profiles_sample_size=10, node_max_val=60,
                                  reward_type="reduced_node_usage"), {}, 2),
    ("rv3_6x10_fair_noclip", "rv3", dict(batch_size=4, node_sample_size=5, profiles_sample_size=10, node_max_val=60,
                                         reward_type="global_dominant", mask_nodes_in_mha=False),
     {"actor.clipnorm": None, "critic.clipnorm": None}, 2),
    ("kv2_5x12", "kv2", dict(batch_size=4, bin_sample_size=4, item_sample_size=12), {}, 2),
]


def main():
    only = sys.argv[1:]
    for name, env_kind, env_over, agent_over, iters in CASES:
        if only and name not in only:
            continue
        run_case(name, env_kind, env_over, agent_over, iters, seed=zlib.crc32(name.encode()) % 100000)


if __name__ == "__main__":
    main()
