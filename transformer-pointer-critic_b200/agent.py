"""Agent with the reference's interface (agents/agent.py:15-305) on top of the libtpc engine.

Two ways to drive it:
  * step-wise, exactly like the reference (`act` / `store` / `compute_*`): every model evaluation is a libtpc
    forward call; arrays cross the API as host NumPy like in the reference (compatibility mode, used by tests);
  * fused (`rollout` + `update`): the whole T-step rollout and the A2C update run on the device behind two
    C-ABI calls (tpc_rollout, tpc_a2c_grads + tpc_adam_apply); this is what `trainer` / `tester` use.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import lib as L
from .engine import Engine, RolloutTensors, model_config_from_agent_config
from .env import HostTensor

TRANSFORMER = "transformer"


def _dev(x, device, dtype=torch.float32):
    if torch.is_tensor(x):
        return x.to(device=device, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(device)


def reshape_into_vertical_format(data, batch_size):
    """environment/custom/resource_v3/misc/utils.py:13-17: (B,T) -> (T*B,1), step-major."""
    return np.reshape(np.transpose(np.asarray(data), [1, 0]), [batch_size, 1])


def reshape_into_horizontal_format(data, batch_size, decoding_steps):
    """environment/custom/resource_v3/misc/utils.py:7-11."""
    return np.transpose(np.reshape(np.asarray(data), [decoding_steps, batch_size]), [1, 0])


class _Net:
    """Stands in for the Keras model objects `agent.bin_actor` / `agent.critic`."""

    def __init__(self, agent: "Agent", which: int):
        self.agent, self.which = agent, which

    @property
    def engine(self) -> Engine:
        return self.agent.engine

    @property
    def trainable_weights(self):
        offs, sizes = self.engine.layout(self.which)
        flat = self.engine.params[self.which]
        return [flat[o:o + n] for o, n in zip(offs, sizes)]

    def get_flat(self) -> np.ndarray:
        return self.engine.params[self.which].detach().cpu().numpy().copy()

    def set_flat(self, flat) -> None:
        self.engine.set_params(self.which, flat)

    def _named_tensors(self) -> dict:
        """{Keras attribute path: ndarray} in `trainable_weights` order (tf_checkpoint.keras_path)."""
        from . import tf_checkpoint as tfc
        names = self.engine.tensor_names(self.which)
        offs, sizes = self.engine.layout(self.which)
        shapes = self.engine.tensor_shapes(self.which)
        flat = self.get_flat()
        return {tfc.keras_path(n): flat[o:o + k].reshape(s) for n, o, k, s in zip(names, offs, sizes, shapes)}

    def save_weights(self, location: str) -> None:
        """keras.Model.save_weights(location) (agents/agent.py:289-290). A path without `.npz` gets what Keras writes
        for it -- a TF checkpoint `<location>.index` + `<location>.data-00000-of-00001` (tf_checkpoint.py; the format
        could not be verified against TensorFlow itself) -- and a flat `<location>.npz` beside it; `*.npz` gets the
        flat file only."""
        os.makedirs(os.path.dirname(os.path.abspath(location)) or ".", exist_ok=True)
        offs, sizes = self.engine.layout(self.which)
        np.savez(location if location.endswith(".npz") else location + ".npz", flat=self.get_flat(),
                 offsets=np.asarray(offs), sizes=np.asarray(sizes), names=np.asarray(self.engine.tensor_names(self.which)))
        if not location.endswith(".npz"):
            from . import tf_checkpoint as tfc
            tfc.write_checkpoint(location, self._named_tensors())
            base = os.path.basename(location)
            with open(os.path.join(os.path.dirname(os.path.abspath(location)), "checkpoint"), "w") as f:
                f.write(f'model_checkpoint_path: "{base}"\nall_model_checkpoint_paths: "{base}"\n')

    def load_weights(self, location: str) -> None:
        """keras.Model.load_weights(location) (agents/agent.py:293-294): a TF checkpoint written by the reference (or
        by save_weights above) when `<location>.index` exists, else the flat `.npz`."""
        if not location.endswith(".npz") and os.path.exists(location + ".index"):
            from . import tf_checkpoint as tfc
            stored = tfc.read_checkpoint(location)
            names = self.engine.tensor_names(self.which)
            offs, sizes = self.engine.layout(self.which)
            shapes = self.engine.tensor_shapes(self.which)
            flat = np.zeros(self.engine.n_params[self.which], dtype=np.float32)
            for n, o, k, s in zip(names, offs, sizes, shapes):
                key = tfc.keras_path(n)
                if key not in stored:
                    raise ValueError(f"{location}: variable {key} is missing from the checkpoint")
                if tuple(stored[key].shape) != tuple(s):
                    raise ValueError(f"{location}: {key} has shape {stored[key].shape}, this network expects {tuple(s)}")
                flat[o:o + k] = stored[key].reshape(-1)
            self.set_flat(flat)
            return
        path = location if location.endswith(".npz") else location + ".npz"
        with np.load(path) as z:
            offs, sizes = self.engine.layout(self.which)
            if "offsets" in z.files and (list(z["offsets"]) != offs or list(z["sizes"]) != sizes):
                raise ValueError(f"{path}: stored tensor layout does not match this network "
                                 f"({len(z['sizes'])} tensors / {int(np.sum(z['sizes']))} parameters stored, "
                                 f"{len(sizes)} / {sum(sizes)} expected): different architecture or a critic file")
            self.set_flat(z["flat"])

    def __call__(self, *args, **kwargs):
        if self.which == 0:
            # ActorTransformer.call(enc_input, dec_input, attention_mask, training, num_bins, enc_padding_mask=...)
            state, dec_input, attention_mask, _training, num_bins = args[:5]
            mha = kwargs.get("enc_padding_mask")
            dev = self.engine.device
            st = _dev(state, dev)
            B, S = st.shape[0], st.shape[1]
            logits, probs, idx = self.engine.actor_forward(st, _dev(dec_input, dev).reshape(B, -1),
                                                           _dev(attention_mask, dev), _dev(mha, dev).reshape(B, S),
                                                           num_bins)
            i = idx.cpu().numpy()
            decoded = np.asarray(state)[np.arange(B), i] if not torch.is_tensor(state) else state[torch.arange(B), idx.long()]
            return (logits.cpu().numpy().view(HostTensor), probs.cpu().numpy().view(HostTensor), i.view(HostTensor),
                    decoded)
        # CriticTransformer.call(encoder_input, training, num_bins, enc_padding_mask)
        state, _training, num_bins = args[:3]
        mha = kwargs.get("enc_padding_mask", args[3] if len(args) > 3 else None)
        dev = self.engine.device
        st = _dev(state, dev)
        v = self.engine.critic_forward(st, _dev(mha, dev).reshape(st.shape[0], st.shape[1]), num_bins)
        return v.cpu().numpy().reshape(-1, 1).view(HostTensor)


class _Adam:
    """tf.keras.optimizers.Adam(learning_rate, clipnorm) (agents/agent.py:39-49) -> tpc_adam_apply."""

    def __init__(self, agent: "Agent", which: int, learning_rate: float, clipnorm):
        self.agent, self.which = agent, which
        self.learning_rate, self.clipnorm = learning_rate, clipnorm

    def apply_gradients(self, grads_and_vars=None, grad_scale: float = 1.0):
        """Applies the gradients currently held in the engine's gradient buffer of this network."""
        self.agent.engine.adam_apply(self.which, self.learning_rate, self.clipnorm, grad_scale)


class Agent:
    def __init__(self, name, opts, device="cuda:0", num_features=None, seed: int = 0):
        self.name = name
        self.agent_config = opts
        self.training = True
        self.batch_size: int = opts["batch_size"]
        self.num_resources: int = opts["num_resources"]
        self.num_bins: int = opts["num_bins"]
        self.tensor_size: int = opts["tensor_size"]
        self.gamma: float = opts["gamma"]
        self.values_loss_coefficient: float = opts["values_loss_coefficient"]
        self.entropy_coefficient: float = opts["entropy_coefficient"]
        self.stochastic_action_selection: bool = opts["stochastic_action_selection"]
        self.critic_learning_rate = opts["critic"]["learning_rate"]
        self.critic_clipnorm = opts["critic"]["clipnorm"]
        self.actor_learning_rate = opts["actor"]["learning_rate"]
        self.actor_clipnorm = opts["actor"]["clipnorm"]
        if name != TRANSFORMER:
            raise ValueError(f"unknown agent type {name!r} (model_factory.py:35 returns None for it)")
        emb = opts["encoder_embedding"]
        if num_features is None:
            num_features = emb.get("num_resource_features") or 3
        self.num_features = num_features
        self.engine = Engine(model_config_from_agent_config(opts, num_features), device)
        self.bin_actor, self.critic = _Net(self, 0), _Net(self, 1)
        self.critic_opt = _Adam(self, 1, self.critic_learning_rate, self.critic_clipnorm)
        self.pointer_opt = _Adam(self, 0, self.actor_learning_rate, self.actor_clipnorm)
        self.init_weights(seed)
        self.sample_seed = 0x5EED + seed
        self.episode_id0 = 0
        self._draw = 0
        self.clear_memory()

    # ------------------------------------------------------------------ weights
    def init_weights(self, seed: int = 0) -> None:
        """Keras defaults: zero biases, LayerNorm gamma 1 / beta 0, kernels `glorot_uniform`
        (use_default_initializer) or U(+-1/sqrt(dims)) with dims = dim_model for the encoder / decoder / MHA / FFN /
        critic head and dims = attention_dense_units for the pointer's W1, W2, V (common/utils.py:49-64,
        actor/pointer_attention.py:13)."""
        rng = np.random.default_rng(seed)
        for which, sub in ((0, "actor"), (1, "critic")):
            offs, sizes = self.engine.layout(which)
            flat = np.zeros(self.engine.n_params[which], dtype=np.float32)
            default = self.agent_config[sub].get("use_default_initializer", True)
            n_pairs = len(offs) // 2
            # the canonical list alternates (kernel, bias) for Dense and (gamma, beta) for LayerNorm; a LayerNorm
            # pair is the only one with equal sizes (no Dense here has a single input feature)
            for i in range(0, len(offs), 2):
                n, nb = sizes[i], sizes[i + 1]
                if n == nb:
                    flat[offs[i]:offs[i] + n] = 1.0
                else:
                    fan_out, fan_in = nb, n // nb
                    pointer = which == 0 and i // 2 >= n_pairs - 3        # last_decoder_layer {W1, W2, V}
                    dims = self.agent_config[sub]["attention_dense_units" if pointer else "dim_model"]
                    lim = np.sqrt(6.0 / (fan_in + fan_out)) if default else 1.0 / np.sqrt(dims)
                    flat[offs[i]:offs[i] + n] = rng.uniform(-lim, lim, size=n).astype(np.float32)
            self.engine.set_params(which, flat)

    def save_weights(self, location):
        self.bin_actor.save_weights(location)

    def load_weights(self, location):
        self.bin_actor.load_weights(location)

    def set_testing_mode(self, batch_size, num_bins, num_resources):
        """agents/agent.py:296-305."""
        self.training = False
        self.stochastic_action_selection = False
        self.batch_size = batch_size
        self.num_resources = num_resources
        self.num_bins = num_bins

    # ------------------------------------------------------------------ memory (agents/agent.py:75-108)
    def store(self, state, dec_input, bin_mask, mha_mask, bin, reward, training_step):
        self.states.append(state)
        self.actor_decoder_input.append(dec_input)
        self.bin_masks.append(bin_mask)
        self.mha_masks.append(mha_mask)
        self.bins.append(bin)
        self.rewards[:, training_step] = np.asarray(reward)[:, 0]

    def clear_memory(self):
        self.states = []
        self.actor_decoder_input = []
        self.bin_masks = []
        self.mha_masks = []
        self.bins = []
        self.rewards = np.zeros((self.batch_size, self.num_resources), dtype="float32")

    # ------------------------------------------------------------------ step-wise API
    def act(self, state, dec_input, bins_mask, mha_used_mask, build_feasible_mask):
        """agents/agent.py:244-287."""
        bins_mask = build_feasible_mask(state, dec_input, bins_mask)
        _logits, bins_probs, bin_ids, _ = self.bin_actor(state, dec_input, bins_mask, self.training, self.num_bins,
                                                         enc_padding_mask=mha_used_mask,
                                                         dec_padding_mask=mha_used_mask)
        if self.stochastic_action_selection:
            dev = self.engine.device
            p = _dev(np.asarray(bins_probs), dev)
            out = torch.empty(p.shape[0], dtype=torch.int32, device=dev)
            L.check(self.engine.lib.tpc_sample_categorical(self.engine.h, L.ptr(p), p.shape[0], p.shape[1],
                                                           self.sample_seed, self.episode_id0, self._draw, L.ptr(out),
                                                           self.engine.stream), "tpc_sample_categorical")
            self._draw += 1      # one Philox draw index per decision
            bin_ids = out.cpu().numpy().view(HostTensor)
        return bin_ids, bins_mask, bins_probs

    def compute_discounted_rewards(self, bootstrap_state_value):
        """agents/agent.py:110-124 -> (B,T) returns (device kernel tpc_discounted_returns)."""
        eng = self.engine
        B, T = self.rewards.shape
        rw = _dev(self.rewards, eng.device)
        boot = _dev(np.asarray(bootstrap_state_value, dtype=np.float32).reshape(-1), eng.device)
        if boot.numel() != B:
            raise ValueError(f"bootstrap_state_value must hold one value per episode ({B}), got {boot.numel()}")
        out = torch.empty(B, T, dtype=torch.float32, device=eng.device)
        L.check(eng.lib.tpc_discounted_returns(eng.h, L.ptr(rw), L.ptr(boot), B, T, float(self.gamma), L.ptr(out),
                                               eng.stream), "tpc_discounted_returns")
        return out.cpu().numpy()

    def compute_value_loss(self, discounted_rewards):
        """agents/agent.py:126-166 -> (value_loss, state_values, advantages); arithmetic in tpc_value_loss."""
        eng = self.engine
        states = np.concatenate(self.states, axis=0)
        n = states.shape[0]
        mha = np.concatenate(self.mha_masks, axis=0)
        values = np.asarray(self.critic(states, self.training, self.num_bins, enc_padding_mask=mha))
        R = reshape_into_vertical_format(discounted_rewards, n).astype(np.float32)
        Rd, Vd = _dev(R.reshape(-1), eng.device), _dev(values.reshape(-1), eng.device)
        adv = torch.empty(n, dtype=torch.float32, device=eng.device)
        loss = torch.empty(1, dtype=torch.float32, device=eng.device)
        L.check(eng.lib.tpc_value_loss(eng.h, L.ptr(Rd), L.ptr(Vd), n, float(self.values_loss_coefficient), L.ptr(adv),
                                       L.ptr(loss), eng.stream), "tpc_value_loss")
        return (np.float32(loss.cpu().numpy()[0]), values.view(HostTensor),
                adv.cpu().numpy().reshape(n, 1).view(HostTensor))

    def compute_actor_loss(self, model, masks, actions, decoder_inputs, advantages):
        """agents/agent.py:168-242 -> (total_loss, dec_output, entropy, policy_loss, probs); arithmetic in tpc_actor_loss."""
        eng = self.engine
        states = np.concatenate(self.states, axis=0)
        attention_mask = np.concatenate(masks, axis=0)
        mha = np.concatenate(self.mha_masks, axis=0)
        dec_input = np.concatenate(decoder_inputs, axis=0)
        logits, probs, _idx, dec_out = model(states, dec_input, attention_mask, self.training, self.num_bins,
                                             enc_padding_mask=mha, dec_padding_mask=mha)
        logits = np.ascontiguousarray(np.asarray(logits, dtype=np.float32))
        probs = np.asarray(probs, dtype=np.float32)
        n, S = logits.shape
        acts = _dev(np.concatenate([np.asarray(a) for a in actions], axis=0).reshape(-1), eng.device, torch.int32)
        adv = _dev(np.asarray(advantages, dtype=np.float32).reshape(-1), eng.device)
        lg = _dev(logits, eng.device)
        policy_loss = torch.empty(n, dtype=torch.float32, device=eng.device)
        entropy = torch.empty(n, dtype=torch.float32, device=eng.device)
        sums = torch.empty(3, dtype=torch.float32, device=eng.device)
        L.check(eng.lib.tpc_actor_loss(eng.h, L.ptr(lg), L.ptr(acts), L.ptr(adv), n, S, float(self.entropy_coefficient),
                                       L.ptr(policy_loss), L.ptr(entropy), L.ptr(sums), eng.stream), "tpc_actor_loss")
        total = np.float32(sums.cpu().numpy()[1])
        return (total, dec_out, entropy.cpu().numpy().view(HostTensor), policy_loss.cpu().numpy().view(HostTensor),
                probs.view(HostTensor))

    # ------------------------------------------------------------------ fused device path
    def memory_to_device(self) -> RolloutTensors:
        """Upload the step-wise memory (lists filled by `store`) into device trajectory buffers."""
        T, B = len(self.states), self.states[0].shape[0]
        N, R, F = self.num_bins, self.num_resources, self.num_features
        S = N + R
        states = np.stack(self.states)
        cost = states.shape[-1] > F
        bufs = RolloutTensors(B, N, R, F, self.engine.device, cost=cost)
        bufs.states.copy_(_dev(states[..., :F], self.engine.device))
        if cost:
            bufs.is_empties.copy_(_dev(states[..., F], self.engine.device))
        bufs.dec_inputs.copy_(_dev(np.stack(self.actor_decoder_input).reshape(T, B, F), self.engine.device))
        bufs.ptr_masks.copy_(_dev(np.stack(self.bin_masks), self.engine.device))
        bufs.mha_masks.copy_(_dev(np.stack(self.mha_masks).reshape(T, B, S), self.engine.device))
        bufs.actions.copy_(_dev(np.stack([np.asarray(b) for b in self.bins]), self.engine.device, torch.int32))
        bufs.rewards.copy_(_dev(self.rewards, self.engine.device))
        return bufs

    def update(self, bufs: RolloutTensors, world_size: int = 1, allreduce=None, chunk_states: int = 0):
        """trainer.py:103-156 on device: returns, value loss, actor loss, both backward passes, both Adam steps.
        `allreduce(tensor)` starts a SUM over the data-parallel ranks and returns a handle with `.wait()` (or None
        when it already completed); None for one process. The critic gradients are complete after the critic tape,
        so their all-reduce runs on the communication stream while the actor tape computes; the actor
        gradients and the 4 loss sums follow, and both Adam steps wait for the handles.
        Returns the 4 reported scalars (value_loss, mean policy loss, total actor loss, mean entropy) on the device."""
        n_local = bufs.B * bufs.R
        total = n_local * world_size
        eng = self.engine
        hyper = (self.gamma, self.values_loss_coefficient, self.entropy_coefficient)
        if allreduce is not None and world_size > 1:
            eng.a2c_grads(bufs, *hyper, chunk_states=chunk_states, total_states=total, phases=L.A2C_CRITIC)
            pending = [allreduce(eng.grads[1])]
            eng.a2c_grads(bufs, *hyper, chunk_states=chunk_states, total_states=total, phases=L.A2C_ACTOR)
            pending += [allreduce(eng.grads[0]), allreduce(eng.losses)]
            for w in pending:
                if w is not None:
                    w.wait()
        else:
            eng.a2c_grads(bufs, *hyper, chunk_states=chunk_states, total_states=total)
        self.critic_opt.apply_gradients()
        self.pointer_opt.apply_gradients()
        out = torch.empty(4, dtype=torch.float32, device=eng.device)
        L.check(eng.lib.tpc_scale(eng.h, L.ptr(eng.losses), 4, 1.0 / float(total), L.ptr(out), eng.stream), "tpc_scale")
        return out

    def rollout(self, env, bufs: RolloutTensors, greedy=None):
        """T steps of act -> env.step on device (trainer.py:49-100 / tester.py:173-209)."""
        greedy = (not self.stochastic_action_selection) if greedy is None else greedy
        # the kernel uses draw index epoch * T + t (net.cu): advance the shared counter by whole rollouts of THIS
        # length so that neither a later rollout with another T nor step-wise act() calls can reuse a
        # (seed, episode, draw) triple
        T = bufs.T
        epoch = -(-self._draw // T)
        self.engine.rollout(env._cfg_struct(), bufs, greedy, self.sample_seed, env.episode_id0, epoch)
        self._draw = (epoch + 1) * T
