"""CPU oracle (torch restatement) of the reference actor / critic / losses / optimizer.

TEST INFRASTRUCTURE ONLY (see oracle/env.py header). PARITY UNPINNED for model numerics: the arithmetic lives
in un-vendored tensorflow==2.3.1 / tensorflow-probability==0.11.1 (requirements.txt:4-5), which cannot be
installed here, and the reference's own tests hold no numeric golden logits / values / losses
(tests/unit/agent/agent_test.py checks shapes and act-vs-loss self-consistency only). This file restates the
documented Keras semantics op for op; each function cites the reference lines it follows.

All functions take a flat parameter vector in the canonical order of `actor_spec` / `critic_spec`
(Keras attribute-tracking order: SURVEY appendix B) and run in the dtype of that vector (fp32 = the
reference's arithmetic, fp64 = arbiter).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

BIG = 1e9


# ------------------------------------------------------------------------------------------------------
# configuration + parameter inventory
# ------------------------------------------------------------------------------------------------------
@dataclass
class NetCfg:
    """Subset of configs/*.json `<agent>.agent_config.{actor,critic}` + env-injected keys (env.py:314-334)."""
    num_layers: int = 1
    d_model: int = 128
    num_heads: int = 8
    dff: int = 128
    common_embedding: bool = True
    num_bin_features: int = 3
    num_resource_features: int = 3
    dec_features: int = 3
    # actor only
    logit_clipping_C: float | None = 10.0
    attention_dense_units: int = 128
    # critic only
    last_layer_units: int = 128
    last_layer_activation: str = "linear"
    tensor_size: int = 0  # S = num_bins + num_resources (critic Flatten width = S * d_model)
    use_default_initializer: bool = True


def _mha_spec(prefix, d):
    out = []
    for n in ("wq", "wk", "wv", "dense"):
        out += [(f"{prefix}.{n}.kernel", (d, d)), (f"{prefix}.{n}.bias", (d,))]
    return out


def _ffn_spec(prefix, d, dff):
    return [(f"{prefix}.0.kernel", (d, dff)), (f"{prefix}.0.bias", (dff,)),
            (f"{prefix}.1.kernel", (dff, d)), (f"{prefix}.1.bias", (d,))]


def _ln_spec(prefix, d):
    return [(f"{prefix}.gamma", (d,)), (f"{prefix}.beta", (d,))]


def _enc_layer_spec(prefix, d, dff):
    # common/encoder_layer.py:9-13: mha, ffn, layernorm1, layernorm2
    return _mha_spec(f"{prefix}.mha", d) + _ffn_spec(f"{prefix}.ffn", d, dff) + \
        _ln_spec(f"{prefix}.layernorm1", d) + _ln_spec(f"{prefix}.layernorm2", d)


def _dec_layer_spec(prefix, d, dff):
    # actor/decoder_layer.py:9-16: mha1, mha2, ffn, layernorm1..3
    return _mha_spec(f"{prefix}.mha1", d) + _mha_spec(f"{prefix}.mha2", d) + _ffn_spec(f"{prefix}.ffn", d, dff) + \
        _ln_spec(f"{prefix}.layernorm1", d) + _ln_spec(f"{prefix}.layernorm2", d) + _ln_spec(f"{prefix}.layernorm3", d)


def _encoder_spec(prefix, c: NetCfg):
    # common/encoder.py:31-63: embedding1 [, embedding2], enc_layers, enc_layers_bins, enc_layers_resources
    d = c.d_model
    fb = c.num_bin_features
    out = [(f"{prefix}.embedding1.kernel", (fb, d)), (f"{prefix}.embedding1.bias", (d,))]
    if not c.common_embedding:
        out += [(f"{prefix}.embedding2.kernel", (c.num_resource_features, d)), (f"{prefix}.embedding2.bias", (d,))]
    for grp in ("enc_layers", "enc_layers_bins", "enc_layers_resources"):
        for i in range(c.num_layers):
            out += _enc_layer_spec(f"{prefix}.{grp}.{i}", d, c.dff)
    return out


def actor_spec(c: NetCfg):
    """actor/model.py:22-39 -> encoder, decoder{embedding, dec_layers, last_decoder_layer{W1,W2,V}}."""
    d, u = c.d_model, c.attention_dense_units
    out = _encoder_spec("encoder", c)
    out += [("decoder.embedding.kernel", (c.dec_features, d)), ("decoder.embedding.bias", (d,))]
    for i in range(c.num_layers):
        out += _dec_layer_spec(f"decoder.dec_layers.{i}", d, c.dff)
    out += [("decoder.pointer.W1.kernel", (d, u)), ("decoder.pointer.W1.bias", (u,)),
            ("decoder.pointer.W2.kernel", (d, u)), ("decoder.pointer.W2.bias", (u,)),
            ("decoder.pointer.V.kernel", (u, 1)), ("decoder.pointer.V.bias", (1,))]
    return out


def critic_spec(c: NetCfg):
    """critic/model.py:31-53 -> encoder, final_layer0, final_layer1, final_layer."""
    d, u = c.d_model, c.last_layer_units
    out = _encoder_spec("encoder", c)
    out += [("final_layer0.kernel", (c.tensor_size * d, u)), ("final_layer0.bias", (u,)),
            ("final_layer1.kernel", (u, u)), ("final_layer1.bias", (u,)),
            ("final_layer.kernel", (u, 1)), ("final_layer.bias", (1,))]
    return out


def spec_offsets(spec):
    offs, o = {}, 0
    for name, shape in spec:
        n = int(np.prod(shape))
        offs[name] = (o, shape)
        o += n
    return offs, o


def init_params(spec, seed: int, use_default_initializer=True, dims=128) -> np.ndarray:
    """Keras defaults: Dense kernel glorot_uniform (or U(+-1/sqrt(dims)), common/utils.py:49-64), bias 0,
    LayerNormalization gamma 1 / beta 0. A NumPy Generator replaces TF's RNG (weights are an input to parity
    runs, not something to reproduce)."""
    rng = np.random.default_rng(seed)
    offs, total = spec_offsets(spec)
    flat = np.zeros(total, dtype=np.float32)
    for name, shape in spec:
        o, _ = offs[name]
        n = int(np.prod(shape))
        if name.endswith(".kernel"):
            lim = math.sqrt(6.0 / (shape[0] + shape[1])) if use_default_initializer else 1.0 / math.sqrt(dims)
            flat[o:o + n] = rng.uniform(-lim, lim, size=n).astype(np.float32)
        elif name.endswith(".gamma"):
            flat[o:o + n] = 1.0
    return flat


def golden_params(spec, seed: int) -> np.ndarray:
    """Deterministic NON-trivial parameters for the golden model vectors (tests/golden/make_model_golden.py writes
    these into the reference's variables in ITS `trainable_weights` order; the tests rebuild them from `spec`, so
    any disagreement between the two orders shows up as a numeric mismatch): kernels glorot-uniform, biases and
    LayerNorm beta U(-0.1, 0.1), gamma U(0.9, 1.1), one draw sequence consumed tensor by tensor."""
    rng = np.random.default_rng(seed)
    chunks = []
    for name, shape in spec:
        n = int(np.prod(shape))
        if name.endswith(".kernel"):
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            chunks.append(rng.uniform(-lim, lim, size=n))
        elif name.endswith(".gamma"):
            chunks.append(rng.uniform(0.9, 1.1, size=n))
        else:
            chunks.append(rng.uniform(-0.1, 0.1, size=n))
    return np.concatenate(chunks).astype(np.float32)


def sample_indices(n: int, k: int = 64) -> np.ndarray:
    """Fixed positions at which the goldens keep gradient / weight entries of a tensor with n elements."""
    return np.unique(np.linspace(0, n - 1, min(n, k)).astype(np.int64))


class P:
    """View helper: P(flat, spec)['name'] -> tensor view with the right shape (keeps autograd graph)."""

    def __init__(self, flat: torch.Tensor, spec):
        self.flat = flat
        self.offs, self.total = spec_offsets(spec)
        assert flat.numel() == self.total, (flat.numel(), self.total)

    def __getitem__(self, name):
        o, shape = self.offs[name]
        return self.flat[o:o + int(np.prod(shape))].view(*shape)


# ------------------------------------------------------------------------------------------------------
# layers
# ------------------------------------------------------------------------------------------------------
def dense(p: P, name, x):
    return x @ p[f"{name}.kernel"] + p[f"{name}.bias"]


def layer_norm(p: P, name, x, eps=1e-6):
    """tf.keras.layers.LayerNormalization(epsilon=1e-6) (encoder_layer.py:12-13): biased variance, last axis."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * p[f"{name}.gamma"] + p[f"{name}.beta"]


def sdpa(q, k, v, mask):
    """common/utils.py:13-47."""
    logits = q @ k.transpose(-1, -2) / math.sqrt(k.shape[-1])
    if mask is not None:
        logits = logits + mask * -BIG
    w = torch.softmax(logits, dim=-1)
    return w @ v, w


def mha(p: P, name, v, k, q, mask, num_heads):
    """common/attention.py:30-52."""
    B = q.shape[0]
    d = q.shape[-1]
    depth = d // num_heads

    def split(x):
        return x.view(B, -1, num_heads, depth).permute(0, 2, 1, 3)

    qh = split(dense(p, f"{name}.wq", q))
    kh = split(dense(p, f"{name}.wk", k))
    vh = split(dense(p, f"{name}.wv", v))
    att, w = sdpa(qh, kh, vh, mask)
    att = att.permute(0, 2, 1, 3).reshape(B, -1, d)
    return dense(p, f"{name}.dense", att), w


def ffn(p: P, name, x):
    """common/utils.py:5-11."""
    return dense(p, f"{name}.1", torch.relu(dense(p, f"{name}.0", x)))


def encoder_layer(p: P, name, x, mask, num_heads):
    """common/encoder_layer.py:15-23."""
    attn, _ = mha(p, f"{name}.mha", x, x, x, mask, num_heads)
    out1 = layer_norm(p, f"{name}.layernorm1", x + attn)
    return layer_norm(p, f"{name}.layernorm2", out1 + ffn(p, f"{name}.ffn", out1))


def encoder(p: P, c: NetCfg, x, num_bins, mask, prefix="encoder"):
    """common/encoder.py:67-121. mask: (B,1,1,S), 1 = hidden key."""
    if c.common_embedding:
        e = dense(p, f"{prefix}.embedding1", x)
        eb, er = e[:, :num_bins], e[:, num_bins:]
    else:
        eb = dense(p, f"{prefix}.embedding1", x[:, :num_bins, : c.num_bin_features])
        er = dense(p, f"{prefix}.embedding2", x[:, num_bins:, : c.num_resource_features])
    for i in range(c.num_layers):
        eb = encoder_layer(p, f"{prefix}.enc_layers_bins.{i}", eb, mask[:, :, :, :num_bins], c.num_heads)
    for i in range(c.num_layers):
        er = encoder_layer(p, f"{prefix}.enc_layers_resources.{i}", er, mask[:, :, :, num_bins:], c.num_heads)
    cat = torch.cat([eb, er], dim=1)
    for i in range(c.num_layers):
        cat = encoder_layer(p, f"{prefix}.enc_layers.{i}", cat, mask, c.num_heads)
    return cat


def decoder_layer(p: P, name, x, enc_out, look_ahead_mask, padding_mask, num_heads):
    """actor/decoder_layer.py:18-32."""
    a1, _ = mha(p, f"{name}.mha1", x, x, x, look_ahead_mask, num_heads)
    o1 = layer_norm(p, f"{name}.layernorm1", a1 + x)
    a2, _ = mha(p, f"{name}.mha2", enc_out, enc_out, o1, padding_mask, num_heads)
    o2 = layer_norm(p, f"{name}.layernorm2", a2 + o1)
    return layer_norm(p, f"{name}.layernorm3", ffn(p, f"{name}.ffn", o2) + o2)


def pointer_attention(p: P, c: NetCfg, dec_out, enc_out, mask):
    """actor/pointer_attention.py:21-71. dec_out (B,1,d), enc_out (B,S,d), mask (B,S)."""
    score = dense(p, "decoder.pointer.V",
                  torch.tanh(dense(p, "decoder.pointer.W1", dec_out) + dense(p, "decoder.pointer.W2", enc_out)))
    logits = score.squeeze(2)
    if c.logit_clipping_C is not None:
        logits = c.logit_clipping_C * torch.tanh(logits)
    logits = logits - mask * BIG
    probs = torch.softmax(logits, dim=-1)
    return logits, probs, torch.argmax(probs, dim=-1)


def actor_forward(flat, c: NetCfg, state, dec_input, ptr_mask, mha_mask, num_bins):
    """actor/model.py:41-72 + actor/decoder.py:62-101. Returns (logits, probs, argmax index)."""
    p = P(flat, actor_spec(c))
    enc_out = encoder(p, c, state, num_bins, mha_mask)
    y = dense(p, "decoder.embedding", dec_input)
    for i in range(c.num_layers):
        y = decoder_layer(p, f"decoder.dec_layers.{i}", y, enc_out, None, mha_mask, c.num_heads)
    return pointer_attention(p, c, y, enc_out, ptr_mask)


def critic_forward(flat, c: NetCfg, state, mha_mask, num_bins):
    """critic/model.py:55-85. Returns (B,1) state values."""
    p = P(flat, critic_spec(c))
    enc_out = encoder(p, c, state, num_bins, mha_mask)
    x = enc_out.reshape(enc_out.shape[0], -1)
    act = {"linear": lambda t: t, "relu": torch.relu, "tanh": torch.tanh}[c.last_layer_activation or "linear"]
    x = act(dense(p, "final_layer0", x))
    x = act(dense(p, "final_layer1", x))
    return dense(p, "final_layer", x)


# ------------------------------------------------------------------------------------------------------
# agent math: agents/agent.py
# ------------------------------------------------------------------------------------------------------
def compute_discounted_rewards(rewards: np.ndarray, gamma: float, bootstrap: np.ndarray | None = None):
    """agents/agent.py:110-124. rewards (B,T) -> returns (B,T); bootstrap (B,) defaults to 0 (trainer.py:103-105)."""
    B, T = rewards.shape
    boot = np.zeros(B, dtype=np.float32) if bootstrap is None else np.asarray(bootstrap, dtype=np.float32).reshape(B)
    out = np.zeros_like(rewards)
    for step in reversed(range(T)):
        est = rewards[:, step] + np.float32(gamma) * boot
        out[:, step] = est
        boot = est
    return out


def value_loss(values, returns_vertical, coef):
    """agents/agent.py:155-166: c_v * mean((R - V)^2); advantages = R - V. values/returns: (T*B,1)."""
    adv = returns_vertical - values
    loss = coef * ((returns_vertical - values) ** 2).mean(dim=-1)
    return loss.mean(dim=-1), adv


def actor_loss(logits, actions, advantages, entropy_coef):
    """agents/agent.py:189-240. logits (T*B,S) (masked entries hold -1e9), actions (T*B,), advantages (T*B,1)
    already stop-gradient. Returns (total, entropy(T*B,), policy_loss(T*B,), probs)."""
    logp = torch.log_softmax(logits, dim=-1)
    probs = torch.softmax(logits, dim=-1).detach()
    policy_loss = -logp.gather(1, actions.view(-1, 1).long()).squeeze(1)
    pl_adv = policy_loss * advantages.squeeze(-1)
    entropy = -(probs * logp).sum(dim=-1)   # CategoricalCrossentropy(from_logits)(stop_grad(p), logits)
    total = (pl_adv - entropy_coef * entropy).mean(dim=-1)
    return total, entropy, policy_loss, probs


def keras_adam_step(params, grads, m, v, step, lr, spec, clipnorm=None, beta1=0.9, beta2=0.999, eps=1e-7):
    """tf.keras.optimizers.Adam (agents/agent.py:39-49): per-variable clip_by_norm(clipnorm) then
    m,v update and theta -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v)+eps). In-place on NumPy arrays; step is 1-based."""
    offs, _ = spec_offsets(spec)
    lr_t = lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    for name, (o, shape) in offs.items():
        n = int(np.prod(shape))
        g = grads[o:o + n].astype(np.float32).copy()
        if clipnorm is not None:
            norm = float(np.sqrt(np.sum(g.astype(np.float64) ** 2)))
            if norm > clipnorm:
                g = g * np.float32(clipnorm / norm)
        m[o:o + n] = beta1 * m[o:o + n] + (1 - beta1) * g
        v[o:o + n] = beta2 * v[o:o + n] + (1 - beta2) * g * g
        params[o:o + n] -= np.float32(lr_t) * m[o:o + n] / (np.sqrt(v[o:o + n]) + np.float32(eps))
    return params, m, v
