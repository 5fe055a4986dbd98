"""Oracle model: parameter inventory, self-consistency properties the reference's agent_test.py checks, and the
hand-verifiable pieces (returns, zero-gradient entropy term, Keras Adam)."""
import numpy as np
import torch

from oracle import model as om


def test_parameter_counts_match_survey():
    a = om.NetCfg(num_layers=1, dff=128)
    assert om.spec_offsets(om.actor_spec(a))[1] == 498_817
    assert om.spec_offsets(om.critic_spec(om.NetCfg(num_layers=3, dff=512, tensor_size=151)))[1] == 4_275_713
    assert om.spec_offsets(om.critic_spec(om.NetCfg(num_layers=3, dff=512, tensor_size=31)))[1] == 2_309_633
    cost = om.NetCfg(num_layers=1, dff=128, common_embedding=False, num_bin_features=4, num_resource_features=3)
    assert om.spec_offsets(om.actor_spec(cost))[1] == 499_457
    kv2 = om.NetCfg(num_layers=1, dff=128, common_embedding=False, num_bin_features=2, num_resource_features=2,
                    dec_features=2)
    assert om.spec_offsets(om.actor_spec(kv2))[1] == 498_945


def _problem(B=2, N=4, R=2, seed=0):
    rng = np.random.default_rng(seed)
    S = N + R
    st = rng.random((B, S, 3)).astype(np.float32)
    st[:, 0] = -2
    pm = np.zeros((B, S), np.float32)
    pm[:, N:] = 1
    mm = np.zeros((B, 1, 1, S), np.float32)
    return torch.tensor(st), torch.tensor(st[:, N:N + 1]), torch.tensor(pm), torch.tensor(mm), N


def test_act_shapes_and_probs_consistency():
    """agent_test.py:103-137 (shapes, bin_ids within the node range) and :217-220 (probs at act() time equal the
    probs recomputed in the loss to 5 decimals)."""
    c = om.NetCfg(num_layers=1, dff=128)
    flat = torch.tensor(om.init_params(om.actor_spec(c), 3))
    st, dec, pm, mm, N = _problem()
    logits, probs, idx = om.actor_forward(flat, c, st, dec, pm, mm, N)
    assert logits.shape == (2, 6) and probs.shape == (2, 6) and idx.shape == (2,)
    assert int(idx.max()) <= 3
    assert torch.all(probs[:, N:] == 0)                      # masked positions: exactly 0 (fp32 fact)
    _, _, _, p2 = om.actor_loss(logits, idx, torch.zeros(2, 1), 0.01)
    np.testing.assert_almost_equal(probs.numpy(), p2.numpy(), decimal=5)


def test_entropy_term_has_zero_gradient():
    """labels = stop_gradient(softmax(l)) => dCE/dl = softmax(l) - p = 0."""
    logits = torch.randn(5, 7, requires_grad=True)
    adv = torch.zeros(5, 1)
    total, ent, pl, _ = om.actor_loss(logits, torch.zeros(5, dtype=torch.long), adv, 0.5)
    total.backward()
    assert float(logits.grad.abs().max()) < 1e-6


def test_discounted_rewards():
    """agents/agent.py:110-124 with bootstrap 0."""
    r = np.array([[1, 0, 2], [0, 1, 1]], dtype=np.float32)
    got = om.compute_discounted_rewards(r, 0.5)
    assert got.tolist() == [[1 + 0.5 * (0 + 0.5 * 2), 0 + 0.5 * 2, 2], [0 + 0.5 * (1 + 0.5), 1.5, 1]]


def test_keras_adam_first_step_and_clipnorm():
    """First Adam step moves every weight by ~lr * sign(g) (m/sqrt(v) = 1 after bias correction); a tensor whose
    gradient norm exceeds clipnorm is rescaled before the moments are updated."""
    spec = [("a.kernel", (2, 2)), ("a.bias", (2,))]
    p = np.zeros(6, np.float32)
    g = np.array([3, -4, 0, 0, 0.1, -0.1], np.float32)     # kernel norm 5 -> clipped to 1; bias norm 0.14 -> kept
    m, v = np.zeros(6, np.float32), np.zeros(6, np.float32)
    om.keras_adam_step(p, g, m, v, 1, 0.01, spec, clipnorm=1.0)
    np.testing.assert_allclose(m[:2], 0.1 * np.array([0.6, -0.8]), rtol=1e-6)
    np.testing.assert_allclose(p[:2], [-0.01, 0.01], rtol=1e-4)
    np.testing.assert_allclose(p[4:], [-0.01, 0.01], rtol=1e-4)
    assert p[2] == 0 and p[3] == 0


def test_critic_is_tied_to_training_shape():
    c = om.NetCfg(num_layers=1, dff=128, tensor_size=6)
    flat = torch.tensor(om.init_params(om.critic_spec(c), 1))
    st, dec, pm, mm, N = _problem()
    assert om.critic_forward(flat, c, st, mm, N).shape == (2, 1)
