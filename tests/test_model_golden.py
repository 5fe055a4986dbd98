"""oracle/model.py against golden vectors produced by the reference's OWN model / agent / trainer code
(tests/golden/make_model_golden.py executes agents/models/transformer/**, agents/agent.py and agents/trainer.py
unmodified on a torch-backed TensorFlow shim). This is what pins the oracle's model numerics, its canonical parameter
order, the step-major loss reshapes and the Adam update."""
import numpy as np
import pytest
import torch

from oracle import model as om
from tests import golden_model_util as gu

TOL = 5e-6   # fp32 restatement vs fp32 reference (different LayerNorm / softmax operation order): a few ulp of O(10) logits


def _params(z, meta):
    a, c = gu.net_cfgs(meta)
    return (a, om.golden_params(om.actor_spec(a), int(z["actor_seed"])),
            c, om.golden_params(om.critic_spec(c), int(z["critic_seed"])))


@pytest.mark.parametrize("case", gu.CASES)
def test_parameter_order_is_the_references_trainable_weights_order(case):
    z, meta = gu.load(case)
    a, c = gu.net_cfgs(meta)
    for which, spec in (("actor", om.actor_spec(a)), ("critic", om.critic_spec(c))):
        ref_names = [gu.normalise_path(p) for p in z[f"{which}_paths"]]
        ref_shapes = [tuple(__import__("json").loads(s)) for s in z[f"{which}_shapes"]]
        assert ref_names == [n for n, _ in spec]
        assert ref_shapes == [tuple(s) for _, s in spec]
    a_, af, c_, cf = _params(z, meta)
    assert abs(af.astype(np.float64).sum() - float(z["actor_param_checksum"])) < 1e-9
    assert abs(cf.astype(np.float64).sum() - float(z["critic_param_checksum"])) < 1e-9


def _oracle_iteration(a, af, c, cf, it, meta):
    """One trainer iteration on the oracle from the recorded trajectory -> dict of everything the golden holds."""
    ac = meta["agent_config"]
    fi = gu.flat_inputs(it)
    N = ac["num_bins"]
    at = torch.tensor(af, requires_grad=True)
    ct = torch.tensor(cf, requires_grad=True)
    st, dec = torch.tensor(fi["states"]), torch.tensor(fi["dec"])
    pm, mm = torch.tensor(fi["ptr_mask"]), torch.tensor(fi["mha"])
    returns = om.compute_discounted_rewards(it["rewards"], ac["gamma"])
    rv = torch.tensor(np.ascontiguousarray(returns.T).reshape(-1, 1))           # misc/utils.py:13-17
    values = om.critic_forward(ct, c, st, mm, N)
    vloss, adv = om.value_loss(values, rv, ac["values_loss_coefficient"])
    cg, = torch.autograd.grad(vloss, ct)
    logits, probs, idx = om.actor_forward(at, a, st, dec, pm, mm, N)
    total, ent, pl, p2 = om.actor_loss(logits, torch.tensor(fi["actions"]), adv.detach(), ac["entropy_coefficient"])
    ag, = torch.autograd.grad(total, at)
    return dict(returns=returns, values=values.detach().numpy(), advantages=adv.detach().numpy(),
                value_loss=float(vloss.detach()), logits=logits.detach().numpy(), probs=probs.detach().numpy(),
                argmax=idx.numpy(), actor_loss=float(total.detach()), entropy=ent.detach().numpy(),
                policy_loss=pl.detach().numpy(), actor_grad=ag.numpy(), critic_grad=cg.numpy())


@pytest.mark.parametrize("case", gu.CASES)
def test_oracle_matches_reference_forward_losses_gradients_and_updates(case):
    z, meta = gu.load(case)
    a, af, c, cf = _params(z, meta)
    ac = meta["agent_config"]
    aspec, cspec = om.actor_spec(a), om.critic_spec(c)
    am, av, cm, cv = (np.zeros_like(x) for x in (af, af, cf, cf))
    for k in range(int(z["n_iters"])):
        it = gu.iteration(z, k)
        o = _oracle_iteration(a, af, c, cf, it, meta)
        live = gu.flat_inputs(it)["ptr_mask"] == 0
        # iteration 0 starts from bit-identical parameters; later ones from parameters that went through Adam
        # steps whose first-step direction is sign(g): a gradient entry at rounding-noise level may move a weight
        # by 2 lr the other way, so the comparison of later iterations is a drift bound, not an ulp bound.
        # Measured fp32-vs-fp32 drift of the critic values: 4e-6, 2e-4, 9e-4, 2e-3 after 1..4 updates (lr 5e-4, value
        # losses of O(100): the reference's own training is that sensitive), so the bound grows 3x per update.
        tol, ltol = (TOL, 1e-5) if k == 0 else (1e-4 * 3 ** k, 3e-4 * 3 ** k)
        assert np.array_equal(o["returns"], it["returns"])
        np.testing.assert_allclose(o["values"], it["values"], rtol=tol, atol=tol * 10)
        np.testing.assert_allclose(o["advantages"], it["advantages"], rtol=tol, atol=tol * 10)
        np.testing.assert_allclose(o["value_loss"], it["value_loss"], rtol=ltol)
        np.testing.assert_allclose(o["logits"][live], it["logits"][live], rtol=tol, atol=tol * 10)
        assert np.array_equal(o["logits"][~live], it["logits"][~live])          # exactly -1e9
        np.testing.assert_allclose(o["probs"], it["probs"], atol=tol)
        assert np.all(it["probs"][~live] == 0) and np.all(o["probs"][~live] == 0)
        # act-time probabilities (batch B, one step) equal the loss-time recomputation (agent_test.py:217-220)
        T, B, S = it["act_probs"].shape
        np.testing.assert_allclose(it["act_probs"].reshape(T * B, S), it["probs"], atol=1e-5)
        assert k > 0 or np.array_equal(o["argmax"], it["argmax"])
        np.testing.assert_allclose(o["policy_loss"], it["policy_loss"], rtol=ltol, atol=tol)
        np.testing.assert_allclose(o["entropy"], it["entropy"], rtol=ltol, atol=tol)
        np.testing.assert_allclose(o["actor_loss"], it["actor_loss"], rtol=ltol, atol=tol)
        for which, spec, g in (("actor", aspec, o["actor_grad"]), ("critic", cspec, o["critic_grad"])):
            norms, samp = gu.sample_flat(g, spec)
            ref_n, ref_s = it[f"{which}_grad_norms"], it[f"{which}_grad_sample"]
            np.testing.assert_allclose(norms, ref_n, rtol=2e-4 if k == 0 else 2e-2, atol=1e-6 * ref_n.max())   # wk.bias: zero by shift invariance
            own = gu.sample_owner(spec)
            rms = ref_n[own] / np.sqrt([np.prod(s) for _, s in spec])[own]
            assert np.max(np.abs(samp - ref_s) / (rms + 1e-6 * ref_n.max())) < (2e-2 if k == 0 else 2e-1), which
        # both Adam steps (trainer.py:144-156), then compare the parameters the next iteration starts from
        om.keras_adam_step(cf, o["critic_grad"], cm, cv, k + 1, ac["critic"]["learning_rate"], cspec,
                           clipnorm=ac["critic"]["clipnorm"])
        om.keras_adam_step(af, o["actor_grad"], am, av, k + 1, ac["actor"]["learning_rate"], aspec,
                           clipnorm=ac["actor"]["clipnorm"])
        for which, spec, flat, lr in (("actor", aspec, af, ac["actor"]["learning_rate"]),
                                      ("critic", cspec, cf, ac["critic"]["learning_rate"])):
            _, samp = gu.sample_flat(flat, spec)
            err = np.abs(samp - it[f"{which}_weights_after"]) / lr        # in units of one Adam step
            assert np.quantile(err, 0.99) < 5e-2 and err.max() < 2.0 * (k + 1), (which, k, err.max())


@pytest.mark.parametrize("case", gu.CASES)
def test_trainer_buffers_are_the_recorded_scalars(case):
    """The 7 lists trainer() returns (trainer.py:186-192) are the per-iteration statistics of the recorded tensors."""
    z, _ = gu.load(case)
    buf = z["trainer_buffers"]
    for k in range(int(z["n_iters"])):
        it = gu.iteration(z, k)
        tot = it["rewards"].sum(axis=-1)
        assert buf[0][k] == np.float32(np.average(tot)) and buf[1][k] == tot.min() and buf[2][k] == tot.max()
        np.testing.assert_allclose(buf[3][k], it["value_loss"], rtol=1e-6)
        np.testing.assert_allclose(buf[4][k], it["policy_loss"].mean(), rtol=1e-5)
        np.testing.assert_allclose(buf[5][k], it["actor_loss"], rtol=1e-6)
        np.testing.assert_allclose(buf[6][k], it["entropy"].mean(), rtol=1e-5)
