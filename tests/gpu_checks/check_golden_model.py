"""GPU check: libtpc (through the Python mirror of the reference Agent) against golden vectors recorded from the
reference's OWN model / agent / trainer code (tests/golden/model_*.npz, see tests/golden/make_model_golden.py).

For every case the recorded trajectory is replayed through the step-wise API (`Agent.act` with the recorded feasible
mask, `Agent.store` with the recorded action / reward), then the three loss calls of trainer.py:112-136 and the fused
device path (`tpc_a2c_grads` + both `tpc_adam_apply`) are compared with what the reference computed; the next
iteration starts from the parameters the CUDA path produced, so the last iteration bounds the drift of K updates.
Returns {metric: value}; tests/test_gpu_golden_model.py holds the tolerances."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import model as om  # noqa: E402
from tests import golden_model_util as gu  # noqa: E402
from tpc_b200.agent import Agent  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def run_case(case, verbose=True):
    z, meta = gu.load(case)
    ac = meta["agent_config"]
    a, c = gu.net_cfgs(meta)
    aspec, cspec = om.actor_spec(a), om.critic_spec(c)
    F = meta["env_config"]["num_features"]
    agent = Agent("transformer", ac, num_features=F)
    agent.stochastic_action_selection = False
    eng = agent.engine
    # the engine's layout is the reference's trainable_weights order (asserted name by name on the CPU side)
    for which, spec in ((0, aspec), (1, cspec)):
        offs, sizes = eng.layout(which)
        ooffs, _ = om.spec_offsets(spec)
        assert offs == [o for (o, _) in ooffs.values()] and sizes == [int(np.prod(s)) for _, s in spec]
    agent.bin_actor.set_flat(om.golden_params(aspec, int(z["actor_seed"])))
    agent.critic.set_flat(om.golden_params(cspec, int(z["critic_seed"])))
    out = {}
    K = int(z["n_iters"])
    for k in range(K):
        it = gu.iteration(z, k)
        T, B, S, Fs = it["states"].shape
        tag = f"it{k}"
        agent.clear_memory()
        e_act, agree, decided = 0.0, 0, 0
        for t in range(T):
            mask_t = it["bin_masks"][t]
            ids, mask_out, probs = agent.act(it["states"][t], it["dec_inputs"][t], mask_t, it["mha_masks"][t],
                                             lambda s, d, m, _m=mask_t: _m)
            e_act = max(e_act, float(np.abs(np.asarray(probs) - it["act_probs"][t]).max()))
            # greedy pick == the reference's argmax wherever its top-2 probability margin is decisive
            ref_lg = it["logits"].reshape(T, B, S)[t]
            top2 = np.sort(ref_lg, axis=-1)[:, -2:]
            clear = (top2[:, 1] - top2[:, 0]) > 1e-3 * np.abs(top2[:, 1]).clip(1.0)
            decided += int(clear.sum())
            agree += int((np.asarray(ids)[clear] == it["argmax"].reshape(T, B)[t][clear]).sum())
            agent.store(it["states"][t].copy(), it["dec_inputs"][t].copy(), mask_t.copy(), it["mha_masks"][t].copy(),
                        it["actions"][t].copy(), it["rewards"][:, t:t + 1].copy(), t)
        out[f"{tag}_act_probs_abs"] = e_act
        out[f"{tag}_greedy_disagree"] = decided - agree
        returns = agent.compute_discounted_rewards(np.zeros((B, 1), np.float32))
        out[f"{tag}_returns_abs"] = float(np.abs(returns - it["returns"]).max())
        vloss, values, adv = agent.compute_value_loss(returns)
        out[f"{tag}_values"] = rel(values, it["values"])
        out[f"{tag}_advantages"] = rel(adv, it["advantages"])
        out[f"{tag}_loss_value"] = rel(vloss, it["value_loss"])
        total, _dec, ent, pl, probs = agent.compute_actor_loss(agent.bin_actor, agent.bin_masks, agent.bins,
                                                               agent.actor_decoder_input, adv)
        out[f"{tag}_probs_abs"] = float(np.abs(np.asarray(probs) - it["probs"]).max())
        out[f"{tag}_loss_policy"] = rel(np.mean(pl), it["policy_loss"].mean())
        out[f"{tag}_loss_policy_rows"] = rel(pl, it["policy_loss"])
        out[f"{tag}_loss_entropy"] = rel(np.mean(ent), it["entropy"].mean())
        out[f"{tag}_loss_total"] = rel(total, it["actor_loss"])
        # logits through the model call itself (agents/agent.py:180-189)
        fi = gu.flat_inputs(it)
        lg, pr, idx, _ = agent.bin_actor(fi["states"], fi["dec"], fi["ptr_mask"], True, agent.num_bins,
                                         enc_padding_mask=fi["mha"], dec_padding_mask=fi["mha"])
        live = fi["ptr_mask"] == 0
        out[f"{tag}_logits"] = rel(np.asarray(lg)[live], it["logits"][live])
        assert np.all(np.asarray(pr)[~live] == 0), "masked probabilities must be exactly 0"

        # fused device path: tpc_a2c_grads, then both Adam steps (trainer.py:112-156)
        bufs = agent.memory_to_device()
        agent.engine.a2c_grads(bufs, agent.gamma, agent.values_loss_coefficient, agent.entropy_coefficient)
        torch.cuda.synchronize()
        fused = (eng.losses / float(T * B)).cpu().numpy()
        out[f"{tag}_loss_fused_value"] = rel(fused[0], it["value_loss"])
        out[f"{tag}_loss_fused_policy"] = rel(fused[1], it["policy_loss"].mean())
        out[f"{tag}_loss_fused_total"] = rel(fused[2], it["actor_loss"])
        out[f"{tag}_loss_fused_entropy"] = rel(fused[3], it["entropy"].mean())
        live_t = {}
        for which, name, spec in ((0, "actor", aspec), (1, "critic", cspec)):
            g = eng.grads[which].cpu().numpy()
            norms, samp = gu.sample_flat(g, spec)
            ref_n, ref_s = it[f"{name}_grad_norms"], it[f"{name}_grad_sample"]
            # tensors whose true gradient is zero (wk.bias: softmax shift invariance; decoder mha1.wq / wk: one key)
            # hold rounding noise in the reference too: they are excluded from the relative measures
            big = ref_n > 1e-3 * ref_n.max()
            live_t[name] = big
            out[f"{tag}_grad_norm_{name}"] = float(np.max(np.abs(norms[big] - ref_n[big]) / ref_n[big]))
            own = gu.sample_owner(spec)
            rms = ref_n[own] / np.sqrt([np.prod(s) for _, s in spec])[own]
            e = np.abs(samp - ref_s)[big[own]] / rms[big[own]]
            out[f"{tag}_grad_elem_{name}"] = float(e.max())
            if verbose and k == 0:
                worst = own[big[own]][int(np.argmax(e))]
                print(f"    {name}: worst element error {e.max():.2e} x rms in {spec[worst][0]} (norm {ref_n[worst]:.2e}, max norm {ref_n.max():.2e})")
        # the same update in chunks of ONE step: every chunk leaves out the rule tokens placed before its step
        # (TokWin, csrc/kernels.cuh) -- losses and actor gradients must not move
        full = [eng.grads[0].clone(), eng.losses.clone()]
        agent.engine.a2c_grads(bufs, agent.gamma, agent.values_loss_coefficient, agent.entropy_coefficient,
                               chunk_states=B)
        torch.cuda.synchronize()
        out[f"{tag}_window_grad_actor"] = float((eng.grads[0] - full[0]).norm() / full[0].norm())
        out[f"{tag}_window_losses"] = float(((eng.losses - full[1]).abs() / full[1].abs().clamp_min(1e-12)).max())
        agent.critic_opt.apply_gradients()
        agent.pointer_opt.apply_gradients()
        torch.cuda.synchronize()
        for which, name, spec in ((0, "actor", aspec), (1, "critic", cspec)):
            _, samp = gu.sample_flat(eng.params[which].cpu().numpy(), spec)
            own = gu.sample_owner(spec)
            err = (np.abs(samp - it[f"{name}_weights_after"]) / ac[name]["learning_rate"])[live_t[name][own]]
            out[f"{tag}_update_q95_{name}"] = float(np.quantile(err, 0.95))      # in Adam-step units
            out[f"{tag}_update_mean_{name}"] = float(err.mean())
            out[f"{tag}_update_max_{name}"] = float(err.max())
        if verbose:
            print(f"[{case}] " + " ".join(f"{kk[len(tag) + 1:]}={v:.2e}" for kk, v in out.items() if kk.startswith(tag)))
    agent.engine.close()
    return out


def main(cases=None):
    res = {}
    for case in (cases or gu.CASES):
        for k, v in run_case(case).items():
            res[f"{case}:{k}"] = v
    return res


if __name__ == "__main__":
    main(sys.argv[1:] or None)
