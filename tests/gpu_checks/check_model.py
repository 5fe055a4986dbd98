"""GPU check: actor / critic forward and A2C gradients of libtpc vs the torch CPU oracle (fp32 and fp64)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import model as om  # noqa: E402
from tpc_b200 import lib as L  # noqa: E402
from tpc_b200.engine import Engine, RolloutTensors  # noqa: E402


def make_cfg(N, R, nl_a=1, dff_a=128, nl_c=3, dff_c=512, common=True, Fb=3, Fr=3, F=3):
    mc = L.ModelConfig()
    mc.actor_num_layers, mc.actor_inner_layer_dim = nl_a, dff_a
    mc.critic_num_layers, mc.critic_inner_layer_dim = nl_c, dff_c
    mc.dim_model, mc.num_heads, mc.attention_dense_units, mc.last_layer_units = 128, 8, 128, 128
    mc.logit_clipping_C, mc.has_logit_clipping = 10.0, 1
    mc.common_embedding = 1 if common else 0
    mc.num_bin_features, mc.num_resource_features, mc.dec_features = Fb, Fr, F
    mc.num_bins, mc.num_resources = N, R
    a = om.NetCfg(num_layers=nl_a, dff=dff_a, common_embedding=common, num_bin_features=Fb if not common else F,
                  num_resource_features=Fr, dec_features=F)
    c = om.NetCfg(num_layers=nl_c, dff=dff_c, common_embedding=common, num_bin_features=Fb if not common else F,
                  num_resource_features=Fr, dec_features=F, tensor_size=N + R)
    return mc, a, c


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def random_problem(rng, n, N, R, F, Fstate):
    S = N + R
    state = rng.integers(0, 100, size=(n, S, Fstate)).astype(np.float32) / 100
    state[:, 0, :] = -2.0
    state[:, N:, :F] = rng.integers(1, 30, size=(n, R, F)).astype(np.float32) / 100
    if Fstate > F:
        state[:, :, F:] = rng.integers(0, 2, size=(n, S, Fstate - F)).astype(np.float32)
    t = rng.integers(0, R, size=n)                      # current rule index
    mha = np.zeros((n, S), dtype=np.float32)
    ptr = np.zeros((n, S), dtype=np.float32)
    ptr[:, N:] = 1
    for i in range(n):
        mha[i, N:N + t[i]] = 1                           # placed rules are hidden
        full = rng.random(N) < 0.2
        full[0] = False
        mha[i, :N][full] = 1
        ptr[i, :N][full] = 1
        ptr[i, :N][rng.random(N) < 0.2] = 1
        ptr[i, 0] = 0
    dec = state[np.arange(n), N + t, :F].copy()
    return state, dec, ptr, mha


def main():
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    dev = torch.device("cuda:0")
    worst = {}
    for (N, R, n, common, Fb) in [(6, 9, 37, True, 3), (11, 20, 64, True, 3), (51, 100, 24, True, 3), (8, 12, 33, False, 4)]:
        S, F = N + R, 3
        Fstate = F if common else Fb
        mc, ca, cc = make_cfg(N, R, common=common, Fb=Fb)
        eng = Engine(mc, dev)
        pa = om.init_params(om.actor_spec(ca), 1)
        pc = om.init_params(om.critic_spec(cc), 2)
        # make biases / LN params non-trivial so every term is exercised
        pa += (rng.standard_normal(pa.shape) * 0.02).astype(np.float32)
        pc += (rng.standard_normal(pc.shape) * 0.02).astype(np.float32)
        assert eng.n_params[0] == pa.size and eng.n_params[1] == pc.size, (eng.n_params, pa.size, pc.size)
        offs, sizes = eng.layout(0)
        ooffs, _ = om.spec_offsets(om.actor_spec(ca))
        assert offs == [o for (o, _) in ooffs.values()], "actor layout mismatch"
        offs, sizes = eng.layout(1)
        ooffs, _ = om.spec_offsets(om.critic_spec(cc))
        assert offs == [o for (o, _) in ooffs.values()], "critic layout mismatch"
        eng.set_params(0, pa)
        eng.set_params(1, pc)

        state, dec, ptr, mha = random_problem(rng, n, N, R, F, Fstate)
        tstate, tdec, tptr, tmha = (torch.tensor(x, device=dev) for x in (state, dec, ptr, mha))
        logits, probs, idx = eng.actor_forward(tstate, tdec, tptr, tmha, N)
        values = eng.critic_forward(tstate, tmha, N)
        torch.cuda.synchronize()
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            fa, fc = torch.tensor(pa, dtype=dt), torch.tensor(pc, dtype=dt)
            st = torch.tensor(state, dtype=dt)
            ol, op, oi = om.actor_forward(fa, ca, st, torch.tensor(dec, dtype=dt).unsqueeze(1), torch.tensor(ptr, dtype=dt),
                                          torch.tensor(mha, dtype=dt).view(n, 1, 1, S), N)
            ov = om.critic_forward(fc, cc, st, torch.tensor(mha, dtype=dt).view(n, 1, 1, S), N)
            unm = ptr == 0
            e_l = rel(logits.cpu().numpy()[unm], ol.detach().numpy()[unm])
            e_p = float(np.abs(probs.cpu().numpy() - op.detach().numpy()).max())
            e_v = rel(values.cpu().numpy(), ov.detach().numpy().ravel())
            agree = float((idx.cpu().numpy() == oi.numpy()).mean())
            print(f"[N={N} R={R} common={common}] vs {tag}: logits {e_l:.2e} probs(abs) {e_p:.2e} values {e_v:.2e} argmax agree {agree:.3f}")
            worst[f"fwd_{N}_{R}_{tag}"] = max(e_l, e_v)
        assert np.all(probs.cpu().numpy()[ptr == 1] == 0), "masked probs must be exactly 0"

        # ---------------- gradients on a synthetic trajectory ----------------
        B, T = 5, R
        nT = B * T
        state, dec, ptr, mha = random_problem(rng, nT, N, R, F, Fstate)
        actions = np.array([rng.choice(np.flatnonzero(ptr[i] == 0)) for i in range(nT)], dtype=np.int32)
        rewards = rng.standard_normal((B, T)).astype(np.float32)
        cost = not common
        bufs = RolloutTensors(B, N, R, F, dev, cost=cost)
        bufs.states.copy_(torch.tensor(state[:, :, :F]).view(T, B, S, F))
        if cost:
            bufs.is_empties.copy_(torch.tensor(state[:, :, F]).view(T, B, S))
        bufs.dec_inputs.copy_(torch.tensor(dec).view(T, B, F))
        bufs.ptr_masks.copy_(torch.tensor(ptr).view(T, B, S))
        bufs.mha_masks.copy_(torch.tensor(mha).view(T, B, S))
        bufs.actions.copy_(torch.tensor(actions).view(T, B))
        bufs.rewards.copy_(torch.tensor(rewards))
        gamma, c_v, c_e = 0.99, 1.0, 0.01
        for chunk in (0, 17):
            vals, adv = eng.a2c_grads(bufs, gamma, c_v, c_e, chunk_states=chunk)
            torch.cuda.synchronize()
            ga, gc = eng.grads[0].cpu().numpy().copy(), eng.grads[1].cpu().numpy().copy()
            losses = eng.losses.cpu().numpy().copy()
            dt = torch.float64
            fa = torch.tensor(pa, dtype=dt, requires_grad=True)
            fc = torch.tensor(pc, dtype=dt, requires_grad=True)
            st = torch.tensor(state, dtype=dt)
            m4 = torch.tensor(mha, dtype=dt).view(nT, 1, 1, S)
            Rv = torch.tensor(om.compute_discounted_rewards(rewards, gamma).T.reshape(nT, 1).astype(np.float64))
            ov = om.critic_forward(fc, cc, st, m4, N)
            vl, oadv = om.value_loss(ov, Rv, c_v)
            vl.backward()
            ol, op, oi = om.actor_forward(fa, ca, st, torch.tensor(dec, dtype=dt).unsqueeze(1), torch.tensor(ptr, dtype=dt), m4, N)
            tot, ent, pl, _ = om.actor_loss(ol, torch.tensor(actions), oadv.detach(), c_e)
            tot.backward()
            e_gc = rel(gc, fc.grad.numpy())
            e_ga = rel(ga, fa.grad.numpy())
            # per-tensor relative errors (clipnorm / Adam act per tensor)
            def per_tensor(g, ref, spec):
                offs, _ = om.spec_offsets(spec)
                out = {}
                for name, (o, shape) in offs.items():
                    k = int(np.prod(shape))
                    nr = np.linalg.norm(ref[o:o + k])
                    out[name] = float(np.linalg.norm(g[o:o + k] - ref[o:o + k]) / max(nr, 1e-12)) if nr > 1e-9 else float(np.linalg.norm(g[o:o + k]))
                return out
            pta = per_tensor(ga, fa.grad.numpy(), om.actor_spec(ca))
            ptc = per_tensor(gc, fc.grad.numpy(), om.critic_spec(cc))
            wa = max(pta, key=pta.get)
            wc = max(ptc, key=ptc.get)
            print(f"  grads chunk={chunk}: critic max-rel {e_gc:.2e} (worst tensor {wc} {ptc[wc]:.2e})  actor max-rel {e_ga:.2e} (worst tensor {wa} {pta[wa]:.2e})")
            print(f"  losses: value {losses[0]/nT:.6f} vs {float(vl):.6f} | policy {losses[1]/nT:.6f} vs {float(pl.mean()):.6f} | total {losses[2]/nT:.6f} vs {float(tot):.6f} | entropy {losses[3]/nT:.6f} vs {float(ent.mean()):.6f}")
            for nm, got, ref in (("value", losses[0] / nT, float(vl)), ("policy", losses[1] / nT, float(pl.mean())),
                                 ("total", losses[2] / nT, float(tot)), ("entropy", losses[3] / nT, float(ent.mean()))):
                # the total is a signed mean of policy_loss * advantage (random advantages here): its error is
                # measured against the magnitude of the terms, not against a sum that may cancel
                scale = float((pl * oadv.detach().squeeze(-1)).abs().mean()) if nm == "total" else abs(ref)
                worst[f"loss_{nm}_{N}_{R}_{chunk}"] = abs(got - ref) / max(abs(ref), scale, 1e-12)
            e_adv = rel(adv.cpu().numpy(), oadv.detach().numpy().ravel())
            print(f"  values/adv err {e_adv:.2e}")
            worst[f"grad_{N}_{R}_{chunk}"] = max(ptc[wc], pta[wa])
        eng.close()
    print("WORST", {k: f"{v:.2e}" for k, v in worst.items()})
    return worst


if __name__ == "__main__":
    t0 = time.time()
    main()
    print(f"done in {time.time() - t0:.1f}s")
