"""Quick perf probe: rollout / update / Adam split and host launch time (default: the headline shape B=1024, 50 nodes x
100 rules; B= N= R= select another one, e.g. B=128 N=11 R=20 for the shipped config)."""
import ctypes as C, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import model as om
from tpc_b200 import lib as L
from tpc_b200.engine import Engine, RolloutTensors, env_config_struct
from tests.gpu_checks.check_model import make_cfg

dev = torch.device("cuda:0")
B = int(os.environ.get("B", 1024)); N, R, F = int(os.environ.get("N", 51)), int(os.environ.get("R", 100)), 3
chunk = int(os.environ.get("CHUNK", 10240))
cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "env_rv3_shipped_greedy.json")))
mc, ca, cc = make_cfg(N, R)
eng = Engine(mc, dev)
eng.set_params(0, om.init_params(om.actor_spec(ca), 1))
eng.set_params(1, om.init_params(om.critic_spec(cc), 2))
e = env_config_struct(cfg, 0)
prof = torch.empty(cfg["num_profiles"], F, device=dev)
L.check(eng.lib.tpc_env_generate_profiles(eng.h, C.byref(e), 1235, L.ptr(prof), eng.stream))
bufs = RolloutTensors(B, N, R, F, dev)
def ev(): return torch.cuda.Event(enable_timing=True)
def reset(ep):
    L.check(eng.lib.tpc_env_reset(eng.h, C.byref(e), 1235, 0, ep, L.ptr(prof), B, N, R, L.ptr(bufs.batch), L.ptr(bufs.bin_net_mask), L.ptr(bufs.mha_mask), None, eng.stream))
for it in range(int(os.environ.get("ITERS", 3))):
    e0, e1, e2, e3 = ev(), ev(), ev(), ev()
    reset(it)
    torch.cuda.synchronize(); h0 = time.perf_counter()
    e0.record(); eng.rollout(e, bufs, False, 5, 0, it); e1.record()
    h1 = time.perf_counter()
    eng.a2c_grads(bufs, 0.99, 1.0, 0.01, chunk_states=chunk); e2.record()
    h2 = time.perf_counter()
    print(f"   host launch time: rollout {1e3*(h1-h0):.1f} ms, a2c {1e3*(h2-h1):.1f} ms")
    eng.adam_apply(0, 1e-4, 1.0); eng.adam_apply(1, 5e-4, 1.0); e3.record()
    torch.cuda.synchronize()
    tr, tu, ta = e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3)
    tot = tr + tu + ta
    print(f"iter {it}: rollout {tr:.1f} ms, a2c grads {tu:.1f} ms, adam {ta:.2f} ms, total {tot:.1f} ms -> {B*R/tot*1e3:.0f} decisions/s; "
          f"losses {eng.losses.cpu().numpy()/ (B*R)}; mean reward {bufs.rewards.sum(1).mean().item():.2f}")
print("max mem GB", torch.cuda.max_memory_allocated()/1e9)
