"""Correctness (vs torch fp64) and live timing of the attention kernels at the update-chunk shapes."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tpc_b200 import lib as L
lib = L.load()
h = C.c_void_p(); L.check(lib.tpc_create(C.byref(h)))
dev = torch.device("cuda:0")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(0)

def ref(qkv, mask, Cn, Lt):
    q, k, v = (qkv[:, i * 128:(i + 1) * 128].double().view(Cn, Lt, 8, 16).permute(0, 2, 1, 3) for i in range(3))
    s = q @ k.transpose(-1, -2) / 4 + mask.double().view(Cn, 1, 1, Lt) * -1e9
    p = torch.softmax(s, -1)
    return (p @ v).permute(0, 2, 1, 3).reshape(Cn * Lt, 128)

def run(Cn, Lt, frac_masked, check):
    qkv = torch.randn(Cn * Lt, 384, device=dev)
    mask = torch.zeros(Cn, Lt, device=dev)
    nm = int(Lt * frac_masked)
    if nm:
        mask[:, Lt - nm:] = 1            # contiguous block of masked keys (placed rules)
    att = torch.empty(Cn * Lt, 128, device=dev); lse = torch.empty(Cn * Lt, 8, device=dev)
    datt = torch.randn(Cn * Lt, 128, device=dev); dqkv = torch.empty(Cn * Lt, 384, device=dev)
    f = lambda: L.check(lib.tpc_attention_fwd(h, L.ptr(qkv), L.ptr(mask), Lt, 0, Cn, Lt, L.ptr(att), L.ptr(lse), st))
    b = lambda: L.check(lib.tpc_attention_bwd(h, L.ptr(qkv), L.ptr(att), L.ptr(lse), L.ptr(datt), L.ptr(mask), Lt, 0, Cn, Lt, L.ptr(dqkv), st))
    f(); b(); torch.cuda.synchronize()
    if check:
        x = qkv.double().clone().requires_grad_(True)
        o = ref(x, mask, Cn, Lt)
        o.backward(datt.double())
        e1 = (att.double() - o).abs().max().item() / o.abs().max().item()
        e2 = (dqkv.double() - x.grad).abs().max().item() / x.grad.abs().max().item()
        print(f"  check C={Cn} L={Lt} masked={frac_masked}: fwd {e1:.2e} bwd {e2:.2e}")
        assert e1 < 1e-3 and e2 < 5e-3, (e1, e2)
        return
    for name, fn in (("fwd", f), ("bwd", b)):
        for _ in range(2): fn()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 5 * 1e3
        print(f"attention {name} C={Cn} L={Lt} masked={frac_masked}: {us:8.1f} us  ({us / Cn * 148:.2f} us per state-SM)")

if not os.environ.get('SKIP_CHECK'):
    for (Cn, Lt, fm) in [(5, 151, 0.3), (7, 100, 0.5), (9, 51, 0.0), (3, 16, 0.25)]:
        run(Cn, Lt, fm, True)
for (Lt, fm) in [(151, 0.0), (151, 0.33), (100, 0.5), (51, 0.1)]:
    run(int(os.environ.get('STATES', 2048)), Lt, fm, False)
