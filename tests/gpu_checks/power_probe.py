"""Board power and SM clock while ONE kernel runs back to back for a few seconds (nvidia-smi sampled every 100 ms):
shows which launches are power-limited (the forward GEMMs are: the SM clock drops with random operands and not
with zeros). KIND = ffn2_3x | ffn2_1x | qkv_3x | ffn1_3x | copy;  ZERO=1 uses all-zero operands."""
import ctypes as C, os, subprocess, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tpc_b200 import lib as L
lib = L.load()
h = C.c_void_p(); L.check(lib.tpc_create(C.byref(h)))
dev = torch.device("cuda:0")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
M = int(os.environ.get("CHUNK", 10240)) * 151
kind = os.environ.get("KIND", "ffn2_3x")
zero = os.environ.get("ZERO", "0") == "1"
mk = (lambda *s: torch.zeros(*s, device=dev)) if zero else (lambda *s: torch.randn(*s, device=dev))
def image(Wt):
    img = torch.empty_like(Wt)
    L.check(lib.tpc_gemm_split_image(h, L.ptr(Wt), Wt.shape[1], Wt.shape[0], Wt.shape[1], L.ptr(img), st))
    return img
g = torch.ones(128, device=dev); b = torch.zeros(128, device=dev); rstd = torch.empty(M, device=dev)
res = mk(M, 128); D = torch.empty(M, 128, device=dev)
if kind.startswith("ffn2"):
    A = mk(M, 512); W = mk(128, 512) / 22; Wi = image(W) if kind.endswith("3x") else None
    fn = lambda: lib.tpc_gemm(h, L.ptr(A), 512, L.ptr(W), 512, L.ptr(D), 128, M, 128, 512, L.ptr(b), L.ptr(res), 128, 1,
                              L.ptr(g), L.ptr(b), L.ptr(rstd), 2, 1, L.ptr(Wi) if Wi is not None else None, st)
    nbytes = 6 * M * 128 * 4
elif kind == "qkv_3x":
    A = mk(M, 128); W = mk(384, 128) / 11; Wi = image(W); Q = torch.empty(M, 384, device=dev)
    fn = lambda: lib.tpc_gemm(h, L.ptr(A), 128, L.ptr(W), 128, L.ptr(Q), 384, M, 384, 128, None, None, 0, 0, None, None, None, 0, 1, L.ptr(Wi), st)
    nbytes = 4 * M * 128 * 4
elif kind == "ffn1_3x":
    A = mk(M, 128); W = mk(512, 128) / 11; Wi = image(W); H = torch.empty(M, 512, device=dev)
    fn = lambda: lib.tpc_gemm(h, L.ptr(A), 128, L.ptr(W), 128, L.ptr(H), 512, M, 512, 128, None, None, 0, 0, None, None, None, 1, 1, L.ptr(Wi), st)
    nbytes = 5 * M * 128 * 4
else:
    A = mk(M, 512); B2 = torch.empty_like(A)
    fn = lambda: B2.copy_(A)
    nbytes = 2 * M * 512 * 4
for _ in range(5): fn()
torch.cuda.synchronize()
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=power.draw,clocks.sm,clocks_throttle_reasons.active", "--format=csv,noheader", "-lms", "100"],
                       stdout=subprocess.PIPE, text=True)
secs = float(os.environ.get("SECS", 3))
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
t0 = time.time(); n = 0
e0.record()
while time.time() - t0 < secs:
    for _ in range(50): fn()
    n += 50
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
smi.terminate()
lines = [l.strip() for l in smi.stdout.read().splitlines() if l.strip()]
us = e0.elapsed_time(e1) / n * 1e3
mid = lines[len(lines) // 3:]
pw = sorted(float(l.split(",")[0].split()[0]) for l in mid)
ck = sorted(int(l.split(",")[1].split()[0]) for l in mid)
rs = sorted(set(l.split(",")[2].strip() for l in mid))
print(f"{kind:8s} zero={int(zero)}: {us:8.1f} us/launch {nbytes/us/1e6:5.2f} TB/s | power median {pw[len(pw)//2]:.0f} W max {pw[-1]:.0f} W | "
      f"SM clock median {ck[len(ck)//2]} MHz min {ck[0]} | throttle reasons {rs}")
