"""Live CUDA-event timings of the dominant kernels at the update-chunk shape."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tpc_b200 import lib as L
lib = L.load()
h = C.c_void_p(); L.check(lib.tpc_create(C.byref(h)))
dev = torch.device("cuda:0")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
C_ = int(os.environ.get("CHUNK", 2048))
M = C_ * 151
def split_image(Wt):
    img = torch.empty_like(Wt)
    L.check(lib.tpc_gemm_split_image(h, L.ptr(Wt), Wt.shape[1], Wt.shape[0], Wt.shape[1], L.ptr(img), st))
    return img
A = torch.randn(M, 128, device=dev); W = torch.randn(128, 128, device=dev) / 11
Wlo = split_image(W)
D = torch.empty(M, 128, device=dev); res = torch.randn(M, 128, device=dev)
g = torch.ones(128, device=dev); b = torch.zeros(128, device=dev); rstd = torch.empty(M, device=dev)
H = torch.randn(M, 512, device=dev); W1 = torch.randn(512, 128, device=dev) / 11
W1lo = split_image(W1)
W2 = torch.randn(128, 512, device=dev) / 22
W2lo = split_image(W2)
Q = torch.randn(M, 384, device=dev); Wq = torch.randn(384, 128, device=dev) / 11
Wqlo = split_image(Wq)
dW = torch.zeros(128, 128, device=dev); db = torch.zeros(128, device=dev)
dW4 = torch.zeros(128, 512, device=dev); db4 = torch.zeros(512, device=dev)
dW5 = torch.zeros(512, 128, device=dev)
def timeit(name, fn, bytes_):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    print(f"{name:44s} {us:8.1f} us  {bytes_/us/1e6:7.2f} TB/s (algorithmic)")
n4 = M * 128 * 4
timeit("fwd 3x proj 128->128", lambda: lib.tpc_gemm(h, L.ptr(A), 128, L.ptr(W), 128, L.ptr(D), 128, M, 128, 128, None, None, 0, 0, None, None, None, 0, 1, L.ptr(Wlo), st), 2 * n4)
timeit("fwd 3x proj +res+LN", lambda: lib.tpc_gemm(h, L.ptr(A), 128, L.ptr(W), 128, L.ptr(D), 128, M, 128, 128, L.ptr(b), L.ptr(res), 128, 1, L.ptr(g), L.ptr(b), L.ptr(rstd), 2, 1, L.ptr(Wlo), st), 3 * n4)
timeit("fwd 3x QKV 128->384", lambda: lib.tpc_gemm(h, L.ptr(A), 128, L.ptr(Wq), 128, L.ptr(Q), 384, M, 384, 128, None, None, 0, 0, None, None, None, 0, 1, L.ptr(Wqlo), st), 4 * n4)
timeit("fwd 3x FFN1 128->512 relu", lambda: lib.tpc_gemm(h, L.ptr(A), 128, L.ptr(W1), 128, L.ptr(H), 512, M, 512, 128, None, None, 0, 0, None, None, None, 1, 1, L.ptr(W1lo), st), 5 * n4)
timeit("fwd 3x FFN2 512->128 +res+LN", lambda: lib.tpc_gemm(h, L.ptr(H), 512, L.ptr(W2), 512, L.ptr(D), 128, M, 128, 512, L.ptr(b), L.ptr(res), 128, 1, L.ptr(g), L.ptr(b), L.ptr(rstd), 2, 1, L.ptr(W2lo), st), 6 * n4)
timeit("bwd 1x dX 128->128", lambda: lib.tpc_gemm(h, L.ptr(A), 128, L.ptr(W), 128, L.ptr(D), 128, M, 128, 128, None, None, 0, 0, None, None, None, 4, 1, None, st), 2 * n4)
timeit("bwd 1x dX +add", lambda: lib.tpc_gemm(h, L.ptr(A), 128, L.ptr(W), 128, L.ptr(D), 128, M, 128, 128, None, L.ptr(res), 128, 1, None, None, None, 4, 1, None, st), 3 * n4)
timeit("bwd 1x dH gate 128->512 (in place)", lambda: lib.tpc_gemm(h, L.ptr(A), 128, L.ptr(W1), 128, L.ptr(H), 512, M, 512, 128, None, L.ptr(H), 512, 2, None, None, None, 4, 1, None, st), 9 * n4)
timeit("bwd 1x dout1 512->128 +add", lambda: lib.tpc_gemm(h, L.ptr(H), 512, L.ptr(W2), 512, L.ptr(D), 128, M, 128, 512, None, L.ptr(res), 128, 1, None, None, None, 4, 1, None, st), 6 * n4)
timeit("bwd 1x dx 384->128 +add", lambda: lib.tpc_gemm(h, L.ptr(Q), 384, L.ptr(W1), 384, L.ptr(D), 128, M, 128, 384, None, L.ptr(res), 128, 1, None, None, None, 4, 1, None, st), 5 * n4)
timeit("wgrad 128x128 + bias", lambda: lib.tpc_wgrad(h, L.ptr(A), 128, L.ptr(res), 128, L.ptr(dW), 128, L.ptr(db), M, 128, 128, st), 2 * n4)
timeit("wgrad 128x128 no bias", lambda: lib.tpc_wgrad(h, L.ptr(A), 128, L.ptr(res), 128, L.ptr(dW), 128, None, M, 128, 128, st), 2 * n4)
timeit("wgrad 128x384 + bias (qkv)", lambda: lib.tpc_wgrad(h, L.ptr(A), 128, L.ptr(Q), 384, L.ptr(dW4), 512, L.ptr(db4), M, 128, 384, st), 4 * n4)
timeit("wgrad 128x512 + bias (ffn1)", lambda: lib.tpc_wgrad(h, L.ptr(A), 128, L.ptr(H), 512, L.ptr(dW4), 512, L.ptr(db4), M, 128, 512, st), 5 * n4)
timeit("wgrad 512x128 + bias (ffn2)", lambda: lib.tpc_wgrad(h, L.ptr(H), 512, L.ptr(res), 128, L.ptr(dW5), 128, L.ptr(db), M, 512, 128, st), 5 * n4)
x = torch.empty(M * 128, device=dev); y = torch.empty_like(x)
timeit("torch copy (HBM reference)", lambda: y.copy_(x), 2 * n4)
