"""Short driver for ncu --set full: the dominant kernels once each at the update-chunk shape (2048 states x 151)."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tpc_b200 import lib as L
lib = L.load()
h = C.c_void_p(); L.check(lib.tpc_create(C.byref(h)))
dev = torch.device("cuda:0")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
M = 2048 * 151
A = torch.randn(M, 128, device=dev); W = torch.randn(128, 128, device=dev) / 11
Wlo = torch.empty_like(W); L.check(lib.tpc_gemm_split_image(h, L.ptr(W), 128, 128, 128, L.ptr(Wlo), st))
D = torch.empty(M, 128, device=dev); res = torch.randn(M, 128, device=dev)
g = torch.ones(128, device=dev); b = torch.zeros(128, device=dev); rstd = torch.empty(M, device=dev)
H = torch.randn(M, 512, device=dev); W1 = torch.randn(512, 128, device=dev) / 11
W1lo = torch.empty_like(W1); L.check(lib.tpc_gemm_split_image(h, L.ptr(W1), 128, 512, 128, L.ptr(W1lo), st))
dW = torch.zeros(128, 128, device=dev); db = torch.zeros(128, device=dev)
for it in range(2):
    # forward projection (split product) + residual + LayerNorm epilogue
    lib.tpc_gemm(h, L.ptr(A), 128, L.ptr(W), 128, L.ptr(D), 128, M, 128, 128, L.ptr(b), L.ptr(res), 128, 1, L.ptr(g), L.ptr(b), L.ptr(rstd), 2, 1, L.ptr(Wlo), st)
    # forward FFN1 (split product) + relu (N=512)
    lib.tpc_gemm(h, L.ptr(A), 128, L.ptr(W1), 128, L.ptr(H), 512, M, 512, 128, None, None, 0, 0, None, None, None, 1, 1, L.ptr(W1lo), st)
    # backward dX (1 pass) with add epilogue
    lib.tpc_gemm(h, L.ptr(A), 128, L.ptr(W), 128, L.ptr(D), 128, M, 128, 128, None, L.ptr(res), 128, 1, None, None, None, 4, 1, None, st)
    # wgrad + bias
    lib.tpc_wgrad(h, L.ptr(A), 128, L.ptr(res), 128, L.ptr(dW), 128, L.ptr(db), M, 128, 128, st)
torch.cuda.synchronize()
print("done")
