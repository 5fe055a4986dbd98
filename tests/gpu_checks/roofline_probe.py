"""The roofline launch of bench.py (gemm_tc_kernel, FFN2 + residual + LayerNorm at one update chunk) on its own, for
`ncu --set full -k regex:gemm_tc -c 1`: dram__bytes_read.sum + dram__bytes_write.sum go to profiles/r02f_ncu_full_gemm_tc_ffn2.json (round 1: profiles/gemm_tc_traffic.json)."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from tpc_b200 import lib as L
lib = L.load()
h = C.c_void_p(); L.check(lib.tpc_create(C.byref(h)))
dev = torch.device("cuda:0")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
M = int(os.environ.get("M", 10240 * 151))
call, alg, label = bench.roofline_probe(lib, h, st, M, dev)
for _ in range(3):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(10):
    call()
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10 / 1e3
print(f"{label}: M={M} {t * 1e6:.1f} us/launch, algorithmic {alg / 1e6:.1f} MB -> {alg / t / 1e9:.0f} GB/s")
