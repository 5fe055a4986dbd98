"""Stand-alone GPU check of the tcgen05 GEMM / wgrad kernels against torch fp64 (run under gpurun).

Not collected by pytest (no test_ prefix); tests/test_gemm_gpu.py holds the pytest version.
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
lib = ctypes.CDLL(os.path.join(ROOT, "transformer-pointer-critic_b200", "libtpc.so"))
P = ctypes.c_void_p
I = ctypes.c_int
lib.tpc_create.argtypes = [ctypes.POINTER(P)]
lib.tpc_gemm.argtypes = [P, P, I, P, I, P, I, I, I, I, P, P, I, I, P, P, P, I, I, P, P]
lib.tpc_gemm_split_image.argtypes = [P, P, I, I, I, P, P]


def split_image(h, Bt, stream):
    """fp16 split image of the weights (tpc.h: tpc_gemm_split_image), as an fp32-sized buffer"""
    img = torch.empty_like(Bt)
    rc = lib.tpc_gemm_split_image(h, ptr(Bt), Bt.shape[1], Bt.shape[0], Bt.shape[1], ptr(img), stream)
    assert rc == 0, rc
    return img
lib.tpc_wgrad.argtypes = [P, P, I, P, I, P, I, P, I, I, I, P]


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def tf32(x):
    # round-to-nearest-even emulation is close enough to rna for a tolerance check
    return (x.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32) if False else x


def main():
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    h = P()
    rc = lib.tpc_create(ctypes.byref(h))
    assert rc == 0, rc
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    worst = 0.0
    for (M, N, K) in [(128, 128, 128), (300, 128, 128), (1000, 384, 128), (4096, 128, 512), (777, 512, 128),
                      (260, 128, 19328)]:
        A = torch.randn(M, K, device=dev)
        Bt = torch.randn(N, K, device=dev) / K ** 0.5
        bias = torch.randn(N, device=dev)
        D = torch.full((M, N), float("nan"), device=dev)
        rc = lib.tpc_gemm(h, ptr(A), K, ptr(Bt), K, ptr(D), N, M, N, K, ptr(bias), None, 0, 0, None, None, None, 0, 1,
                          None, stream)
        assert rc == 0, rc
        torch.cuda.synchronize()
        ref = (A.double() @ Bt.double().T + bias.double())
        err = (D.double() - ref).abs().max().item() / ref.abs().max().item()
        print(f"gemm  M={M} N={N} K={K}: rel-max err {err:.3e} finite={torch.isfinite(D).all().item()}")
        worst = max(worst, err)

    # split product: fp32-accurate
    for (M, N, K) in [(300, 128, 128), (1000, 384, 128), (4096, 128, 512), (260, 128, 19328)]:
        A = torch.randn(M, K, device=dev)
        Bt = torch.randn(N, K, device=dev) / K ** 0.5
        Blo = split_image(h, Bt, stream)
        bias = torch.randn(N, device=dev)
        D = torch.full((M, N), float("nan"), device=dev)
        rc = lib.tpc_gemm(h, ptr(A), K, ptr(Bt), K, ptr(D), N, M, N, K, ptr(bias), None, 0, 0, None, None, None, 0, 1,
                          ptr(Blo), stream)
        assert rc == 0, rc
        torch.cuda.synchronize()
        ref = (A.double() @ Bt.double().T + bias.double())
        err = (D.double() - ref).abs().max().item() / ref.abs().max().item()
        print(f"gemm3x M={M} N={N} K={K}: rel-max err {err:.3e}")
        assert err < (2e-5 if K <= 512 else 3e-4), err  # long-K: fp32 accumulation order inside the tensor core

    # weight-resident kernel (K == 128, many row tiles): 1-pass and split product, bias + relu, aux add / gate
    for (M, N) in [(40000, 128), (39999, 384), (40001, 512)]:
        K = 128
        A = torch.randn(M, K, device=dev)
        Bt = torch.randn(N, K, device=dev) / K ** 0.5
        Blo = split_image(h, Bt, stream)
        bias = torch.randn(N, device=dev)
        aux = torch.randn(M, N, device=dev)
        ref0 = A.double() @ Bt.double().T
        for split in (False, True):
            D = torch.full((M, N), float("nan"), device=dev)
            rc = lib.tpc_gemm(h, ptr(A), K, ptr(Bt), K, ptr(D), N, M, N, K, ptr(bias), None, 0, 0, None, None, None, 1, 1,
                              ptr(Blo) if split else None, stream)
            assert rc == 0, rc
            torch.cuda.synchronize()
            ref = torch.relu(ref0 + bias.double())
            err = (D.double() - ref).abs().max().item() / ref.abs().max().item()
            print(f"bres M={M} N={N} split={split} bias+relu: {err:.3e}")
            assert err < (2e-5 if split else 2e-3), err
        D = torch.full((M, N), float("nan"), device=dev)
        rc = lib.tpc_gemm(h, ptr(A), K, ptr(Bt), K, ptr(D), N, M, N, K, None, ptr(aux), N, 1, None, None, None, 4, 1, None, stream)
        assert rc == 0
        torch.cuda.synchronize()
        err = (D.double() - (ref0 + aux.double())).abs().max().item() / ref0.abs().max().item()
        print(f"bres M={M} N={N} aux add: {err:.3e}")
        assert err < 2e-3, err
        D = aux.clone()
        rc = lib.tpc_gemm(h, ptr(A), K, ptr(Bt), K, ptr(D), N, M, N, K, None, ptr(D), N, 2, None, None, None, 4, 1, None, stream)
        assert rc == 0
        torch.cuda.synchronize()
        err = (D.double() - ref0 * (aux.double() > 0)).abs().max().item() / ref0.abs().max().item()
        print(f"bres M={M} N={N} gate in place: {err:.3e}")
        assert err < 2e-3, err

    # relu + aux add + round
    M, N, K = 555, 256, 128
    A = torch.randn(M, K, device=dev)
    Bt = torch.randn(N, K, device=dev) / K ** 0.5
    aux = torch.randn(M, N, device=dev)
    D = torch.empty(M, N, device=dev)
    rc = lib.tpc_gemm(h, ptr(A), K, ptr(Bt), K, ptr(D), N, M, N, K, None, ptr(aux), N, 1, None, None, None, 1, 1, None, stream)
    assert rc == 0
    torch.cuda.synchronize()
    ref = torch.relu(A.double() @ Bt.double().T) + aux.double()
    err = (D.double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"gemm relu+add: {err:.3e}")
    worst = max(worst, err)
    # gate in place
    Hm = torch.randn(M, N, device=dev)
    D = Hm.clone()
    rc = lib.tpc_gemm(h, ptr(A), K, ptr(Bt), K, ptr(D), N, M, N, K, None, ptr(D), N, 2, None, None, None, 0, 1, None, stream)
    assert rc == 0
    torch.cuda.synchronize()
    ref = (A.double() @ Bt.double().T) * (Hm.double() > 0)
    err = (D.double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"gemm gate in-place: {err:.3e}")
    worst = max(worst, err)

    # residual + layernorm
    M, N, K = 1000, 128, 512
    A = torch.randn(M, K, device=dev)
    Bt = torch.randn(N, K, device=dev) / K ** 0.5
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev)
    gamma = torch.rand(N, device=dev) + 0.5
    beta = torch.randn(N, device=dev)
    D = torch.empty(M, N, device=dev)
    rstd = torch.empty(M, device=dev)
    rc = lib.tpc_gemm(h, ptr(A), K, ptr(Bt), K, ptr(D), N, M, N, K, ptr(bias), ptr(res), N, 1, ptr(gamma), ptr(beta),
                      ptr(rstd), 2 | 4, 1, None, stream)
    assert rc == 0
    torch.cuda.synchronize()
    pre = A.double() @ Bt.double().T + bias.double() + res.double()
    ref = torch.nn.functional.layer_norm(pre, (N,), gamma.double(), beta.double(), 1e-6)
    err = (D.double() - ref).abs().max().item() / ref.abs().max().item()
    rref = 1.0 / torch.sqrt(pre.var(dim=1, unbiased=False) + 1e-6)
    rerr = ((rstd.double() - rref).abs() / rref).max().item()
    print(f"gemm +res+LN: {err:.3e}  rstd {rerr:.3e}")
    worst = max(worst, err)

    # split-K atomic
    M, N, K = 300, 128, 19328
    A = torch.randn(M, K, device=dev)
    Bt = torch.randn(N, K, device=dev) / K ** 0.5
    D = torch.zeros(M, N, device=dev)
    rc = lib.tpc_gemm(h, ptr(A), K, ptr(Bt), K, ptr(D), N, M, N, K, None, None, 0, 0, None, None, None, 8, 40, None, stream)
    assert rc == 0
    torch.cuda.synchronize()
    ref = A.double() @ Bt.double().T
    err = (D.double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"gemm split-K atomic: {err:.3e}")
    worst = max(worst, err)

    for (M, Kin, Nout) in [(256, 128, 128), (5000, 128, 384), (100000, 512, 128), (3000, 128, 512), (999, 1280, 128)]:
        X = torch.randn(M, Kin, device=dev)
        dY = torch.randn(M, Nout, device=dev)
        dW = torch.zeros(Kin, Nout, device=dev)
        db = torch.zeros(Nout, device=dev)
        rc = lib.tpc_wgrad(h, ptr(X), Kin, ptr(dY), Nout, ptr(dW), Nout, ptr(db), M, Kin, Nout, stream)
        assert rc == 0, rc
        torch.cuda.synchronize()
        ref = X.double().T @ dY.double()
        err = (dW.double() - ref).abs().max().item() / ref.abs().max().item()
        rb = dY.double().sum(0)
        errb = (db.double() - rb).abs().max().item() / rb.abs().max().item()
        print(f"wgrad M={M} Kin={Kin} Nout={Nout}: rel-max err {err:.3e}  bias {errb:.3e}")
        worst = max(worst, err, errb)

    # timing of the common shape
    M, N, K = 154624, 128, 128
    A = torch.randn(M, K, device=dev)
    Bt = torch.randn(N, K, device=dev)
    D = torch.empty(M, N, device=dev)
    for _ in range(3):
        lib.tpc_gemm(h, ptr(A), K, ptr(Bt), K, ptr(D), N, M, N, K, None, None, 0, 0, None, None, None, 4, 1, None, stream)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(20):
        lib.tpc_gemm(h, ptr(A), K, ptr(Bt), K, ptr(D), N, M, N, K, None, None, 0, 0, None, None, None, 4, 1, None, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"gemm {M}x{N}x{K}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s  {(M*K+M*N)*4/ms/1e6:.0f} GB/s")
    print("WORST", worst)
    assert worst < 2e-3, worst
    print("OK")


if __name__ == "__main__":
    main()


def check_split_image_and_range():
    """tpc_gemm_split_image is bit-exact against the same two-term FP16 split in torch; operands beyond the fp16 range
    saturate (finite output, no NaN); tiny operands keep an absolute error far below the fp32 rounding of the result."""
    dev = torch.device("cuda:0")
    h = P()
    assert lib.tpc_create(ctypes.byref(h)) == 0
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    torch.manual_seed(5)
    N, K = 384, 128
    Bt = torch.randn(N, K, device=dev) * torch.logspace(-9, 1, K, device=dev)   # columns from 1e-9 to 10
    Bt[0, 0], Bt[1, 1], Bt[2, 2] = 7.0e4, -1.0e6, 65504.0                      # beyond / at the fp16 range
    img = split_image(h, Bt, stream)
    torch.cuda.synchronize()
    planes = img.view(torch.float16).reshape(N, 2 * K)   # fp32 row of K floats == 2K halves; the image is [2][N][K] halves
    planes = img.view(torch.float16).reshape(-1)[: 2 * N * K].reshape(2, N, K)
    hi = Bt.clamp(-65504.0, 65504.0).to(torch.float16)
    lo = (Bt - hi.float()).clamp(-65504.0, 65504.0).to(torch.float16)
    assert torch.equal(planes[0], hi), "B_h plane differs from fp16(B)"
    assert torch.equal(planes[1], lo), "B_r plane differs from fp16(B - B_h)"
    out = {}
    # saturation: huge activations and weights give a finite (wrong, documented) result, never NaN / Inf
    M = 256
    A = torch.randn(M, K, device=dev)
    A[3, 5] = 3.0e5
    D = torch.full((M, N), float("nan"), device=dev)
    rc = lib.tpc_gemm(h, ptr(A), K, ptr(Bt), K, ptr(D), N, M, N, K, None, None, 0, 0, None, None, None, 0, 1, ptr(img), stream)
    assert rc == 0, rc
    torch.cuda.synchronize()
    assert torch.isfinite(D).all(), "split product produced NaN / Inf on out-of-range operands"
    # tiny operands: activations of 1e-6 against unit weights
    A2 = torch.randn(M, K, device=dev) * 1e-6
    B2 = torch.randn(N, K, device=dev)
    img2 = split_image(h, B2, stream)
    rc = lib.tpc_gemm(h, ptr(A2), K, ptr(B2), K, ptr(D), N, M, N, K, None, None, 0, 0, None, None, None, 0, 1, ptr(img2), stream)
    assert rc == 0, rc
    torch.cuda.synchronize()
    ref = A2.double() @ B2.double().T
    out["tiny_abs"] = (D.double() - ref).abs().max().item()
    out["tiny_rel"] = out["tiny_abs"] / ref.abs().max().item()
    print(f"split image bit-exact; saturation finite; tiny operands: abs err {out['tiny_abs']:.2e} (rel {out['tiny_rel']:.2e})")
    # |a| ~ 1e-6 sits below the fp16 normal range: each element carries an absolute error <= 3e-8, i.e. up to a few 1e-3
    # relative on such a product -- documented limit of the split; the bound checked is the absolute one
    assert out["tiny_abs"] < 128 ** 0.5 * 4 * 3e-8 * 4, out
    lib.tpc_destroy(h)
    return out


lib.tpc_gemm_ln_bwd.argtypes = [P, P, I, P, I, P, I, I, I, P, I, P, P, P, P, P, P, P]


def check_gemm_ln_bwd(timing=False):
    """tpc_gemm_ln_bwd (backward GEMM whose epilogue is the LayerNorm backward) against torch autograd in fp64:
    pre -> y = LayerNorm(pre); dy = A @ Bt^T + aux; the kernel must return d(pre), and add d(gamma), d(beta)."""
    dev = torch.device("cuda:0")
    h = P()
    assert lib.tpc_create(ctypes.byref(h)) == 0
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    torch.manual_seed(11)
    out = {}
    for (M, K, with_aux) in [(128, 128, True), (1000, 512, True), (40001, 384, True), (777, 128, False), (19000, 512, False)]:
        pre = (torch.randn(M, 128, device=dev) * 2 + 0.3).double().requires_grad_(True)
        gamma = (1 + 0.2 * torch.randn(128, device=dev)).double().requires_grad_(True)
        beta = (0.1 * torch.randn(128, device=dev)).double().requires_grad_(True)
        y64 = torch.nn.functional.layer_norm(pre, (128,), gamma, beta, 1e-6)
        rstd = (1.0 / torch.sqrt(pre.var(dim=1, unbiased=False) + 1e-6)).float().contiguous()
        A = torch.randn(M, K, device=dev)
        Bt = torch.randn(128, K, device=dev) / K ** 0.5
        aux = torch.randn(M, 128, device=dev) if with_aux else None
        dy = A.double() @ Bt.double().T + (aux.double() if with_aux else 0.0)
        y64.backward(dy)
        y = y64.detach().float().contiguous()
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        D = torch.full((M, 128), float("nan"), device=dev)
        dg = torch.zeros(128, device=dev)
        db = torch.zeros(128, device=dev)
        rc = lib.tpc_gemm_ln_bwd(h, ptr(A), K, ptr(Bt), K, ptr(D), 128, M, K, ptr(aux), 128, ptr(y), ptr(rstd), ptr(g32),
                                 ptr(b32), ptr(dg), ptr(db), stream)
        assert rc == 0, rc
        torch.cuda.synchronize()
        assert torch.isfinite(D).all()
        e_d = ((D.double() - pre.grad).abs().max() / pre.grad.abs().max()).item()
        e_g = ((dg.double() - gamma.grad).abs().max() / gamma.grad.abs().max()).item()
        e_b = ((db.double() - beta.grad).abs().max() / beta.grad.abs().max()).item()
        print(f"gemm + LN backward M={M} K={K} aux={with_aux}: dpre {e_d:.3e} dgamma {e_g:.3e} dbeta {e_b:.3e}")
        out[(M, K, with_aux)] = (e_d, e_g, e_b)
        # one TF32 pass on unrounded operands + TF32-rounded output: the class of the other backward GEMMs (2e-3)
        assert e_d < 3e-3 and e_g < 3e-3 and e_b < 3e-3, (M, K, e_d, e_g, e_b)
    if timing:
        for K in (512, 384):
            M = 1546240
            A = torch.randn(M, K, device=dev)
            Bt = torch.randn(128, K, device=dev) / K ** 0.5
            aux = torch.randn(M, 128, device=dev)
            y = torch.randn(M, 128, device=dev)
            rstd = torch.rand(M, device=dev) + 0.5
            g32 = torch.ones(128, device=dev)
            b32 = torch.zeros(128, device=dev)
            D = torch.empty(M, 128, device=dev)
            dg = torch.zeros(128, device=dev)
            db = torch.zeros(128, device=dev)

            def fused():
                lib.tpc_gemm_ln_bwd(h, ptr(A), K, ptr(Bt), K, ptr(D), 128, M, K, ptr(aux), 128, ptr(y), ptr(rstd), ptr(g32),
                                    ptr(b32), ptr(dg), ptr(db), stream)

            def plain():
                lib.tpc_gemm(h, ptr(A), K, ptr(Bt), K, ptr(D), 128, M, 128, K, None, ptr(aux), 128, 1, None, None, None, 4, 1,
                             None, stream)

            for name, fn in (("gemm + aux add (unfused producer)", plain), ("gemm + aux add + LN backward epilogue", fused)):
                for _ in range(3):
                    fn()
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                e0.record()
                for _ in range(10):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / 10 * 1e3
                byts = M * (K + 128 * (2 if fn is plain else 3)) * 4   # A + aux + D (+ the saved LayerNorm output)
                print(f"  K={K} M={M} {name}: {us:.0f} us, {byts / us / 1e6:.2f} TB/s of its own bytes")
    lib.tpc_destroy(h)
    return out
