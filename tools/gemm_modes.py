#!/usr/bin/env python
"""CUDA-event timing of every (shape, epilogue) the c2 training step sends to the tcgen05 GEMM kernels.

    python tools/gemm_modes.py [R]        # R = flow rows (default 40960 = batch 4096 x atoms 10)
"""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from floppity_deprecated_b200 import _lib  # noqa: E402

L = _lib.lib()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 40960
H, D, P, LDP, NCTX = 128, 20, 290, 292, 3840
dev = "cuda"
keep = []


def T(*shape):
    t = torch.randn(*shape, device=dev)
    keep.append(t)
    return t


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def nt(name, M, N, K, lda, ldc, flags, extra_bytes=0, no_u=False, **aux):
    A = T(M, lda)
    B = T(N, lda)           # weight [N, K] with the same leading dimension
    Bh, Bl = torch.empty_like(B), torch.empty_like(B)
    keep.extend([Bh, Bl])
    L.fp_split_tf32(B.data_ptr(), Bh.data_ptr(), Bl.data_ptr(), B.numel(), _lib.stream())
    C = T(M, ldc)
    a = _lib.GemmArgs()
    a.A, a.lda, a.a_kmajor = A.data_ptr(), lda, 1
    a.B, a.ldb, a.b_kmajor = B.data_ptr(), lda, 1
    a.C, a.ldc, a.M, a.N, a.K, a.flags, a.split_k = C.data_ptr(), ldc, M, N, K, flags, 1
    if flags & _lib.G_BIAS:
        a.bias = T(max(N, 4)).data_ptr()
    if flags & (_lib.G_ADD_CTX | _lib.G_GLU_RES):
        a.G, a.ldg, a.ctx_div = T(M // 10 + 1, NCTX).data_ptr(), NCTX, 10
    if flags & _lib.G_GLU_RES:
        a.Rsd, a.ldr = T(M, ldc).data_ptr(), ldc
        if not no_u:
            a.U, a.ldu = T(M, ldc).data_ptr(), ldc
    if flags & _lib.G_RELUMASK:
        a.Msk, a.ldm = T(M, ldc).data_ptr(), ldc
    s = _lib.stream()
    rc = L.fp_gemm_tc(ctypes.byref(a), Bh.data_ptr(), Bl.data_ptr(), s)
    if rc != 0:
        print(name, "declined", rc)
        return None
    us = timeit(lambda: L.fp_gemm_tc(ctypes.byref(a), Bh.data_ptr(), Bl.data_ptr(), s))
    by = 4.0 * (M * K + M * N) + extra_bytes
    return dict(kernel="nt", name=name, M=M, N=N, K=K, flags=flags, us=round(us, 1), gbs=round(by / us / 1e3, 0),
                tflops=round(2.0 * M * N * K / us / 1e6, 1))


def tn(name, M, N, K, lda, ldb, relu_b=False):
    A = T(K, lda)
    B = T(K, ldb)
    C = torch.zeros(M, N, device=dev)
    keep.append(C)
    a = _lib.GemmArgs()
    a.A, a.lda, a.a_kmajor = A.data_ptr(), lda, 0
    a.B, a.ldb, a.b_kmajor = B.data_ptr(), ldb, 0
    a.C, a.ldc, a.M, a.N, a.K = C.data_ptr(), N, M, N, K
    a.flags, a.split_k = _lib.G_ATOMIC | (_lib.G_RELU_B if relu_b else 0), 1
    s = _lib.stream()
    rc = L.fp_gemm_tc_tn(ctypes.byref(a), s)
    if rc != 0:
        print(name, "declined", rc)
        return None
    us = timeit(lambda: L.fp_gemm_tc_tn(ctypes.byref(a), s))
    by = 4.0 * (K * M + K * N)
    return dict(kernel="tn", name=name, M=M, N=N, K=K, us=round(us, 1), gbs=round(by / us / 1e3, 0),
                tflops=round(2.0 * M * N * K / us / 1e6, 1))


G = _lib
res = [
    nt("first_fwd  [R,D]x[H,D] +ctx", R, H, D, D, H, G.G_ADD_CTX),
    nt("w1_fwd     relu,bias", R, H, H, H, H, G.G_RELU_A | G.G_BIAS),
    nt("w2_fwd     relu,bias,glu", R, H, H, H, H, G.G_RELU_A | G.G_BIAS | G.G_GLU_RES, extra_bytes=4.0 * 2 * R * H),
    nt("final_fwd  bias N=290", R, P, H, H, LDP, G.G_BIAS),
    nt("final_dgrad K=290", R, H, P, LDP, H, 0),
    nt("w2_dgrad   relumask", R, H, H, H, H, G.G_RELUMASK, extra_bytes=4.0 * R * H),
    nt("w1_dgrad   relumask,accum", R, H, H, H, H, G.G_RELUMASK | G.G_ACCUM, extra_bytes=4.0 * 2 * R * H),
    nt("first_dgrad accum N=20", R, D, H, H, D, G.G_ACCUM, extra_bytes=4.0 * R * D),
    nt("x: glu without U", R, H, H, H, H, G.G_RELU_A | G.G_BIAS | G.G_GLU_RES, extra_bytes=4.0 * R * H, no_u=True),
    nt("x: accum only", R, H, H, H, H, G.G_ACCUM, extra_bytes=4.0 * R * H),
    nt("x: add_ctx K=128", R, H, H, H, H, G.G_ADD_CTX),
    nt("x: plain", R, H, H, H, H, 0),
    tn("w_wgrad    128x128", H, H, R, H, H, True),
    tn("final_wgrad 290x128", P, H, R, LDP, H),
    tn("first_wgrad 128x20", H, D, R, H, D),
    tn("lu_wgrad   20x20", D, D, R, D, D),
]
res = [r for r in res if r]
for r in res:
    print(r, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gemm_modes.json"), "w"), indent=1)
