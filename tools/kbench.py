#!/usr/bin/env python
"""Kernel micro-benchmarks at the BASELINE.json shapes: spline forward / inverse / backward (HBM roofline) and the
conditioner GEMM shapes (tensor roofline).  Inputs are larger than the 126 MB L2; CUDA-event timing, 3 warm-ups.

    python tools/kbench.py [rqs] [gemm]
"""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from floppity_deprecated_b200 import _lib  # noqa: E402

PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
    else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def rqs():
    L = _lib.lib()
    out = []
    only = os.environ.get("KB_SHAPE")
    for name, (D, K, H) in {"c1": (10, 8, 50), "c2": (20, 10, 128), "c3": (30, 16, 512), "c4": (40, 10, 128)}.items():
        if only and name != only:
            continue
        cfg = _lib.NsfConfig(D=D, C=8, T=2, K=K, H=H, Bk=1, tail_bound=3.0, min_bin_width=1e-3, min_bin_height=1e-3,
                             min_derivative=1e-3, lu_eps=1e-3, dropout_p=0.0)
        d_tr, P = D // 2, 3 * K - 1
        R = int(400e6 / (d_tr * P * 4)) // 1024 * 1024          # ~400 MB of params >> L2
        g = torch.Generator(device="cuda").manual_seed(0)
        theta = torch.rand(R, D, device="cuda", generator=g) * 6.2 - 3.1
        params = torch.randn(R, d_tr, P, device="cuda", generator=g)
        y, ld = torch.empty_like(theta), torch.zeros(R, device="cuda")
        dth, dp = torch.empty_like(theta), torch.empty_like(params)
        gy, gl = torch.randn(R, D, device="cuda", generator=g), torch.randn(R, device="cuda", generator=g)
        s = _lib.stream()
        f = lambda: L.fp_rqs_forward(cfg, 0, theta.data_ptr(), params.data_ptr(), y.data_ptr(), ld.data_ptr(), None, R, None, s)
        i = lambda: L.fp_rqs_inverse(cfg, 0, theta.data_ptr(), params.data_ptr(), y.data_ptr(), ld.data_ptr(), None, R, None, s)
        b = lambda: L.fp_rqs_backward(cfg, 0, theta.data_ptr(), params.data_ptr(), gy.data_ptr(), gl.data_ptr(), dp.data_ptr(), dth.data_ptr(), R, s)
        by_f = 4.0 * R * (d_tr * P + 2 * d_tr + 1)
        by_b = 4.0 * R * (2 * d_tr * P + 3 * d_tr + 1)
        for tag, fn, by in (("fwd", f, by_f), ("inv", i, by_f), ("bwd", b, by_b)):
            ms = timeit(fn)
            gbs = by / ms / 1e6
            out.append(dict(kernel=f"rqs_{tag}", shape=name, rows=R, ms=round(ms, 4), gbs=round(gbs, 1), frac_hbm=round(gbs / PK["hbm_gbs"], 3)))
            print(out[-1], flush=True)
    return out


def gemm():
    L = _lib.lib()
    out = []
    shapes = [("ctx_fwd_c2", 4096, 3840, 500, 1, 1), ("blk_fwd_c2", 40960, 128, 128, 1, 1), ("final_fwd_c2", 40960, 290, 128, 1, 1),
              ("blk_dgrad_c2", 40960, 128, 128, 1, 0), ("blk_wgrad_c2", 128, 128, 40960, 0, 0), ("ctx_wgrad_c2", 3840, 500, 4096, 0, 0),
              ("ctx_fwd_c3", 4096, 38400, 2000, 1, 1), ("blk_fwd_c3", 4096, 512, 512, 1, 1)]
    for name, M, N, K, ak, bk in shapes:
        A = torch.randn((M, K) if ak else (K, M), device="cuda")
        B = torch.randn((N, K) if bk else (K, N), device="cuda")
        C = torch.zeros(M, N, device="cuda")
        a = _lib.GemmArgs()
        a.A, a.lda, a.a_kmajor = A.data_ptr(), A.shape[1], ak
        a.B, a.ldb, a.b_kmajor = B.data_ptr(), B.shape[1], bk
        a.C, a.ldc, a.M, a.N, a.K = C.data_ptr(), N, M, N, K
        split = 1
        if not ak:
            tiles = ((M + 127) // 128) * ((N + 63) // 64)
            split = max(1, min((K + 255) // 256, (2 * 148 + tiles - 1) // tiles))
        a.flags, a.split_k = (_lib.G_ATOMIC if split > 1 else 0), split
        s = _lib.stream()
        ms = timeit(lambda: L.fp_gemm(ctypes.byref(a), s))
        tf = 2.0 * M * N * K / ms / 1e9
        out.append(dict(kernel="gemm_simt", shape=name, M=M, N=N, K=K, ms=round(ms, 4), tflops=round(tf, 2)))
        print(out[-1], flush=True)
    return out


def gemm_tc():
    L = _lib.lib()
    out = []
    shapes = [("ctx_fwd_c2", 4096, 3840, 500), ("blk_fwd_c2", 40960, 128, 128), ("final_fwd_c2", 40960, 292, 128),
              ("first_fwd_c2", 40960, 128, 20), ("ctx_fwd_c3", 4096, 38400, 2000), ("blk_fwd_c3", 4096, 512, 512),
              ("blk_fwd_c3_big", 65536, 512, 512)]
    only = os.environ.get("KB_SHAPE")
    for name, M, N, K in shapes:
        if only and name != only:
            continue
        A, B = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
        Bh, Bl = torch.empty_like(B), torch.empty_like(B)
        L.fp_split_tf32(B.data_ptr(), Bh.data_ptr(), Bl.data_ptr(), B.numel(), _lib.stream())
        C = torch.zeros(M, N, device="cuda")
        a = _lib.GemmArgs()
        a.A, a.lda, a.a_kmajor = A.data_ptr(), K, 1
        a.B, a.ldb, a.b_kmajor = B.data_ptr(), K, 1
        a.C, a.ldc, a.M, a.N, a.K, a.flags, a.split_k = C.data_ptr(), N, M, N, K, 0, 1
        s = _lib.stream()
        rc = L.fp_gemm_tc(ctypes.byref(a), Bh.data_ptr(), Bl.data_ptr(), s)
        if rc != 0:
            print(name, "declined", rc)
            continue
        ms = timeit(lambda: L.fp_gemm_tc(ctypes.byref(a), Bh.data_ptr(), Bl.data_ptr(), s))
        tf = 2.0 * M * N * K / ms / 1e9
        out.append(dict(kernel="gemm_tcgen05_3xtf32", shape=name, M=M, N=N, K=K, ms=round(ms, 4), tflops_fp32_equiv=round(tf, 2),
                        tf32_tflops=round(3 * tf, 2), frac_bf16_burst=round(tf / PK["bf16_tflops"], 4)))
        print(out[-1], flush=True)
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["rqs", "gemm", "tc"]
    res = []
    if "rqs" in which:
        res += rqs()
    if "gemm" in which:
        res += gemm()
    if "tc" in which:
        res += gemm_tc()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "kbench.json"), "w"), indent=1)
