"""Timing ablation of the two GEMMs that carry half of the c3 training step (context projection forward, NT, and its weight
gradient, TN) plus the hidden-512 block GEMMs.  Debug bits (fp_tc_set_debug; results are wrong while set):
1 = splitter warps do no work, 2 = epilogue skips its global stores / reductions, 4 = two of three products,
8 = (NT) the B_lo tile is not loaded, 32 = (NT) CTA pairs (cta_group::2; results stay correct), 128 = the tile order that ignores the L2 (results stay correct)."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from floppity_deprecated_b200 import _lib
L = _lib.lib()
s = _lib.stream()


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def nt(M, N, K, flags=0):
    A, B = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    hi, lo = torch.empty_like(B), torch.empty_like(B)
    L.fp_split_tf32(B.data_ptr(), hi.data_ptr(), lo.data_ptr(), B.numel(), s)
    C = torch.empty(M, N, device="cuda"); bias = torch.randn(N, device="cuda")
    a = _lib.GemmArgs()
    a.A, a.lda, a.a_kmajor, a.B, a.ldb, a.b_kmajor = A.data_ptr(), K, 1, B.data_ptr(), K, 1
    a.C, a.ldc, a.M, a.N, a.K, a.flags, a.bias, a.split_k = C.data_ptr(), N, M, N, K, flags | _lib.G_BIAS, bias.data_ptr(), 1
    out = {}
    for bits in (0, 128, 1, 4, 15, 32):
        L.fp_tc_set_debug(bits)
        us = timeit(lambda: L.fp_gemm_tc(ctypes.byref(a), hi.data_ptr(), lo.data_ptr(), s))
        out[bits] = (round(us, 1), round(2.0 * M * N * K / us / 1e6, 1))
    L.fp_tc_set_debug(0)
    return out


def tn(M, N, R):
    A, B = torch.randn(R, M, device="cuda"), torch.randn(R, N, device="cuda")
    C = torch.zeros(M, N, device="cuda")
    a = _lib.GemmArgs()
    a.A, a.lda, a.a_kmajor, a.B, a.ldb, a.b_kmajor = A.data_ptr(), M, 0, B.data_ptr(), N, 0
    a.C, a.ldc, a.M, a.N, a.K, a.flags, a.split_k = C.data_ptr(), N, M, N, R, _lib.G_ATOMIC, 1
    out = {}
    for bits in (0, 128, 1, 4, 7):
        L.fp_tc_set_debug(bits)
        us = timeit(lambda: L.fp_gemm_tc_tn(ctypes.byref(a), s))
        out[bits] = (round(us, 1), round(2.0 * M * N * R / us / 1e6, 1))
    L.fp_tc_set_debug(0)
    return out


R = 18944
print("(us, TFLOP/s fp32-equivalent) per debug-bit set")
print("NT ctx fwd   [18944 x 2000] x [38400 x 2000]^T", nt(R, 38400, 2000), flush=True)
print("NT block     [18944 x 512]  x [512 x 512]^T   ", nt(R, 512, 512, _lib.G_RELU_A), flush=True)
print("NT final     [18944 x 512]  x [705 x 512]^T   ", nt(R, 708, 512), flush=True)
print("TN ctx wgrad [38400 x 2000] over 18944 rows   ", tn(38400, 2000, R), flush=True)
print("TN block     [512 x 512] over 18944 rows      ", tn(512, 512, R), flush=True)
