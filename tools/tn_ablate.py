"""Timing ablation of the tcgen05 weight-gradient (TN) kernel: which role bounds a k-block?
bits: 1 = splitter does no work, 2 = epilogue skips the global reductions, 4 = two of three products."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from floppity_deprecated_b200 import _lib
L = _lib.lib()
R = 40960


def run(M, N, lda, ldb, colsum):
    A, B = torch.randn(R, lda, device="cuda"), torch.randn(R, ldb, device="cuda")
    C = torch.zeros(M, N, device="cuda"); bsum = torch.zeros(((M + 3) // 4) * 4, device="cuda")
    a = _lib.GemmArgs()
    a.A, a.lda, a.a_kmajor = A.data_ptr(), lda, 0
    a.B, a.ldb, a.b_kmajor = B.data_ptr(), ldb, 0
    a.C, a.ldc, a.M, a.N, a.K = C.data_ptr(), N, M, N, R
    a.flags, a.split_k = _lib.G_ATOMIC | _lib.G_RELU_B | (_lib.G_COLSUM_A if colsum else 0), 1
    a.U = bsum.data_ptr()
    s = _lib.stream()
    out = {}
    for bits in (0, 1, 2, 3, 4, 7):
        L.fp_tc_set_debug(bits)
        for _ in range(3):
            assert L.fp_gemm_tc_tn(ctypes.byref(a), s) == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            L.fp_gemm_tc_tn(ctypes.byref(a), s)
        e1.record(); torch.cuda.synchronize()
        out[bits] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
    L.fp_tc_set_debug(0)
    return out


for name, M, N, lda, ldb in [("w 128x128", 128, 128, 128, 128), ("final 290x128", 290, 128, 292, 128), ("first 128x20", 128, 20, 128, 20),
                             ("ctx 3840x500 (R=4096)", 3840, 500, 3840, 500)]:
    if name.startswith("ctx"):
        R = 4096
    for cs in (0, 1):
        print(name, "colsum" if cs else "plain ", run(M, N, lda, ldb, cs), flush=True)
