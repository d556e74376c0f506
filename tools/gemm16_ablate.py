"""Timing ablation of the 3xFP16 NT GEMM (debug bits as in c3_gemm_ablate.py: 1 no conversion, 2 no epilogue stores, 64 one product)."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from floppity_deprecated_b200 import _lib
L = _lib.lib(); s = _lib.stream()
from tools.gemm16_check import timeit


def nt(M, N, K, flags=0):
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda") / K ** 0.5
    h16 = torch.empty(N, K, device="cuda", dtype=torch.float16); l16 = torch.empty_like(h16)
    L.fp_split_f16(B.data_ptr(), h16.data_ptr(), l16.data_ptr(), B.numel(), s)
    C = torch.empty(M, N, device="cuda"); bias = torch.randn(N, device="cuda")
    b = _lib.GemmArgs()
    b.A, b.lda, b.a_kmajor, b.B, b.ldb, b.b_kmajor = A.data_ptr(), K, 1, B.data_ptr(), K, 1
    b.C, b.ldc, b.M, b.N, b.K, b.flags, b.bias, b.split_k = C.data_ptr(), N, M, N, K, flags | _lib.G_BIAS, bias.data_ptr(), 1
    out = {}
    for bits in (0, 1, 2, 64, 65, 67):
        L.fp_tc_set_debug(bits)
        out[bits] = round(timeit(lambda: L.fp_gemm_tc16(ctypes.byref(b), h16.data_ptr(), l16.data_ptr(), s)), 1)
    L.fp_tc_set_debug(0)
    return out


def tn(M, N, R, flags=0):
    A = torch.randn(R, M, device="cuda") * 1e-3; B = torch.randn(R, N, device="cuda")
    C = torch.zeros(M, N, device="cuda"); db = torch.zeros(M, device="cuda")
    sc = torch.tensor([512.0], device="cuda")
    b = _lib.GemmArgs()
    b.A, b.lda, b.a_kmajor, b.B, b.ldb, b.b_kmajor = A.data_ptr(), M, 0, B.data_ptr(), N, 0
    b.C, b.ldc, b.M, b.N, b.K, b.split_k = C.data_ptr(), N, M, N, R, 1
    b.flags, b.U, b.a_scale = _lib.G_ATOMIC | _lib.G_COLSUM_A | flags, db.data_ptr(), sc.data_ptr()
    out = {}
    for bits in (0, 1, 2, 3, 64, 65, 67):
        L.fp_tc_set_debug(bits)
        out[bits] = round(timeit(lambda: L.fp_gemm_tc16_tn(ctypes.byref(b), s)), 1)
    L.fp_tc_set_debug(0)
    return out


if __name__ == "__main__":
    R = 18944
    print("NT16 block", nt(R, 512, 512, _lib.G_RELU_A), flush=True)
    print("NT16 ctx  ", nt(R, 38400, 2000), flush=True)
    print("TN16 block [512 x 512] over 18944, relu_b", tn(512, 512, R, _lib.G_RELU_B), flush=True)
    print("TN16 final [708 x 512] over 18944        ", tn(708, 512, R), flush=True)
