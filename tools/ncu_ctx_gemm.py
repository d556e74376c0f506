"""The two GEMMs that carry a third of the c3 training step, launched three times each (for an `ncu --set full` capture of
one launch of known shape), as the engine launches them: the context projection [18944 x 2000] x [38400 x 2000]^T
(k_gemm_tc_nt, bias epilogue, 3xFP16 pieces, context rows pre-split) and its weight gradient [38400 x 2000] over 18944 rows
(k_gemm_tc_tn, gradient operand converted in the kernel and scaled, context rows pre-split, bias gradient riding along).
`python tools/ncu_ctx_gemm.py tf32` launches the 3xTF32 kernels instead."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from floppity_deprecated_b200 import _lib
L = _lib.lib(); s = _lib.stream()
tf32 = len(sys.argv) > 1 and sys.argv[1] == "tf32"
R, C, N = 18944, 2000, 38400
x = torch.randn(R, C, device="cuda"); W = torch.randn(N, C, device="cuda") / C ** 0.5
out = torch.empty(R, N, device="cuda"); bias = torch.randn(N, device="cuda")
dW = torch.zeros(N, C, device="cuda"); db = torch.zeros(N, device="cuda")
a = _lib.GemmArgs()
a.A, a.lda, a.a_kmajor, a.B, a.ldb, a.b_kmajor = x.data_ptr(), C, 1, W.data_ptr(), C, 1
a.C, a.ldc, a.M, a.N, a.K, a.flags, a.bias, a.split_k = out.data_ptr(), N, R, N, C, _lib.G_BIAS, bias.data_ptr(), 1
b = _lib.GemmArgs()
b.A, b.lda, b.a_kmajor, b.B, b.ldb, b.b_kmajor = out.data_ptr(), N, 0, x.data_ptr(), C, 0
b.C, b.ldc, b.M, b.N, b.K, b.flags, b.split_k, b.U = dW.data_ptr(), C, N, C, R, _lib.G_ATOMIC | _lib.G_COLSUM_A, 1, db.data_ptr()
if tf32:
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    L.fp_split_tf32(W.data_ptr(), hi.data_ptr(), lo.data_ptr(), W.numel(), s)
    for _ in range(3):
        assert L.fp_gemm_tc(ctypes.byref(a), hi.data_ptr(), lo.data_ptr(), s) == 0
        assert L.fp_gemm_tc_tn(ctypes.byref(b), s) == 0
else:
    f16 = lambda t: (torch.empty_like(t, dtype=torch.float16), torch.empty_like(t, dtype=torch.float16))
    (wh, wl), (xh, xl) = f16(W), f16(x)
    L.fp_split_f16(W.data_ptr(), wh.data_ptr(), wl.data_ptr(), W.numel(), s)
    L.fp_split_f16(x.data_ptr(), xh.data_ptr(), xl.data_ptr(), x.numel(), s)
    sc = torch.tensor([1.0], device="cuda")
    a.pre_hi16, a.pre_lo16, a.pre_ld = xh.data_ptr(), xl.data_ptr(), C
    b.pre_hi16, b.pre_lo16, b.pre_ld = xh.data_ptr(), xl.data_ptr(), C
    b.a_scale = sc.data_ptr()
    for _ in range(3):
        assert L.fp_gemm_tc16(ctypes.byref(a), wh.data_ptr(), wl.data_ptr(), s) == 0
        assert L.fp_gemm_tc16_tn(ctypes.byref(b), s) == 0
torch.cuda.synchronize()
print("algorithmic bytes: nt", 4 * (R * C + N * C + R * N + N), " tn", 4 * (R * N + R * C + 2 * N * C + N))
