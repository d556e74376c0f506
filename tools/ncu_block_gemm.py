"""The hidden-512 block GEMMs of a c3 layer, launched three times each for an `ncu --set full` capture: forward linear with the
GLU + residual epilogue (k_gemm_tc_nt, A converted in the kernel) and the weight gradient of a block linear (k_gemm_tc_tn,
both operands converted, relu prologue on the activations, bias gradient riding along)."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from floppity_deprecated_b200 import _lib
L = _lib.lib(); s = _lib.stream()
R, H, NC = 18944, 512, 38400
t = torch.randn(R, H, device="cuda"); W = torch.randn(H, H, device="cuda") / H ** 0.5
wh, wl = torch.empty(H, H, device="cuda", dtype=torch.float16), torch.empty(H, H, device="cuda", dtype=torch.float16)
L.fp_split_f16(W.data_ptr(), wh.data_ptr(), wl.data_ptr(), W.numel(), s)
h = torch.randn(R, H, device="cuda"); hn = torch.empty(R, H, device="cuda"); u = torch.empty(R, H, device="cuda")
ctx_all = torch.randn(R, NC, device="cuda"); bias = torch.randn(H, device="cuda")
a = _lib.GemmArgs()
a.A, a.lda, a.a_kmajor, a.B, a.ldb, a.b_kmajor = t.data_ptr(), H, 1, W.data_ptr(), H, 1
a.C, a.ldc, a.M, a.N, a.K, a.split_k = hn.data_ptr(), H, R, H, H, 1
a.flags, a.bias = _lib.G_RELU_A | _lib.G_BIAS | _lib.G_GLU_RES, bias.data_ptr()
a.G, a.ldg, a.ctx_div = ctx_all.data_ptr() + 4 * 2048, NC, 1
a.Rsd, a.ldr, a.U, a.ldu = h.data_ptr(), H, u.data_ptr(), H
du = torch.randn(R, H, device="cuda") * 1e-4; dW = torch.zeros(H, H, device="cuda"); db = torch.zeros(H, device="cuda")
sc = torch.tensor([4096.0], device="cuda")
b = _lib.GemmArgs()
b.A, b.lda, b.a_kmajor, b.B, b.ldb, b.b_kmajor = du.data_ptr(), H, 0, t.data_ptr(), H, 0
b.C, b.ldc, b.M, b.N, b.K, b.split_k = dW.data_ptr(), H, H, H, R, 1
b.flags, b.U, b.a_scale = _lib.G_ATOMIC | _lib.G_COLSUM_A | _lib.G_RELU_B, db.data_ptr(), sc.data_ptr()
for _ in range(3):
    assert L.fp_gemm_tc16(ctypes.byref(a), wh.data_ptr(), wl.data_ptr(), s) == 0
    assert L.fp_gemm_tc16_tn(ctypes.byref(b), s) == 0
torch.cuda.synchronize()
print("ok")
