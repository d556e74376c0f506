"""The first conditioner GEMM of a c3 layer (K = 30 parameters, theta rows padded to 32 floats) and its weight gradient on the
3xFP16 tensor-core kernels, against the fp32 SIMT kernel: time per launch, with the kernels' roles ablated."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from floppity_deprecated_b200 import _lib
from tools.gemm16_check import timeit
L = _lib.lib(); s = _lib.stream()
R, D, Dp, H, NC = 18944, 30, 32, 512, 38400
th = torch.randn(R, D, device="cuda"); thp = torch.zeros(R, Dp, device="cuda"); thp[:, :D] = th
W = torch.randn(H, D, device="cuda") / D ** 0.5; Wp = torch.zeros(H, Dp, device="cuda"); Wp[:, :D] = W
h16 = torch.empty(H, Dp, device="cuda", dtype=torch.float16); l16 = torch.empty_like(h16)
L.fp_split_f16(Wp.data_ptr(), h16.data_ptr(), l16.data_ptr(), Wp.numel(), s)
ctx_all = torch.randn(R, NC, device="cuda"); out = torch.empty(R, H, device="cuda")
for flags, name in ((_lib.G_ADD_CTX, "add_ctx"), (0, "plain")):
    a = _lib.GemmArgs()
    a.A, a.lda, a.a_kmajor, a.B, a.ldb, a.b_kmajor = thp.data_ptr(), Dp, 1, Wp.data_ptr(), Dp, 1
    a.C, a.ldc, a.M, a.N, a.K, a.flags, a.split_k = out.data_ptr(), H, R, H, D, flags, 1
    a.G, a.ldg, a.ctx_div = ctx_all.data_ptr() + 4 * 1024, NC, 1
    b = _lib.GemmArgs()
    b.A, b.lda, b.a_kmajor, b.B, b.ldb, b.b_kmajor = th.data_ptr(), D, 1, W.data_ptr(), D, 1
    b.C, b.ldc, b.M, b.N, b.K, b.flags, b.split_k = out.data_ptr(), H, R, H, D, flags, 1
    b.G, b.ldg, b.ctx_div = ctx_all.data_ptr() + 4 * 1024, NC, 1
    res = {"simt": round(timeit(lambda: L.fp_gemm(ctypes.byref(b), s)), 1)}
    for bits in (0, 1, 2, 3):
        L.fp_tc_set_debug(bits)
        r = L.fp_gemm_tc16(ctypes.byref(a), h16.data_ptr(), l16.data_ptr(), s); assert r == 0, r
        res[f"tc16 dbg{bits}"] = round(timeit(lambda: L.fp_gemm_tc16(ctypes.byref(a), h16.data_ptr(), l16.data_ptr(), s)), 1)
    L.fp_tc_set_debug(0)
    print("NT first layer", name, res, flush=True)
dh = torch.randn(R, H, device="cuda") * 1e-4; dW = torch.zeros(H, D, device="cuda"); dWp = torch.zeros(H, Dp, device="cuda")
for ldc, C, name in ((D, dW, "ldc 30"), (Dp, dWp, "ldc 32")):
    t = _lib.GemmArgs()
    t.A, t.lda, t.a_kmajor, t.B, t.ldb, t.b_kmajor = dh.data_ptr(), H, 0, thp.data_ptr(), Dp, 0
    t.C, t.ldc, t.M, t.N, t.K, t.flags, t.split_k = C.data_ptr(), ldc, H, D if ldc == D else Dp, R, _lib.G_ATOMIC, 1
    res = {}
    for bits in (0, 1, 2, 3):
        L.fp_tc_set_debug(bits)
        r = L.fp_gemm_tc16_tn(ctypes.byref(t), s); assert r == 0, r
        res[f"tn16 dbg{bits}"] = round(timeit(lambda: L.fp_gemm_tc16_tn(ctypes.byref(t), s)), 1)
    L.fp_tc_set_debug(0)
    print("TN first layer", name, res, flush=True)
u = _lib.GemmArgs()
u.A, u.lda, u.a_kmajor, u.B, u.ldb, u.b_kmajor = dh.data_ptr(), H, 0, th.data_ptr(), D, 0
u.C, u.ldc, u.M, u.N, u.K, u.flags, u.split_k = dW.data_ptr(), D, H, D, R, _lib.G_ATOMIC, 74
print("TN first layer simt split_k 74:", round(timeit(lambda: L.fp_gemm(ctypes.byref(u), s)), 1))
