"""Does the leading dimension (row alignment) of the spline-parameter buffers matter to the TMA-fed GEMMs?"""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from floppity_deprecated_b200 import _lib
L = _lib.lib()
R, H, P = 40960, 128, 290
s = _lib.stream()

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / iters * 1e3, 1)

for ld in (292, 296, 304, 320):
    # wgrad: dWf[P, H] += dprm[R, ld]^T h[R, H]
    A, B = torch.randn(R, ld, device="cuda"), torch.randn(R, H, device="cuda")
    C = torch.zeros(P, H, device="cuda")
    a = _lib.GemmArgs()
    a.A, a.lda, a.a_kmajor = A.data_ptr(), ld, 0
    a.B, a.ldb, a.b_kmajor = B.data_ptr(), H, 0
    a.C, a.ldc, a.M, a.N, a.K = C.data_ptr(), H, P, H, R
    a.flags, a.split_k = _lib.G_ATOMIC, 1
    assert L.fp_gemm_tc_tn(ctypes.byref(a), s) == 0
    t_w = timeit(lambda: L.fp_gemm_tc_tn(ctypes.byref(a), s))
    # dgrad: dh[R, H] = dprm[R, ld(K=P)] Wf^T   (B = Wf^T [H, ld])
    Bt = torch.randn(H, ld, device="cuda"); Bh, Bl = torch.empty_like(Bt), torch.empty_like(Bt)
    L.fp_split_tf32(Bt.data_ptr(), Bh.data_ptr(), Bl.data_ptr(), Bt.numel(), s)
    Cd = torch.empty(R, H, device="cuda")
    d = _lib.GemmArgs()
    d.A, d.lda, d.a_kmajor = A.data_ptr(), ld, 1
    d.B, d.ldb, d.b_kmajor = Bt.data_ptr(), ld, 1
    d.C, d.ldc, d.M, d.N, d.K, d.flags, d.split_k = Cd.data_ptr(), H, R, H, P, 0, 1
    assert L.fp_gemm_tc(ctypes.byref(d), Bh.data_ptr(), Bl.data_ptr(), s) == 0
    t_d = timeit(lambda: L.fp_gemm_tc(ctypes.byref(d), Bh.data_ptr(), Bl.data_ptr(), s))
    # forward: prm[R, ld] = h[R, H] Wf[P, H]^T + b
    W = torch.randn(P, H, device="cuda"); Wh, Wl = torch.empty_like(W), torch.empty_like(W)
    L.fp_split_tf32(W.data_ptr(), Wh.data_ptr(), Wl.data_ptr(), W.numel(), s)
    bias = torch.randn(320, device="cuda")
    f = _lib.GemmArgs()
    f.A, f.lda, f.a_kmajor = B.data_ptr(), H, 1
    f.B, f.ldb, f.b_kmajor = W.data_ptr(), H, 1
    f.C, f.ldc, f.M, f.N, f.K, f.flags, f.split_k = A.data_ptr(), ld, R, P, H, _lib.G_BIAS, 1
    f.bias = bias.data_ptr()
    assert L.fp_gemm_tc(ctypes.byref(f), Wh.data_ptr(), Wl.data_ptr(), s) == 0
    t_f = timeit(lambda: L.fp_gemm_tc(ctypes.byref(f), Wh.data_ptr(), Wl.data_ptr(), s))
    print(f"ld {ld}: final wgrad {t_w} us, final dgrad {t_d} us, final fwd {t_f} us", flush=True)
