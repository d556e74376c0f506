"""Debug: per-event SM-clock trace of CTA 0 of the tcgen05 NT kernel on the skinny block shape."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from floppity_deprecated_b200 import _lib
L = _lib.lib()
M, N, K = int(sys.argv[1]) if len(sys.argv) > 1 else 40960, 128, int(sys.argv[2]) if len(sys.argv) > 2 else 128
A, B = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
Bh, Bl = torch.empty_like(B), torch.empty_like(B)
L.fp_split_tf32(B.data_ptr(), Bh.data_ptr(), Bl.data_ptr(), B.numel(), _lib.stream())
C = torch.zeros(M, N, device="cuda")
a = _lib.GemmArgs(); a.A, a.lda, a.a_kmajor = A.data_ptr(), K, 1; a.B, a.ldb, a.b_kmajor = B.data_ptr(), K, 1
a.C, a.ldc, a.M, a.N, a.K, a.flags, a.split_k = C.data_ptr(), N, M, N, K, 0, 1
mode = sys.argv[3] if len(sys.argv) > 3 else "plain"
if mode == "glu":      # W2 forward: bias + GLU gate + residual, U stored
    G = torch.randn(M // 10 + 1, 3840, device="cuda"); Rsd = torch.randn(M, N, device="cuda"); U = torch.empty(M, N, device="cuda")
    bias = torch.randn(N, device="cuda")
    a.flags = _lib.G_BIAS | _lib.G_GLU_RES | _lib.G_RELU_A
    a.bias, a.G, a.ldg, a.ctx_div, a.Rsd, a.ldr, a.U, a.ldu = bias.data_ptr(), G.data_ptr(), 3840, 10, Rsd.data_ptr(), N, U.data_ptr(), N
elif mode == "maskacc":  # W1 dgrad: relu mask + accumulate
    Msk = torch.randn(M, N, device="cuda")
    a.flags = _lib.G_RELUMASK | _lib.G_ACCUM
    a.Msk, a.ldm = Msk.data_ptr(), N
for _ in range(3): L.fp_gemm_tc(ctypes.byref(a), Bh.data_ptr(), Bl.data_ptr(), _lib.stream())
tr = torch.zeros(512, dtype=torch.int64, device="cuda")
L.fp_tc_set_trace(tr.data_ptr())
L.fp_gemm_tc(ctypes.byref(a), Bh.data_ptr(), Bl.data_ptr(), _lib.stream())
torch.cuda.synchronize(); L.fp_tc_set_trace(None)
t = tr.cpu().tolist(); t0 = t[500]
f = lambda v: f"{(v - t0) / 1.9e3:7.2f}" if v else "   --  "
print("setup done at", f(t[501]), "us; kernel end", f(t[502]), "(us, assuming 1.9 GHz)")
print(" kb | tma issued | data landed | split done | mma start | mma issued+commit")
for kb in range(0, 16):
    e = t[kb * 8: kb * 8 + 5]
    if not any(e): break
    print(f"{kb:3d} | {f(e[0])} | {f(e[1])} | {f(e[2])} | {f(e[3])} | {f(e[4])}")
for i in range(4):
    if t[480 + 2 * i]: print(f"epilogue tile {i}: start {f(t[480 + 2 * i])} end {f(t[481 + 2 * i])}")
for ti in range(3):
    base = 400 + ti * 16
    if not t[base]: break
    print(f"tile {ti}: per chunk [start | tmem->slab done | b0 loads issued | b0 first row stored | b0 done | b1 loads issued | b1 first row | b1 done]")
    for ch in range(2):
        print("    ", " | ".join(f(t[base + ch * 8 + j]) for j in range(8)))
