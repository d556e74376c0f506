"""Debug: SM-clock trace of CTA 0 of the tcgen05 NT kernel on c3's context projection (M 18944, N 38400, K 2000):
steady-state timing of the ring (TMA issue -> data landed -> split done -> MMA start -> MMAs issued)."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from floppity_deprecated_b200 import _lib
L = _lib.lib()
M, N, K = 18944, int(sys.argv[1]) if len(sys.argv) > 1 else 38400, int(sys.argv[2]) if len(sys.argv) > 2 else 2000
A, B = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
Bh, Bl = torch.empty_like(B), torch.empty_like(B)
L.fp_split_tf32(B.data_ptr(), Bh.data_ptr(), Bl.data_ptr(), B.numel(), _lib.stream())
C = torch.zeros(M, N, device="cuda")
a = _lib.GemmArgs(); a.A, a.lda, a.a_kmajor = A.data_ptr(), K, 1; a.B, a.ldb, a.b_kmajor = B.data_ptr(), K, 1
a.C, a.ldc, a.M, a.N, a.K, a.flags, a.split_k = C.data_ptr(), N, M, N, K, 0, 1
for _ in range(2): L.fp_gemm_tc(ctypes.byref(a), Bh.data_ptr(), Bl.data_ptr(), _lib.stream())
tr = torch.zeros(512, dtype=torch.int64, device="cuda")
L.fp_tc_set_trace(tr.data_ptr())
L.fp_gemm_tc(ctypes.byref(a), Bh.data_ptr(), Bl.data_ptr(), _lib.stream())
torch.cuda.synchronize(); L.fp_tc_set_trace(None)
t = tr.cpu().tolist(); t0 = t[500]
f = lambda v: f"{(v - t0):8d}" if v else "     -- "
print("cycles from kernel start; setup done at", f(t[501]), " kernel end", f(t[502]))
print(" kb | tma issued | data landed | split done | mma start | mma issued |  land-issue  split  mma-wait  per-kb")
prev = None
for kb in range(0, 60):
    e = t[kb * 8: kb * 8 + 5]
    if not all(e): break
    per = (e[4] - prev) if prev else 0
    prev = e[4]
    print(f"{kb:3d} | {f(e[0])} | {f(e[1])} | {f(e[2])} | {f(e[3])} | {f(e[4])} | {e[1]-e[0]:8d} {e[2]-e[1]:6d} {e[3]-e[2]:8d} {per:8d}")
