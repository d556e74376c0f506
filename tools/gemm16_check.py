"""3xFP16 against 3xTF32 in the stand-alone tensor-core GEMMs on the c3 shapes: error against an fp64 product (relative to the
largest output element) and time per launch.  NT: C = pro(A) B^T (+bias); TN: C += A^T B (weight-gradient form)."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from floppity_deprecated_b200 import _lib
L = _lib.lib(); s = _lib.stream()


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


def nt(M, N, K, flags=0, a_mag=1.0, scale=None, ldb16=None, check=True, lda=None):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    lda = lda or K
    Afull = torch.randn(M, lda, device="cuda", generator=g) * a_mag
    A = Afull[:, :K]
    ld16 = ldb16 or K
    Bp = torch.zeros(N, ld16, device="cuda"); Bp[:, :K] = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    B = Bp[:, :K]
    hi, lo = torch.empty_like(Bp), torch.empty_like(Bp)
    L.fp_split_tf32(Bp.data_ptr(), hi.data_ptr(), lo.data_ptr(), Bp.numel(), s)
    h16 = torch.empty(N, ld16, device="cuda", dtype=torch.float16); l16 = torch.empty_like(h16)
    assert L.fp_split_f16(Bp.data_ptr(), h16.data_ptr(), l16.data_ptr(), Bp.numel(), s) == 0
    bias = torch.randn(N, device="cuda", generator=g)
    st = torch.zeros(4, dtype=torch.int32, device="cuda")
    sc = None if scale is None else torch.tensor([scale], device="cuda")
    out = {}
    C32, C16 = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")
    a = _lib.GemmArgs()
    a.A, a.lda, a.a_kmajor, a.B, a.ldb, a.b_kmajor = A.data_ptr(), lda, 1, B.data_ptr(), ld16, 1
    a.C, a.ldc, a.M, a.N, a.K, a.flags, a.bias, a.split_k = C32.data_ptr(), N, M, N, K, flags | _lib.G_BIAS, bias.data_ptr(), 1
    b = _lib.GemmArgs()
    b.A, b.lda, b.a_kmajor, b.B, b.ldb, b.b_kmajor = A.data_ptr(), lda, 1, B.data_ptr(), ld16, 1
    b.C, b.ldc, b.M, b.N, b.K, b.flags, b.bias, b.split_k = C16.data_ptr(), N, M, N, K, flags | _lib.G_BIAS, bias.data_ptr(), 1
    b.status = st.data_ptr()
    if sc is not None:
        b.a_scale = sc.data_ptr()
    f32 = lambda: L.fp_gemm_tc(ctypes.byref(a), hi.data_ptr(), lo.data_ptr(), s)
    f16 = lambda: L.fp_gemm_tc16(ctypes.byref(b), h16.data_ptr(), l16.data_ptr(), s)
    assert f32() == 0
    r = f16(); assert r == 0, r
    torch.cuda.synchronize()
    out["status"] = st.tolist()
    if check:
        rows = slice(0, min(M, 4096))
        Ad = A[rows].double()
        if flags & _lib.G_RELU_A:
            Ad = Ad.clamp_min(0)
        ref = Ad @ B.double().T + bias.double()
        den = ref.abs().max().item()
        out["err_tf32"] = ((C32[rows].double() - ref).abs().max() / den).item()
        out["err_fp16"] = ((C16[rows].double() - ref).abs().max() / den).item()
        # last rows too (ragged tiles)
        Ad = A[-300:].double().clamp_min(0) if flags & _lib.G_RELU_A else A[-300:].double()
        ref = Ad @ B.double().T + bias.double()
        out["err_fp16_tail"] = ((C16[-300:].double() - ref).abs().max() / ref.abs().max()).item()
    out["us_tf32"] = round(timeit(f32), 1); out["us_fp16"] = round(timeit(f16), 1)
    out["tflops_fp16"] = round(2.0 * M * N * K / out["us_fp16"] / 1e6, 1)
    return out


def tn(M, N, R, flags=0, a_mag=1.0, scale=None, colsum=True):
    g = torch.Generator(device="cuda").manual_seed(M + N + R)
    A = torch.randn(R, M, device="cuda", generator=g) * a_mag       # dOut
    B = torch.randn(R, N, device="cuda", generator=g)               # activations
    st = torch.zeros(4, dtype=torch.int32, device="cuda")
    sc = None if scale is None else torch.tensor([scale], device="cuda")
    C32, C16 = torch.zeros(M, N, device="cuda"), torch.zeros(M, N, device="cuda")
    d32, d16 = torch.zeros(M, device="cuda"), torch.zeros(M, device="cuda")
    fl = _lib.G_ATOMIC | flags | (_lib.G_COLSUM_A if colsum else 0)
    a = _lib.GemmArgs()
    a.A, a.lda, a.a_kmajor, a.B, a.ldb, a.b_kmajor = A.data_ptr(), M, 0, B.data_ptr(), N, 0
    a.C, a.ldc, a.M, a.N, a.K, a.flags, a.split_k, a.U = C32.data_ptr(), N, M, N, R, fl, 1, d32.data_ptr()
    b = _lib.GemmArgs()
    b.A, b.lda, b.a_kmajor, b.B, b.ldb, b.b_kmajor = A.data_ptr(), M, 0, B.data_ptr(), N, 0
    b.C, b.ldc, b.M, b.N, b.K, b.flags, b.split_k, b.U = C16.data_ptr(), N, M, N, R, fl, 1, d16.data_ptr()
    b.status = st.data_ptr()
    if sc is not None:
        b.a_scale = sc.data_ptr()
    f32 = lambda: L.fp_gemm_tc_tn(ctypes.byref(a), s)
    f16 = lambda: L.fp_gemm_tc16_tn(ctypes.byref(b), s)
    assert f32() == 0
    r = f16(); assert r == 0, r
    torch.cuda.synchronize()
    out = {"status": st.tolist()}
    Bd = B.double().clamp_min(0) if flags & _lib.G_RELU_B else B.double()
    cols = slice(0, min(M, 1024))
    ref = A[:, cols].double().T @ Bd
    den = ref.abs().max().item()
    out["err_tf32"] = ((C32[cols].double() - ref).abs().max() / den).item()
    out["err_fp16"] = ((C16[cols].double() - ref).abs().max() / den).item()
    if colsum:
        cref = A.double().sum(0)
        out["colsum_err_tf32"] = ((d32.double() - cref).abs().max() / cref.abs().max()).item()
        out["colsum_err_fp16"] = ((d16.double() - cref).abs().max() / cref.abs().max()).item()
    out["us_tf32"] = round(timeit(f32), 1); out["us_fp16"] = round(timeit(f16), 1)
    out["tflops_fp16"] = round(2.0 * M * N * R / out["us_fp16"] / 1e6, 1)
    return out


if __name__ == "__main__":
    R = 18944
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    print("NT small   [300 x 96] x [200 x 96]^T            ", nt(300, 200, 96), flush=True)
    print("NT ragged  [1000 x 712(705)] x [512 x 705]^T     ", nt(1000, 512, 705, ldb16=712, lda=708), flush=True)
    print("NT block   [18944 x 512] x [512 x 512]^T relu   ", nt(R, 512, 512, _lib.G_RELU_A), flush=True)
    print("NT final   [18944 x 512] x [708 x 512]^T        ", nt(R, 708, 512), flush=True)
    print("NT dgrad   tiny A (1e-7), scale 2^20            ", nt(R, 512, 512, 0, a_mag=1e-7, scale=2.0 ** 20), flush=True)
    print("NT dgrad   tiny A (1e-7), no scale              ", nt(R, 512, 512, 0, a_mag=1e-7), flush=True)
    print("NT range   A ~ 1e5 (guard must fire)            ", nt(1000, 512, 512, 0, a_mag=1e5), flush=True)
    print("TN small   [200 x 136] over 1000 rows            ", tn(200, 136, 1000), flush=True)
    print("TN block   [512 x 512] over 18944 rows, relu_b   ", tn(512, 512, R, _lib.G_RELU_B), flush=True)
    print("TN final   [708 x 512] over 18944 rows           ", tn(708, 512, R), flush=True)
    print("TN tiny A (1e-7) scale 2^20                      ", tn(512, 512, R, _lib.G_RELU_B, a_mag=1e-7, scale=2.0 ** 20), flush=True)
    print("TN ragged  [705 x 510(ld 512)]: skipped (ld must be x4)", flush=True)
    if which == "all":
        print("TN ctx wgrad [38400 x 2000] over 18944 rows     ", tn(38400, 2000, R), flush=True)
        # the same two context GEMMs with the shared operand pre-split by the caller
        A = torch.randn(R, 2000, device="cuda"); B = torch.randn(38400, 2000, device="cuda") / 2000 ** 0.5
        ah, al = torch.empty(R, 2000, device="cuda", dtype=torch.float16), torch.empty(R, 2000, device="cuda", dtype=torch.float16)
        bh, bl = torch.empty_like(B, dtype=torch.float16), torch.empty_like(B, dtype=torch.float16)
        L.fp_split_f16(A.data_ptr(), ah.data_ptr(), al.data_ptr(), A.numel(), s)
        L.fp_split_f16(B.data_ptr(), bh.data_ptr(), bl.data_ptr(), B.numel(), s)
        C = torch.empty(R, 38400, device="cuda"); bias = torch.randn(38400, device="cuda")
        a = _lib.GemmArgs()
        a.A, a.lda, a.a_kmajor, a.B, a.ldb, a.b_kmajor = A.data_ptr(), 2000, 1, B.data_ptr(), 2000, 1
        a.C, a.ldc, a.M, a.N, a.K, a.flags, a.bias, a.split_k = C.data_ptr(), 38400, R, 38400, 2000, _lib.G_BIAS, bias.data_ptr(), 1
        a.pre_hi16, a.pre_lo16, a.pre_ld = ah.data_ptr(), al.data_ptr(), 2000
        print("NT ctx fwd, A pre-split: us", round(timeit(lambda: L.fp_gemm_tc16(ctypes.byref(a), bh.data_ptr(), bl.data_ptr(), s)), 1),
              " split of A: us", round(timeit(lambda: L.fp_split_f16(A.data_ptr(), ah.data_ptr(), al.data_ptr(), A.numel(), s)), 1), flush=True)
        dW = torch.zeros(38400, 2000, device="cuda"); db = torch.zeros(38400, device="cuda")
        b = _lib.GemmArgs()
        b.A, b.lda, b.a_kmajor, b.B, b.ldb, b.b_kmajor = C.data_ptr(), 38400, 0, A.data_ptr(), 2000, 0
        b.C, b.ldc, b.M, b.N, b.K, b.split_k = dW.data_ptr(), 2000, 38400, 2000, R, 1
        b.flags, b.U = _lib.G_ATOMIC | _lib.G_COLSUM_A, db.data_ptr()
        b.pre_hi16, b.pre_lo16, b.pre_ld = ah.data_ptr(), al.data_ptr(), 2000
        print("TN ctx wgrad, B pre-split: us", round(timeit(lambda: L.fp_gemm_tc16_tn(ctypes.byref(b), s)), 1), flush=True)
        print("NT ctx fwd [18944 x 2000] x [38400 x 2000]^T    ", nt(R, 38400, 2000), flush=True)
