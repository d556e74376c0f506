"""Debug: SM-clock trace of CTA 0 of the fused conditioner kernel (c2 shapes), eval and training mode."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import floppity_deprecated_b200 as fb
from floppity_deprecated_b200 import _lib
L = _lib.lib()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 40960
torch.manual_seed(0)
net = fb.NSFFlow(20, 500, hidden_features=128, num_transforms=10, num_bins=10, num_blocks=2).cuda()
theta = torch.rand(R, 20, device="cuda") * 4 - 2
x = torch.randn(R // 10, 500, device="cuda")
for mode in ("eval", "train"):
    training = mode == "train"
    for _ in range(2):
        net._forward_raw(theta, x, training)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); net._forward_raw(theta, x, training); e1.record(); torch.cuda.synchronize()
    print(mode, "whole forward (10 layers + ctx GEMM + splines):", round(e0.elapsed_time(e1) * 1e3, 1), "us")
    tr = torch.zeros(512, dtype=torch.int64, device="cuda")
    L.fp_cf_set_trace(tr.data_ptr())
    net._forward_raw(theta, x, training)
    torch.cuda.synchronize(); L.fp_cf_set_trace(None)
    t = tr.cpu().tolist()       # last layer's launch overwrote earlier ones: stamps of layer T-1
    t0 = min(v for v in t if v)
    f = lambda v: f"{(v - t0) / 1.9e3:7.2f}" if v else "   --  "
    print(mode, "MMA thread [operand ready | unit committed] per unit, first 2 tiles:")
    m = [v for v in t[:256] if v]
    print("   ", " ".join(f(v) for v in m[:48]))
    print(mode, "epilogue warp [acc ready | handed on], first 2 tiles:")
    e = [v for v in t[256:] if v]
    print("   ", " ".join(f(v) for v in e[:48]))
    print("    last stamp", f(max(t)))
