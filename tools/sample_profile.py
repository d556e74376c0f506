"""Per-kernel CUDA-event table of posterior sampling (Flow.sample inverse pass) and log_prob on a c2 / c1 flow."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import floppity_deprecated_b200 as fb
from floppity_deprecated_b200 import _lib, synthetic
L = _lib.lib()
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = synthetic.CONFIGS[name]
theta, x, _, _ = synthetic.make_dataset(name, 2048)
net = fb.build_nsf(torch.tensor(theta), torch.tensor(x), hidden_features=cfg.hidden, num_transforms=cfg.transforms, num_bins=cfg.bins,
                   num_blocks=cfg.blocks).cuda().eval()
n = 1 << 20
z = torch.randn(n, cfg.D, device="cuda")
ctx = net._standardize(torch.tensor(x[:1]).cuda())
NT = L.fp_prof_ntags()
def table(fn, reps=5):
    arr = lambda t: (t * NT)()
    ms, cnt, fl, by = arr(ctypes.c_double), arr(ctypes.c_int64), arr(ctypes.c_double), arr(ctypes.c_double)
    for _ in range(2): fn()
    L.fp_prof_enable(1); L.fp_prof_collect(ms, cnt, fl, by)
    for _ in range(reps): fn()
    L.fp_prof_collect(ms, cnt, fl, by); L.fp_prof_enable(0)
    tot = sum(ms)
    print(f"  total {tot / reps:.3f} ms per call of {n} rows -> {n * reps / tot / 1e3:.1f} M rows/s (sum of kernel times)")
    for i in range(NT):
        if cnt[i]:
            print(f"   {L.fp_prof_tag_name(i).decode():18s} share {ms[i] / tot:.3f}  n/call {cnt[i] / reps:5.1f}  avg {1e3 * ms[i] / cnt[i]:8.1f} us"
                  + (f"  {by[i] / ms[i] / 1e6:7.0f} GB/s" if by[i] else ""))
print(name, "sampling (inverse pass):")
table(lambda: net.inverse_rows(z, ctx))
th = torch.tensor(theta[:1]).cuda().repeat(n, 1) + 0.1 * torch.randn(n, cfg.D, device="cuda")
print(name, "log_prob:")
with torch.no_grad():
    table(lambda: net.log_prob(th, torch.tensor(x[:1]).cuda()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(5): net.inverse_rows(z, ctx)
e1.record(); torch.cuda.synchronize()
print("wall (events) per sampling call:", e0.elapsed_time(e1) / 5, "ms")
