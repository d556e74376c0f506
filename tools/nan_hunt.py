"""Diagnostic: train the c2 flow as bench.py's c5 leg does, draw until non-finite samples appear, then evaluate the CPU
oracle on the same base noise to see whether the restated sbi path agrees."""
import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import floppity_deprecated_b200 as fb
from floppity_deprecated_b200 import _lib, synthetic
from floppity_deprecated_b200.inference import _TrainEngine
from oracle import nsf_oracle as O
sys.argv = [sys.argv[0]]
import bench
cfg, theta, x = bench.make_data("c2", 65536, 0)
dev = torch.device("cuda")
torch.manual_seed(0)
net = fb.build_nsf(theta[:8192], x[:8192], hidden_features=cfg.hidden, num_transforms=cfg.transforms, num_bins=cfg.bins,
                   num_blocks=cfg.blocks, generator=torch.Generator().manual_seed(0)).to(dev).train()
prior = fb.BoxUniform(-3 * torch.ones(cfg.D), 3 * torch.ones(cfg.D), device=dev)
eng = _TrainEngine(net, prior, 4096, False, cfg.atoms, True, float(os.environ.get("LR", "2e-3")), 5.0)
th_d, x_d = theta.to(dev), net._standardize(x.to(dev)); m_d = torch.zeros(65536, device=dev)
gp = torch.Generator(device=dev).manual_seed(99)
for i in range(400):
    eng.step(torch.randint(0, 65536, (4096,), device=dev, generator=gp), th_d, x_d, m_d)
eng.check_status()
print("final train loss/row", float(eng.loss_sum.item()) / 4096 / 400)
net.eval()
ctx = net._standardize(bench.make_data("c2", 4096, 0)[2][:1].to(dev))
bad_z = []
for rep in range(8):
    z = torch.randn(1 << 21, cfg.D, device=dev)
    st = torch.zeros(4, dtype=torch.int32, device=dev)
    s = net.inverse_rows(z, ctx, st)
    nf = ~torch.isfinite(s).all(dim=1)
    print("rep", rep, "non-finite rows", int(nf.sum()), "status", st.tolist(), "max |z| of bad rows", float(z[nf].abs().max()) if nf.any() else 0.0)
    if nf.any():
        bad_z.append(z[nf][:8].cpu()); 
        if len(bad_z) >= 2: break
if bad_z:
    zb = torch.cat(bad_z)
    onet = O.build_nsf(theta[:8192], x[:8192], hidden_features=cfg.hidden, num_transforms=cfg.transforms, num_bins=cfg.bins, num_blocks=cfg.blocks)
    onet.load_state_dict({k: v.cpu() for k, v in net.state_dict().items()}, strict=False)
    onet.eval()
    xo = bench.make_data("c2", 4096, 0)[2][:1]
    with torch.no_grad():
        so = onet.sample(zb.shape[0], context=xo, noise=zb)[0]
        s64 = onet.double().sample(zb.shape[0], context=xo.double(), noise=zb.double())[0]
    ours = net.inverse_rows(zb.to(dev), ctx).cpu()
    print("oracle fp32 finite rows:", torch.isfinite(so).all(dim=1).tolist())
    print("oracle fp64 finite rows:", torch.isfinite(s64).all(dim=1).tolist())
    print("ours finite rows:", torch.isfinite(ours).all(dim=1).tolist())
    print("oracle fp64 |sample| max per row:", s64.abs().amax(dim=1).tolist())
    print("bad z rows |z| max:", zb.abs().amax(dim=1).tolist())
