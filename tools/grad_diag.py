"""Diagnostic: gradient error of the CUDA path vs the fp32 oracle vs the fp64 oracle on a golden case (per tensor)."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import golden_oracle_net
import floppity_deprecated_b200 as fb
from floppity_deprecated_b200 import _lib
from oracle import nsf_oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
onet, g, cfg = golden_oracle_net(name)
D, C, T, K, H, Bk = cfg[:6]
theta, x = torch.tensor(g["theta"]), torch.tensor(g["x"])
onet.train(); onet.zero_grad()
(-onet.log_prob(theta, x).mean()).backward()
ref32 = {k: p.grad.numpy().copy() for k, p in onet.named_parameters()}
o64 = O.build_nsf(torch.rand(64, D).double(), torch.randn(64, C).double(), hidden_features=H, num_transforms=T, num_bins=K, num_blocks=Bk).double()
o64.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in onet.state_dict().items()})
o64.train()
(-o64.log_prob(theta.double(), x.double()).mean()).backward()
ref64 = {k: p.grad.numpy().copy() for k, p in o64.named_parameters()}
for mode in [(1, 0, 0, 0)]:
    _lib.lib().fp_set_mode(*mode)
    net = fb.NSFFlow(D, C, hidden_features=H, num_transforms=T, num_bins=K, num_blocks=Bk)
    net.load_state_dict(onet.state_dict()); net = net.cuda().train()
    (-net.log_prob(theta.cuda().requires_grad_(), x.cuda()).mean()).backward()
    got = {k: v.cpu().numpy() for k, v in net.unpack(net._flat.grad).items()}
    worst = []
    for k in ref32:
        m = max(np.abs(ref64[k]).max(), 1e-12)
        worst.append((np.abs(got[k] - ref64[k]).max() / m, np.abs(ref32[k] - ref64[k]).max() / m, np.abs(got[k] - ref32[k]).max() / m, k))
    worst.sort(reverse=True)
    print("mode", mode, "worst (cuda-vs-fp64, fp32oracle-vs-fp64, cuda-vs-fp32oracle) relative to the tensor max:")
    for w in worst[:6]:
        print("   %.2e  %.2e  %.2e  %s" % w)
    k = worst[0][3]
    d = np.abs(got[k] - ref64[k])
    idx = np.argsort(d.ravel())[::-1][:8]
    print("   worst tensor", k, got[k].shape, "worst elements:", [(np.unravel_index(i, d.shape), float(got[k].ravel()[i]), float(ref64[k].ravel()[i])) for i in idx])
    kb = k.replace("weight", "bias")
    d = np.abs(got[kb] - ref64[kb]); idx = np.argsort(d)[::-1][:8]
    print("   bias worst:", [(int(i), float(got[kb][i]), float(ref64[kb][i])) for i in idx])
    print("   median cuda-vs-fp64 %.2e, median fp32oracle-vs-fp64 %.2e" % (np.median([w[0] for w in worst]), np.median([w[1] for w in worst])))
