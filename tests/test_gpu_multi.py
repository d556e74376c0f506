"""Data-parallel correctness on real GPUs: a W-rank optimiser step over NCCL equals the 1-rank step on the
concatenated batch -- gradients (after the all-reduce, scaled 1/W) and the weights after clip + Adam.  Needs >= 2 GPUs
(skipped otherwise; `gpurun --gpus 2`).  Also the sharded sampler / log_prob: every rank draws its share with no
collective on the data path, and the gathered result has the right size and support."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")]

_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
out, D, C, H, B = sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import floppity_deprecated_b200 as fb
from floppity_deprecated_b200 import dist as fdist
from floppity_deprecated_b200.inference import _TrainEngine
g = torch.Generator().manual_seed(0)
theta = torch.rand(4096, D, generator=g) * 6 - 3
x = torch.randn(4096, C, generator=g)
net = fb.build_nsf(theta, x, hidden_features=H, num_transforms=3, num_bins=8, num_blocks=2,
                   generator=torch.Generator().manual_seed(1)).cuda().train()
fdist.broadcast_(net._flat.data)
prior = fb.BoxUniform(-3 * torch.ones(D), 3 * torch.ones(D), device="cuda")
eng = _TrainEngine(net, prior, B, False, 10, True, 5e-4, 5.0)
assert eng.world == world
for step in range(3):
    rows = torch.arange(step * world * B, (step + 1) * world * B)[rank::world]        # this rank's share of the global batch
    eng.step_on_batch(theta[rows].cuda(), net._standardize(x[rows].cuda()))
eng.check_status()
torch.cuda.synchronize()
if rank == 0:
    torch.save({"flat": net._flat.detach().cpu(), "grads": eng.grads[: net.param_count].cpu() / world}, out)
# sharded sampling / log_prob: no collective on the data path, optional gather at the end
post = fb.DirectPosterior(net.eval(), prior).set_default_x(x[:1])
s = post.sample_sharded(10001, gather=True)
assert s.shape == (10001, D) and bool(prior.within(s).all())
lo, hi = fdist.shard_rows(10001)
lp = post.log_prob_sharded(s, norm_posterior=False, gather=True)
assert lp.shape == (10001,) and torch.isfinite(lp).all()
full = post.log_prob(s, norm_posterior=False)
assert torch.allclose(lp.cpu(), full.cpu(), rtol=1e-5, atol=1e-5)
dist.destroy_process_group()
print("ok", rank)
'''


@pytest.mark.parametrize("H", [128, 50])
def test_two_rank_step_equals_one_rank_step_on_the_concatenated_batch(tmp_path, H):
    D, C, B, W = 8, 64, 512, 2
    out = str(tmp_path / "dp.pt")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for r in range(W):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(W), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT, out, str(D), str(C), str(H), str(B)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    got = torch.load(out)

    import floppity_deprecated_b200 as fb
    from floppity_deprecated_b200.inference import _TrainEngine
    g = torch.Generator().manual_seed(0)
    theta = torch.rand(4096, D, generator=g) * 6 - 3
    x = torch.randn(4096, C, generator=g)
    net = fb.build_nsf(theta, x, hidden_features=H, num_transforms=3, num_bins=8, num_blocks=2,
                       generator=torch.Generator().manual_seed(1)).cuda().train()
    prior = fb.BoxUniform(-3 * torch.ones(D), 3 * torch.ones(D), device="cuda")
    eng = _TrainEngine(net, prior, W * B, False, 10, True, 5e-4, 5.0)
    start = net._flat.detach().cpu().clone()
    for step in range(3):
        rows = torch.arange(step * W * B, (step + 1) * W * B)
        eng.step_on_batch(theta[rows].cuda(), net._standardize(x[rows].cuda()))
    grads = eng.grads[: net.param_count].cpu()
    scale = float(grads.abs().max())
    assert float((got["grads"] - grads).abs().max()) < 2e-5 * scale          # same sum, different order of additions
    travelled = float((net._flat.detach().cpu() - start).abs().max())         # ~3 x lr: Adam's first steps move every weight by lr
    apart = float((net._flat.detach().cpu() - got["flat"]).abs().max())
    assert travelled > 1e-3 and apart < 1e-2 * travelled, (apart, travelled)  # (Adam divides by sqrt(v): rounding-level
    #                                                                           gradient differences show up at 1e-3 of a step)
