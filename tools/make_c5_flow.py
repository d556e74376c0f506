"""Train the c2 flow that bench.py's configs[4] leg (posterior samples/s, log_prob/s) loads at every GPU count, and store it
as a compact fixture (fp16 weights + fp32 z-score statistics): tests/golden/c5_flow_c2.npz.

The training itself is not bit-reproducible (the weight-gradient kernels add with atomics), so the acceptance rate of a
flow trained inside the bench differed from run to run (0.34 .. 0.85): a COMMITTED flow makes accepted samples/s
comparable across GPU counts and rounds.  Run on a GPU box:  python tools/make_c5_flow.py gpurun_out/c5_flow_c2.npz"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import floppity_deprecated_b200 as fb
from floppity_deprecated_b200 import synthetic
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "c5_flow_c2.npz")
sys.argv = [sys.argv[0]]
import bench
cfg, theta, x = bench.make_data("c2", 65536, 0)
prior = fb.BoxUniform(-3 * torch.ones(cfg.D), 3 * torch.ones(cfg.D), device="cuda")
inf = fb.SNPE_C(prior, fb.posterior_nn("nsf", hidden_features=cfg.hidden, num_transforms=cfg.transforms, num_bins=cfg.bins,
                                        num_blocks=cfg.blocks), device="cuda")
inf.append_simulations(theta, x)
net = inf.train(training_batch_size=2048, learning_rate=1e-3, stop_after_epochs=5, max_num_epochs=60, seed=0)
print("epochs", inf._summary["epochs_trained"], "best validation log-prob", inf._summary["best_validation_log_prob"])
net.eval()
x_obs = bench.make_data("c2", 4096, 0)[2][:1]
post = fb.DirectPosterior(net, prior, max_sampling_batch_size=1 << 22).set_default_x(x_obs)
s = post.sample((1 << 22,))
acc = post.leakage_correction(x_obs, num_rejection_samples=1 << 20, force_update=True)
print("acceptance", acc, "sample std", s.std(0).mean().item())
np.savez_compressed(out, flat=net._flat.detach().cpu().numpy().astype(np.float16),
                    theta_shift=net._theta_shift.cpu().numpy(), theta_scale=net._theta_scale.cpu().numpy(),
                    x_mean=net._x_mean.cpu().numpy(), x_std=net._x_std.cpu().numpy(), x_obs=x_obs.numpy(),
                    acceptance=np.float64(acc))
print("wrote", out, os.path.getsize(out) / 1e6, "MB")
