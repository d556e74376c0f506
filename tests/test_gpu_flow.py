"""GPU parity, part 2: the whole flow through the reference-facing Python surface (build_nsf / log_prob / sample /
SNPE_C.train / DirectPosterior) against the oracle and the committed golden vectors.  Tolerance: 1e-4 relative
(|a-b| / max(1,|b|)) for log_prob and transformed samples (BASELINE.json north_star), fp32 throughout."""
import copy
import pickle

import numpy as np
import pytest
import torch

import floppity_deprecated_b200 as fb
from conftest import GRAD_TOL, assert_grads_close, golden_oracle_net, knife_edge_units, rel_err
from floppity_deprecated_b200 import synthetic
from oracle import nsf_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _twin(onet, cfg):
    """CUDA flow carrying exactly the oracle network's weights (through the upstream-keyed state_dict)."""
    D, C, T, K, H, Bk = cfg[:6]
    net = fb.NSFFlow(D, C, hidden_features=H, num_transforms=T, num_bins=K, num_blocks=Bk)
    net.load_state_dict(onet.state_dict())
    return net.cuda().eval()


@pytest.mark.parametrize("name", ["c1", "odd", "c2", "c3", "c4"])
def test_log_prob_matches_golden_and_oracle(name):
    onet, g, cfg = golden_oracle_net(name)
    net = _twin(onet, cfg)
    theta, x = torch.tensor(g["theta"]), torch.tensor(g["x"])
    lp = net.log_prob(theta.cuda(), x.cuda()).detach().cpu().numpy()
    assert rel_err(lp, g["logp"]).max() < TOL, rel_err(lp, g["logp"]).max()
    # and not further from the fp64 truth than the fp32 oracle is (plus the bar)
    assert rel_err(lp, g["logp64"]).max() < TOL + rel_err(g["logp"], g["logp64"]).max()
    # single broadcast context row == repeated rows (context hoisting must not change results)
    with torch.no_grad():
        lp1 = net.log_prob(theta.cuda(), x[:1].cuda()).cpu().numpy()
    with torch.no_grad():
        want1 = onet.log_prob(theta, x[:1].expand(theta.shape[0], -1)).numpy()
    assert rel_err(lp1, want1).max() < TOL


@pytest.mark.parametrize("name", ["c1", "odd", "c2", "c3", "c4"])
def test_samples_match_golden_for_fixed_noise(name):
    onet, g, cfg = golden_oracle_net(name)
    net = _twin(onet, cfg)
    x, z = torch.tensor(g["x"]), torch.tensor(g["z"])
    B = z.shape[0]
    s1 = net.sample(B, context=x[:1].cuda(), noise=z.cuda())[0].cpu().numpy()
    assert rel_err(s1, g["samples_one_x"]).max() < TOL, rel_err(s1, g["samples_one_x"]).max()
    s2 = net.sample(1, context=x.cuda(), noise=z.cuda())[:, 0].cpu().numpy()
    assert rel_err(s2, g["samples_per_row"]).max() < TOL
    # encode -> decode round trip on the device: log_prob's z of the samples is the noise
    with torch.no_grad():
        lp = net.log_prob(torch.tensor(s1).cuda(), x[:1].cuda())
        want = onet.log_prob(torch.tensor(s1), x[:1].expand(B, -1))
    assert rel_err(lp.cpu().numpy(), want.numpy()).max() < TOL


@pytest.mark.parametrize("name", ["c1", "odd", "c2", "c3", "c4"])
def test_gradients_match_golden_and_oracle(name):
    onet, g, cfg = golden_oracle_net(name)
    net = _twin(onet, cfg)
    net.train()
    theta = torch.tensor(g["theta"]).cuda().requires_grad_()
    x = torch.tensor(g["x"]).cuda()
    loss = -net.log_prob(theta, x).mean()
    loss.backward()
    grads = {k: v.cpu().numpy() for k, v in net.unpack(net._flat.grad).items()}
    keys = [str(k) for k in g["grad_keys"]]
    assert sorted(grads) == keys
    # full tensor-by-tensor check against the oracle's autograd: 2e-4 of each tensor's maximum (GRAD_TOL), rows of relu
    # knife-edge units excepted (conftest.assert_grads_close)
    onet.zero_grad()
    (-onet.log_prob(torch.tensor(g["theta"]), torch.tensor(g["x"])).mean()).backward()
    knife = knife_edge_units(onet, torch.tensor(g["theta"]), torch.tensor(g["x"]))
    ref = {k: p.grad.numpy() for k, p in onet.named_parameters()}
    worst = assert_grads_close(grads, ref, knife)
    print(f"gradients {name}: worst {worst:.2e} of the tensor maximum")
    # the committed vectors: per-tensor norms, one whole tensor, and the input gradient (rows with a knife-edge unit excepted)
    flipped = worst > GRAD_TOL or any(np.abs(grads[k] - ref[k]).max() > GRAD_TOL * max(np.abs(ref[k]).max(), 1e-4) for k in ref)
    tol_n = 10 * GRAD_TOL if flipped else GRAD_TOL
    norms = np.array([np.linalg.norm(grads[k]) for k in keys])
    bad = [(k, a, b) for k, a, b in zip(keys, norms, g["grad_norms"]) if abs(a - b) > 5 * tol_n * max(b, 1e-3)]
    assert not bad, bad[:5]
    kb = str(g["grad_final_bias_key"])
    assert np.abs(grads[kb] - g["grad_final_bias"]).max() < tol_n * np.abs(g["grad_final_bias"]).max()
    ok_rows = np.array([r not in knife[1] for r in range(theta.shape[0])])
    dth = theta.grad.cpu().numpy()
    assert np.abs(dth[ok_rows] - g["d_theta"][ok_rows]).max() < tol_n * np.abs(g["d_theta"]).max()
    assert np.abs(dth - g["d_theta"]).max() < 2e-3 * np.abs(g["d_theta"]).max()


def test_atomic_loss_path_matches_golden():
    onet, g, cfg = golden_oracle_net("c1")
    net = _twin(onet, cfg)
    D = cfg[0]
    theta = torch.tensor(g["theta"]).clamp(-3.2, 3.2)
    x, choices, masks = torch.tensor(g["x"]), torch.tensor(g["atom_choices"]), torch.tensor(g["atom_masks"])
    B, A = theta.shape[0], choices.shape[1] + 1
    rows = torch.cat([theta[:, None], theta[choices]], dim=1).reshape(B * A, D)
    with torch.no_grad():
        lp = net.log_prob(rows.cuda(), x.cuda())          # R = B*A rows share B context rows (no repeat)
    from floppity_deprecated_b200 import _lib
    out = torch.empty(B, device="cuda")
    low, high = (-3.3 * torch.ones(D)).cuda(), (3.3 * torch.ones(D)).cuda()
    rows_d, masks_d = rows.cuda(), masks.reshape(-1).cuda()
    _lib.check(_lib.lib().fp_atomic_loss(_lib.ptr(lp), _lib.ptr(rows_d), _lib.ptr(low), _lib.ptr(high),
                                         _lib.ptr(masks_d), 1, _lib.ptr(out), None, B, A, D, None, _lib.stream()), "loss")
    assert rel_err(out.cpu().numpy(), g["atomic_lp"]).max() < TOL


def test_training_steps_track_the_oracle():
    """Three optimiser steps (MLE and atomic loss) on identical batches from identical weights: loss value and every
    parameter gradient match torch autograd; the clipped Adam update matches wherever the gradient is above noise
    (Adam's early updates are +-lr whatever the gradient size, so noise-level gradients cannot be compared)."""
    torch.manual_seed(0)
    D, C, B = 6, 12, 64
    theta, x = torch.rand(256, D) * 6 - 3, torch.randn(256, C)
    onet = O.build_nsf(theta, x, hidden_features=24, num_transforms=3, num_bins=6, num_blocks=2)
    O.randomize_lu(onet, torch.Generator().manual_seed(1), 0.1)
    prior_o = O.BoxUniform(-3 * torch.ones(D), 3 * torch.ones(D))
    prior = fb.BoxUniform(-3 * torch.ones(D), 3 * torch.ones(D), device="cuda")
    from floppity_deprecated_b200.inference import _TrainEngine
    for atomic in (False, True):
        o = copy.deepcopy(onet)
        net = fb.NSFFlow(D, C, hidden_features=24, num_transforms=3, num_bins=6, num_blocks=2).cuda().train()
        net.load_state_dict(o.state_dict())                         # z-score buffers must be in place before the engine caches them
        eng = _TrainEngine(net, prior, B, atomic, 5, True, 5e-4, 5.0)
        opt = torch.optim.Adam(o.parameters(), lr=5e-4)
        o.train()
        masks = torch.ones(B, 1)
        for s in range(3):
            tb, xb = theta[s * B:(s + 1) * B], x[s * B:(s + 1) * B]
            net.load_state_dict(o.state_dict())                     # identical weights going into the step
            eng.loss_sum.zero_()
            before = net._flat.detach().clone()
            eng.step_on_batch(tb.cuda(), net._standardize(xb.cuda()), masks.cuda())
            opt.zero_grad()
            if atomic:
                ch = eng.choices.cpu().long()
                lp = O.log_prob_proposal_posterior_atomic(o, prior_o, tb, xb, masks, 5, True, ch)
            else:
                lp = o.log_prob(tb, xb)
            loss = -lp.mean()
            loss.backward()
            assert abs(eng.loss_sum.item() / B - loss.item()) < 1e-4 * max(1.0, abs(loss.item())), (atomic, s)
            got = {k: v.cpu().numpy() for k, v in net.unpack(eng.grads[: net.param_count]).items()}
            rows = torch.cat([tb[:, None], tb[ch]], dim=1).reshape(-1, D) if atomic else tb
            knife = knife_edge_units(o, rows, xb.repeat_interleave(rows.shape[0] // B, dim=0))
            assert_grads_close(got, {k: p.grad.numpy() for k, p in o.named_parameters()}, knife)
            if s == 0:   # first Adam step from zero moments: update = -lr * clip*g / (|clip*g| + eps)
                torch.nn.utils.clip_grad_norm_(o.parameters(), 5.0)
                opt.step()
                upd = net.unpack(net._flat.detach() - before)
                new_o = dict(o.named_parameters())
                old_o = net.unpack(before)
                for k, p in new_o.items():
                    big = p.grad.abs() > 1e-5
                    if big.any():
                        d_ref = (p.detach() - old_o[k].cpu())[big]
                        assert (upd[k].cpu()[big] - d_ref).abs().max() < 2e-6, (atomic, k)
        eng.check_status()


def test_snpe_c_end_to_end_multi_round():
    """The reference's round loop (floppity.py:168-285) on the synthetic c1 task: MLE round, atomic round, posterior."""
    torch.manual_seed(0)
    theta, x, scaler, _ = synthetic.make_dataset("c1", 2048)
    prior = fb.utils.BoxUniform(low=-3 * torch.ones(10), high=3 * torch.ones(10), device="cuda")
    inference = fb.SNPE_C(prior=prior, density_estimator=fb.posterior_nn(model="nsf", num_transforms=5, hidden_features=50,
                                                                         num_bins=8, num_blocks=2, dropout_probability=0.0), device="cuda")
    inference.append_simulations(torch.tensor(theta), torch.tensor(x), proposal=None)
    est = inference.train(show_train_summary=False, stop_after_epochs=2, max_num_epochs=6, num_atoms=10,
                          force_first_round_loss=True, use_combined_loss=True, training_batch_size=256)
    v = inference._summary["validation_log_probs"]
    assert len(v) >= 2 and v[-1] > v[0] and np.isfinite(v).all()
    xo, _ = synthetic.observation("c1", scaler)
    posterior = inference.build_posterior(est).set_default_x(xo)
    s = posterior.sample((5000,))
    assert s.shape == (5000, 10) and s.is_cuda and prior.within(s).all()
    lp = posterior.log_prob(s[:100], norm_posterior=False)
    assert torch.isfinite(lp).all()
    assert posterior.log_prob(torch.full((1, 10), 9.0)).item() == -float("inf")
    # round 2 with the atomic loss actually exercised (force_first_round_loss=False)
    th2 = posterior.sample((1024,)).cpu()
    x2 = synthetic.make_dataset("c1", theta=th2.numpy(), scaler=scaler)[1]
    inference.append_simulations(th2, torch.tensor(x2), proposal=posterior)
    est2 = inference.train(stop_after_epochs=1, max_num_epochs=2, num_atoms=10, force_first_round_loss=False,
                           use_combined_loss=True, training_batch_size=256)
    assert np.isfinite(inference._summary["validation_log_probs"]).all()
    # checkpoint objects survive torch.save / pickle (floppity.py:244-248)
    post2 = pickle.loads(pickle.dumps(inference.build_posterior(est2).set_default_x(xo)))
    assert post2.sample((10,)).shape == (10, 10)
    assert pickle.loads(pickle.dumps(inference))._data_round_index == [0, 1]


def test_density_estimator_adapter_shapes():
    torch.manual_seed(0)
    net = fb.build_nsf(torch.randn(128, 4), torch.randn(128, 9), hidden_features=16, num_transforms=2, num_bins=4).cuda().eval()
    de = fb.NSFDensityEstimator(net)
    inp, cond = torch.randn(5, 3, 4).cuda(), torch.randn(3, 9).cuda()
    lp = de.log_prob(inp, cond)
    assert lp.shape == (5, 3)
    want = torch.stack([net.log_prob(inp[s], cond) for s in range(5)])
    assert (lp - want).abs().max() < 1e-4
    assert de.sample((7,), cond).shape == (7, 3, 4) and de.loss(inp[0], cond).shape == (3,)


def test_map_moves_uphill():
    torch.manual_seed(0)
    theta, x, scaler, _ = synthetic.make_dataset("c1", 1024)
    net = fb.build_nsf(torch.tensor(theta), torch.tensor(x), hidden_features=32, num_transforms=3, num_bins=6).cuda().eval()
    prior = fb.BoxUniform(-3 * torch.ones(10), 3 * torch.ones(10), device="cuda")
    post = fb.DirectPosterior(net, prior).set_default_x(x[:1])
    init = post.sample((200,))
    m = post.map(num_iter=20, num_to_optimize=20, init_method=init)      # same starting population as the comparison
    assert post.log_prob(m[None], norm_posterior=False).item() >= post.log_prob(init, norm_posterior=False).max().item() - 1e-3


# ---------------------------------------------------------------------------------------------------------------------
# The fused conditioner kernel (cond_fused.cu, hidden 128) on MANY row tiles: ragged last tile, CTAs walking several
# tiles, every context-sharing pattern (one row per flow row, A atoms per context row, one context row for all).
# ---------------------------------------------------------------------------------------------------------------------
def _c2_pair(seed=11, name="c2", C=None, T=None):
    cfg = synthetic.CONFIGS[name]
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    theta = torch.rand(2048, cfg.D, generator=g) * 6 - 3
    x = torch.randn(2048, C or cfg.C, generator=g)
    hp = dict(hidden_features=cfg.hidden, num_transforms=T or cfg.transforms, num_bins=cfg.bins, num_blocks=cfg.blocks)
    onet = O.build_nsf(theta, x, **hp)
    O.randomize_lu(onet, g, 0.05)
    with torch.no_grad():
        for c in O.coupling_layers(onet):
            for blk in c.transform_net.blocks:
                blk.linear_layers[1].weight.add_(0.03 * torch.randn(blk.linear_layers[1].weight.shape, generator=g))
    net = fb.NSFFlow(cfg.D, C or cfg.C, **hp)
    net.load_state_dict(onet.state_dict())
    return onet.eval(), net.cuda().eval(), theta, x


@pytest.mark.parametrize("R,A", [(1000, 1), (1290, 10), (129, 1), (2048, 8)])
def test_fused_conditioner_many_tiles_log_prob(R, A):
    onet, net, theta, x = _c2_pair()
    Rc = R // A
    th, xc = theta[:R], x[:Rc]
    with torch.no_grad():
        got = net.log_prob(th.cuda(), xc.cuda()).cpu().numpy()
        want = onet.log_prob(th, xc.repeat_interleave(A, dim=0)).numpy()
    assert rel_err(got, want).max() < TOL, rel_err(got, want).max()


def test_fused_conditioner_many_tiles_sampling_and_round_trip():
    onet, net, theta, x = _c2_pair()
    n = 1500                                                     # 12 tiles, ragged
    z = torch.randn(n, 20, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        s = net.sample(n, context=x[:1].cuda(), noise=z.cuda())[0]
        want = onet.sample(n, context=x[:1], noise=z)[0].numpy()
        assert rel_err(s.cpu().numpy(), want).max() < TOL
        # size-independent property at a BASELINE-scale row count: decode -> encode returns the base noise
        n_big = 300_007
        zb = torch.randn(n_big, 20, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
        sb = net.inverse_rows(zb, net._standardize(x[:1].cuda()))
        lp = net.log_prob(sb, x[:1].cuda())
        # log q(theta) = log N(z) - log|det d theta / d z|: recover |z|^2 through the density of the round trip
        lp2 = net.log_prob(sb[:4096], x[:1].cuda())
        assert torch.equal(lp[:4096], lp2)                       # deterministic, independent of the batch it rode in
        assert torch.isfinite(lp).all()


def test_fused_conditioner_training_gradients_many_tiles():
    from floppity_deprecated_b200.inference import _TrainEngine
    onet, net, theta, x = _c2_pair()
    net.train(); onet.train()
    B = 640                                                      # 5 tiles of flow rows (MLE loss: one atom)
    prior = fb.BoxUniform(-3 * torch.ones(20), 3 * torch.ones(20), device="cuda")
    eng = _TrainEngine(net, prior, B, False, 10, True, 5e-4, 5.0)
    eng.step_on_batch(theta[:B].cuda(), net._standardize(x[:B].cuda()))
    loss = -onet.log_prob(theta[:B], x[:B]).mean()
    loss.backward()
    assert abs(eng.loss_sum.item() / B - loss.item()) < 1e-4 * max(1.0, abs(loss.item()))
    got = {k: v.cpu().numpy() for k, v in net.unpack(eng.grads[: net.param_count]).items()}
    assert_grads_close(got, {k: p.grad.numpy() for k, p in onet.named_parameters()}, knife_edge_units(onet, theta[:B], x[:B]))


def test_cpu_device_request_gets_cpu_results():
    """FlopPITy's default is -device cpu (floppity.py:30) and its helpers call .numpy() on log_prob's result
    (floppityFUN.py:37-39): a posterior built for 'cpu' computes on the GPU but hands results back on the CPU."""
    import warnings
    theta, x, _, _ = synthetic.make_dataset("c1", 1024)
    prior = fb.BoxUniform(-3 * torch.ones(10), 3 * torch.ones(10), device="cpu")
    inf = fb.SNPE_C(prior, density_estimator=fb.posterior_nn(model="nsf", hidden_features=50, num_transforms=3, num_bins=8),
                    device="cpu")
    inf.append_simulations(torch.tensor(theta), torch.tensor(x))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        est = inf.train(training_batch_size=256, max_num_epochs=2, stop_after_epochs=2, seed=0)
    post = inf.build_posterior(est).set_default_x(torch.tensor(x[:1]))
    s = post.sample((100,))
    assert not s.is_cuda and s.shape == (100, 10)
    P = post.log_prob(torch.tensor(theta[:50]), x=x[:1])
    pi = prior.log_prob(torch.tensor(theta[:50]))
    _ = -(P - pi).detach().numpy()                          # the expression of floppityFUN.py:37
    m = post.map(num_iter=5, num_to_optimize=10, num_init_samples=50)
    assert not m.is_cuda and m.shape == (10,)
    from floppity_deprecated_b200 import retrieval as RT
    spec = np.abs(x[:50].astype(np.float64)) + 1.0
    w, eff = RT.IS(post, spec[0], np.full(spec.shape[1], 0.1), spec, theta[:50])
    assert w.shape == (50,) and 0.0 < eff <= 1.0


def test_example_retrieval_round_loop(tmp_path):
    """examples/synthetic_retrieval.py = floppity.py's round loop on the synthetic forward model: preprocess on the GPU,
    MLE round then (force_first_round_loss=0) atomic rounds, IS evidence, pickled checkpoints, MAP."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("synthetic_retrieval", os.path.join(os.path.dirname(__file__), "..", "examples", "synthetic_retrieval.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = mod.main(["--config", "c1", "--rounds", "2", "--samples", "1024", "--naug", "2", "--patience", "2", "--batch", "256",
                    "--npew", "500", "--force-first-round-loss", "0", "--out", str(tmp_path)])
    assert len(out["samples"]) == 2 and out["samples"][0].shape == (500, 10)
    assert np.isfinite(out["MAP"]).all() and len(out["logZs"]) == 1 and np.isfinite(out["logZs"][0][0])
    for f in ("posteriors.pt", "inference.p", "samples.p", "xscaler.p", "pca.p", "map_params.npy"):
        assert os.path.exists(tmp_path / f), f
    posts = torch.load(tmp_path / "posteriors.pt", weights_only=False)          # modules.py:74 (-resume)
    assert posts[-1].sample((10,)).shape == (10, 10)


def test_fused_conditioner_forty_parameters():
    """c4 shapes on the fused kernel: 40 parameters -> the initial layer's contraction spans two k-blocks (D <= 64)."""
    onet, net, theta, x = _c2_pair(13, "c4", C=96, T=4)
    with torch.no_grad():
        got = net.log_prob(theta[:700].cuda(), x[:70].cuda()).cpu().numpy()           # 10 atoms per context row
        want = onet.log_prob(theta[:700], x[:70].repeat_interleave(10, dim=0)).numpy()
    assert rel_err(got, want).max() < TOL, rel_err(got, want).max()
    z = torch.randn(300, 40, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        s = net.sample(300, context=x[:1].cuda(), noise=z.cuda())[0].cpu().numpy()
        ws = onet.sample(300, context=x[:1], noise=z)[0].numpy()
    assert rel_err(s, ws).max() < TOL


@pytest.mark.parametrize("D,Bk,K,A", [(22, 3, 10, 1), (8, 1, 8, 4), (64, 4, 5, 1), (36, 2, 16, 7)])
def test_fused_conditioner_other_shapes(D, Bk, K, A):
    """Block counts beyond the two whose biases sit in shared memory, a single block, the widest supported flow (64
    parameters), odd numbers of spline parameters per row (TMA stores clip at d_tr * (3K-1)), atoms that do not divide 128."""
    g = torch.Generator().manual_seed(100 + D)
    torch.manual_seed(100 + D)
    C, R = 24, 7 * 128 * A // A + 3 * A                      # several tiles, ragged
    Rc = R // A
    R = Rc * A
    theta = torch.rand(1024, D, generator=g) * 6 - 3
    x = torch.randn(1024, C, generator=g)
    hp = dict(hidden_features=128, num_transforms=2, num_bins=K, num_blocks=Bk)
    onet = O.build_nsf(theta, x, **hp)
    O.randomize_lu(onet, g, 0.05)
    with torch.no_grad():
        for c in O.coupling_layers(onet):
            for blk in c.transform_net.blocks:
                blk.linear_layers[1].weight.add_(0.03 * torch.randn(blk.linear_layers[1].weight.shape, generator=g))
    net = fb.NSFFlow(D, C, **hp)
    net.load_state_dict(onet.state_dict())
    net, onet = net.cuda().eval(), onet.eval()
    th = theta[:R] if R <= 1024 else theta.repeat(2, 1)[:R]
    with torch.no_grad():
        got = net.log_prob(th.cuda(), x[:Rc].cuda()).cpu().numpy()
        want = onet.log_prob(th, x[:Rc].repeat_interleave(A, dim=0)).numpy()
    assert rel_err(got, want).max() < TOL, rel_err(got, want).max()
    # training-mode forward + backward through the same kernel (saved activations via TMA stores)
    th_g = th[:256].cuda().requires_grad_(True)
    net.train()
    lp = net.log_prob(th_g, x[:256].cuda())
    lp.sum().backward()
    tho = th[:256].clone().requires_grad_(True)
    onet.train()
    onet.log_prob(tho, x[:256]).sum().backward()
    assert rel_err(th_g.grad.cpu().numpy(), tho.grad.numpy()).max() < 2e-3


def test_wide_conditioner_gemm_chain_hidden_512():
    """c3-style conditioner (hidden 512, 4 blocks, 16 bins): the per-layer tcgen05 GEMM chain (the fused kernel covers
    hidden 128 only) -- log_prob, samples and a training step's gradients against the oracle."""
    from floppity_deprecated_b200.inference import _TrainEngine
    g = torch.Generator().manual_seed(77)
    torch.manual_seed(77)
    D, C = 30, 96
    theta = torch.rand(640, D, generator=g) * 6 - 3
    x = torch.randn(640, C, generator=g)
    hp = dict(hidden_features=512, num_transforms=3, num_bins=16, num_blocks=4)
    onet = O.build_nsf(theta, x, **hp)
    O.randomize_lu(onet, g, 0.04)
    net = fb.NSFFlow(D, C, **hp)
    net.load_state_dict(onet.state_dict())
    net, onet = net.cuda().eval(), onet.eval()
    with torch.no_grad():
        got = net.log_prob(theta[:300].cuda(), x[:300].cuda()).cpu().numpy()
        want = onet.log_prob(theta[:300], x[:300]).numpy()
    assert rel_err(got, want).max() < TOL, rel_err(got, want).max()
    z = torch.randn(200, D, generator=g)
    with torch.no_grad():
        s = net.sample(200, context=x[:1].cuda(), noise=z.cuda())[0].cpu().numpy()
        ws = onet.sample(200, context=x[:1], noise=z)[0].numpy()
    assert rel_err(s, ws).max() < TOL
    net.train(); onet.train()
    B = 256
    prior = fb.BoxUniform(-3 * torch.ones(D), 3 * torch.ones(D), device="cuda")
    eng = _TrainEngine(net, prior, B, False, 10, True, 5e-4, 5.0)
    eng.step_on_batch(theta[:B].cuda(), net._standardize(x[:B].cuda()))
    (-onet.log_prob(theta[:B], x[:B]).mean()).backward()
    got_g = {k: v.cpu().numpy() for k, v in net.unpack(eng.grads[: net.param_count]).items()}
    assert_grads_close(got_g, {k: p.grad.numpy() for k, p in onet.named_parameters()}, knife_edge_units(onet, theta[:B], x[:B]))


def test_fp16_operand_range_guard():
    """The fused conditioner's 3xFP16 operands saturate beyond +-65504: a flow whose activations get there is reported
    through the status word (FloatingPointError naming the FLOPPITY_FUSED_FP16=0 switch), never silently wrong."""
    import os
    from floppity_deprecated_b200 import _lib
    from floppity_deprecated_b200.inference import _TrainEngine
    if os.environ.get("FLOPPITY_FUSED_FP16", "1") == "0" or os.environ.get("FLOPPITY_FUSED", "1") == "0":
        pytest.skip("3xFP16 fused conditioner switched off")
    torch.manual_seed(0)
    D, C = 8, 16
    theta, x = torch.rand(512, D) * 6 - 3, torch.randn(512, C)
    net = fb.build_nsf(theta, x, hidden_features=128, num_transforms=2, num_bins=6, num_blocks=1).cuda()
    prior = fb.BoxUniform(-3 * torch.ones(D), 3 * torch.ones(D), device="cuda")
    post = fb.DirectPosterior(net, prior).set_default_x(x[:1])
    assert post.sample((64,)).shape == (64, D)                           # a sane flow: no complaint
    with torch.no_grad():
        net.view(0, _lib.P_CTX_B, 0).fill_(3.0e5)                        # h0 = 3e5 everywhere: outside fp16
    with pytest.raises(FloatingPointError, match="FLOPPITY_FUSED_FP16=0"):
        post.sample((64,))
    eng = _TrainEngine(net.train(), prior, 128, False, 10, True, 5e-4, 5.0)
    eng.step_on_batch(theta[:128].cuda(), net._standardize(x[:128].cuda()))
    with pytest.raises(FloatingPointError):
        eng.check_status()
