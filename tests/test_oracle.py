"""CPU suite, part 1: the oracle itself -- analytic known answers and the committed golden vectors.
(The reference ships no tests; these are the pins SURVEY.md section 4 asks the build to create.)"""
import numpy as np
import pytest
import torch

from conftest import golden_oracle_net, load_golden, rel_err
from oracle import nsf_oracle as O


def _ident_params(n, K):
    return torch.zeros(n, K), torch.zeros(n, K), torch.full((n, K - 1), float(np.log(np.exp(1 - 1e-3) - 1)))


def test_identity_spline_known_answer():
    x = torch.linspace(-4, 4, 201, dtype=torch.float64)
    uw, uh, ud = (t.double() for t in _ident_params(201, 8))
    y, lad = O.unconstrained_rational_quadratic_spline(x, uw, uh, ud, tail_bound=3.0)
    assert (y - x).abs().max() < 1e-8 and lad.abs().max() < 1e-8


def test_linear_tails_are_identity():
    x = torch.tensor([-7.0, -3.0001, 3.0001, 12.0])
    y, lad = O.unconstrained_rational_quadratic_spline(x, torch.randn(4, 6), torch.randn(4, 6), torch.randn(4, 5), tail_bound=3.0)
    assert torch.equal(y, x) and torch.equal(lad, torch.zeros(4))


def test_boundary_inputs_are_inside():
    x = torch.tensor([-3.0, 3.0])
    _, _, bins, _ = O.unconstrained_rational_quadratic_spline(x, torch.randn(2, 6), torch.randn(2, 6), torch.randn(2, 5),
                                                              tail_bound=3.0, return_bins=True)
    assert bins.tolist() == [0, 5]


@pytest.mark.parametrize("K", [4, 8, 16])
def test_roundtrip_and_logdet(K):
    g = torch.Generator().manual_seed(K)
    n = 4000
    x = (torch.rand(n, generator=g, dtype=torch.float64) * 8 - 4).requires_grad_()
    uw, uh, ud = (torch.randn(n, K, generator=g, dtype=torch.float64), torch.randn(n, K, generator=g, dtype=torch.float64),
                  torch.randn(n, K - 1, generator=g, dtype=torch.float64))
    y, lad = O.unconstrained_rational_quadratic_spline(x, uw, uh, ud, tail_bound=3.0)
    (gx,) = torch.autograd.grad(y.sum(), x)
    assert (gx.log() - lad).abs().max() < 1e-10          # logdet == log dy/dx
    xi, ladi = O.unconstrained_rational_quadratic_spline(y.detach(), uw, uh, ud, inverse=True, tail_bound=3.0)
    assert (xi - x).abs().max() < 1e-9 and (ladi + lad).abs().max() < 1e-9
    assert (y[1:][x[1:] > x[:-1]] is not None)


def test_lu_linear_identity_at_init_and_inverse():
    lu = O.LULinear(6)
    x = torch.randn(5, 6)
    y, ld = lu(x)
    assert torch.allclose(y, x, atol=1e-6) and ld.abs().max() < 1e-6
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for p in lu.parameters():
            p.add_(0.3 * torch.randn(p.shape, generator=g))
    y, ld = lu(x)
    xi, ldi = lu.inverse(y)
    assert torch.allclose(xi, x, atol=1e-5) and torch.allclose(ld, -ldi)
    L, U = lu._create_lower_upper()
    assert torch.allclose(ld[0], torch.linalg.slogdet(L @ U)[1], atol=1e-5)


def test_alternating_mask_and_param_count():
    assert O.create_alternating_binary_mask(5, True).tolist() == [1, 0, 1, 0, 1]
    assert O.create_alternating_binary_mask(5, False).tolist() == [0, 1, 0, 1, 0]
    net = O.build_nsf(torch.randn(64, 10), torch.randn(64, 100), hidden_features=50, num_transforms=5, num_bins=8)
    assert sum(p.numel() for p in net.parameters()) == 157_875      # SURVEY.md 8a row A1 (c1: 0.158 M)


def test_density_normalises_in_2d():
    torch.manual_seed(3)
    net = O.build_nsf(torch.randn(256, 2), torch.randn(256, 3), hidden_features=16, num_transforms=3, num_bins=6).double()
    O.randomize_lu(net, torch.Generator().manual_seed(1), 0.2)
    net.eval()
    grid = torch.linspace(-9, 9, 361, dtype=torch.float64)
    xx, yy = torch.meshgrid(grid, grid, indexing="ij")
    pts = torch.stack([xx.reshape(-1), yy.reshape(-1)], dim=1)
    ctx = torch.randn(1, 3, dtype=torch.float64).expand(pts.shape[0], -1)
    with torch.no_grad():
        mass = net.log_prob(pts, ctx).exp().sum() * (grid[1] - grid[0]) ** 2
    assert abs(mass.item() - 1.0) < 2e-3


def test_flow_sample_inverts_log_prob_transform():
    torch.manual_seed(0)
    net = O.build_nsf(torch.randn(128, 6), torch.randn(128, 9), hidden_features=24, num_transforms=4, num_bins=5)
    O.randomize_lu(net, torch.Generator().manual_seed(1), 0.2)
    net.eval()
    z = torch.randn(50, 6)
    ctx = torch.randn(1, 9)
    with torch.no_grad():
        s = net.sample(50, context=ctx, noise=z)[0]
        zz, _ = net._transform(s, context=net._embedding_net(ctx.expand(50, -1)))
    assert (zz - z).abs().max() < 2e-4


def test_atomic_loss_matches_manual():
    torch.manual_seed(1)
    B, D, C, A = 12, 4, 5, 3
    net = O.build_nsf(torch.randn(64, D), torch.randn(64, C), hidden_features=8, num_transforms=2, num_bins=4)
    net.eval()
    prior = O.BoxUniform(-5 * torch.ones(D), 5 * torch.ones(D))
    theta, x = torch.randn(B, D), torch.randn(B, C)
    choices = torch.stack([torch.tensor([j for j in range(B) if j != b][:A - 1]) for b in range(B)])
    masks = torch.ones(B, 1)
    with torch.no_grad():
        out = O.log_prob_proposal_posterior_atomic(net, prior, theta, x, masks, A, True, choices)
        for b in range(B):
            ths = torch.cat([theta[b:b + 1], theta[choices[b]]])
            lp = net.log_prob(ths, x[b:b + 1].expand(A, -1))
            want = lp[0] - torch.logsumexp(lp, 0) + lp[0]
            assert abs(out[b] - want) < 1e-4


@pytest.mark.parametrize("name", ["c1", "odd", "c2", "c3", "c4"])
def test_oracle_reproduces_golden(name):
    """The committed vectors were written by tests/golden/make_golden.py; the oracle must still produce them."""
    net, g, (D, C, T, K, H, Bk, B, _) = golden_oracle_net(name)
    theta, x, z = torch.tensor(g["theta"]), torch.tensor(g["x"]), torch.tensor(g["z"])
    with torch.no_grad():
        lp = net.log_prob(theta, x).numpy()
        s1 = net.sample(B, context=x[:1], noise=z).numpy()[0]
    assert rel_err(lp, g["logp"]).max() < 1e-5
    assert rel_err(s1, g["samples_one_x"]).max() < 1e-4
    # fp32 oracle vs its own fp64 ground truth: the conditioning floor the 1e-4 bar sits on
    assert rel_err(g["logp"], g["logp64"]).max() < 1e-4


def test_trainer_runs_and_improves():
    torch.manual_seed(0)
    D, C, n = 3, 6, 600
    theta = torch.rand(n, D) * 6 - 3
    x = theta @ torch.randn(D, C) + 0.1 * torch.randn(n, C)
    prior = O.BoxUniform(-3 * torch.ones(D), 3 * torch.ones(D))
    inf = O.SNPE_C(prior, O.posterior_nn("nsf", hidden_features=16, num_transforms=2, num_bins=4))
    inf.append_simulations(theta, x)
    net = inf.train(training_batch_size=50, stop_after_epochs=2, max_num_epochs=6, generator=torch.Generator().manual_seed(0))
    v = inf._summary["validation_log_probs"]
    assert v[-1] > v[0]
    post = inf.build_posterior(net).set_default_x(x[:1])
    s = post.sample((200,))
    assert s.shape == (200, D) and prior.within(s).all()
    lp = post.log_prob(s[:10], norm_posterior=False)
    assert torch.isfinite(lp).all()
    assert post.log_prob(torch.full((1, D), 9.0), norm_posterior=False).item() == -float("inf")
