"""Rows N3 / N4 of the scope table: the retrieval helpers around the flow.

The fixtures in tests/golden/golden_reference_helpers.npz were produced by the REFERENCE'S OWN functions
(tests/golden/make_reference_golden.py imports /root/reference/floppityFUN.py), so these are pinned parity tests:
the CPU oracle restatement and the host-side classes on CPU, the CUDA kernels with `-m gpu`."""
import os
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN

G = dict(np.load(os.path.join(GOLDEN, "golden_reference_helpers.npz")))
NAUG = int(G["pre_naug"])
N, C = G["spec"].shape


def _z(ci, r):
    np.random.seed(int(G["pre_seed0"]) + 10 * ci + r)
    return np.random.randn(N * NAUG, C)


def _spec(r):
    return G["spec"] if r == 0 else G["spec"][::-1].copy()


# ------------------------------------------------------------------------------------------------ CPU: oracle + host classes
def test_oracle_helpers_match_reference_outputs():
    from oracle import retrieval_oracle as R
    assert np.abs(R.normalizer_transform(G["norm_bounds"], 3.0, G["norm_Y"]) - G["norm_T"]).max() < 1e-13
    assert np.abs(R.normalizer_inverse(G["norm_bounds"], 3.0, G["norm_T"]) - G["norm_inv"]).max() < 1e-13
    assert np.abs(R.rm_mean(G["spec"]) - G["rm_mean"]).max() < 1e-14
    assert np.abs(R.sigma_res(G["spec"], G["obs"], G["noise"]) - G["sigma_res"]).max() < 1e-12
    lik = R.likelihood(G["obs"], G["noise"], G["lik_x"])
    assert np.abs(lik - G["lik"]).max() < 1e-9
    w, eff = R.importance_weights(G["lik"], G["is_logq"].astype(np.float64))
    assert np.abs(w - G["is_w"]).max() < 1e-12 and abs(eff - G["is_eff"]) < 1e-12
    logZ, logZ_err = R.evidence_from_IS(w, eff)
    assert abs(logZ - G["is_logZ"]) < 1e-12 and abs(logZ_err - G["is_logZ_err"]) < 1e-9
    ev = R.evidence(G["is_logq"].astype(np.float64), G["ev_logpi"].astype(np.float64), G["lik"])
    assert np.abs(ev - G["ev_logZ"]).max() < 1e-9


@pytest.mark.parametrize("ci", range(len(G["pre_combos"])))
def test_oracle_preprocess_matches_reference_outputs(ci):
    from oracle import retrieval_oracle as R
    rem, pca, xn, res = (bool(v) for v in G["pre_combos"][ci])
    base = R.rm_mean(G["spec"]) if rem else G["spec"]          # the fit never sees residuals (reference quirk)
    pc = R.fit_pca(base, 5) if pca else None
    t = (base - pc[0]) @ pc[1].T if pca else base
    sc = R.fit_scaler(t) if xn else None
    if pca:
        assert np.abs(pc[1] - G[f"pre{ci}_pca_components"]).max() < 1e-12
    if xn:
        assert np.abs(sc[1] - G[f"pre{ci}_scaler_scale"]).max() < 1e-12
    for r in (0, 1):
        out = R.preprocess_transform(_spec(r), _z(ci, r), G["noise"], NAUG, rem, res, G["obs"], pc, sc)
        assert out.dtype == np.float32 and out.shape == G[f"pre{ci}_r{r}_x"].shape
        assert np.abs(out - G[f"pre{ci}_r{r}_x"]).max() < 1e-6


def test_host_classes_match_reference_outputs():
    from floppity_deprecated_b200 import retrieval as RT
    nz = RT.Normalizer(G["norm_bounds"], 3.0)
    assert np.abs(nz.transform(G["norm_Y"]) - G["norm_T"]).max() < 1e-13
    assert np.abs(nz.inverse_transform(G["norm_T"]) - G["norm_inv"]).max() < 1e-13
    with pytest.raises(AssertionError):
        nz.transform(G["norm_Y"][:, :-1])
    rm = RT.rm_mean()
    assert np.abs(rm.transform(G["spec"]) - G["rm_mean"]).max() < 1e-14
    assert np.abs(rm.inverse_transform(G["rm_mean"]) - G["spec"]).max() < 1e-13
    sc = RT.ColumnScaler().fit(G["spec"])
    assert np.abs(sc.scale_ - G["pre0_scaler_scale"]).max() < 1e-14 and np.abs(sc.mean_ - G["pre0_scaler_mean"]).max() < 1e-14
    pc = RT.ColumnPCA(5).fit(G["spec"])
    assert np.abs(pc.components_ - G["pre2_pca_components"]).max() < 1e-12
    assert np.abs(pc.inverse_transform(pc.transform(G["spec"])) - G["spec"]).max() < 1.0     # lossy by construction
    assert RT.evidence_from_IS(G["is_w"], float(G["is_eff"]))[0] == pytest.approx(float(G["is_logZ"]), abs=1e-12)


def test_retrieval_needs_cuda_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from floppity_deprecated_b200 import retrieval as RT
    with pytest.raises(RuntimeError):
        RT.likelihood(G["obs"], G["noise"], G["lik_x"])


# ------------------------------------------------------------------------------------------------ GPU: the kernels
@pytest.mark.gpu
@pytest.mark.parametrize("ci", range(len(G["pre_combos"])))
def test_preprocess_matches_reference_outputs(ci, tmp_path):
    from floppity_deprecated_b200 import retrieval as RT
    rem, pca, xn, res = (bool(v) for v in G["pre_combos"][ci])
    theta = np.random.default_rng(0).uniform(-3, 3, (N, 6))
    args = types.SimpleNamespace(res_scale=res)
    for r in (0, 1):
        th, x_f, xscaler, pca_o = RT.preprocess(theta, _spec(r), r, N, G["obs"], G["noise"], NAUG, pca, 5, xn, C,
                                                rem, str(tmp_path), "cuda", args, z=torch.tensor(_z(ci, r)))
        want = G[f"pre{ci}_r{r}_x"]
        assert x_f.is_cuda and x_f.dtype == torch.float32 and tuple(x_f.shape) == want.shape
        assert th.shape == (N * NAUG, 6) and np.array_equal(th.cpu().numpy(), np.repeat(theta, NAUG, 0).astype(np.float32))
        # fp64 arithmetic in the kernel, fp32 only at the store (PCA projection: fp32 GEMM over 40 terms)
        # relative to max(1, |value|).  The kernel takes the standard-normal draw as fp32 (it is a random draw; the
        # reference's is fp64): in the residual mode that 6e-8 quantisation is amplified by 1 / scaler.scale_ ~ 25
        tol = 2e-5 if pca else (5e-6 if res else 1e-6)
        err = (np.abs(x_f.cpu().numpy() - want) / np.maximum(1.0, np.abs(want))).max()
        assert err < tol, err
    assert os.path.exists(tmp_path / "xscaler.p") and os.path.exists(tmp_path / "pca.p")
    if xn:
        assert np.abs(xscaler.scale_ - G[f"pre{ci}_scaler_scale"]).max() < 1e-12


@pytest.mark.gpu
def test_preprocess_draws_its_own_noise_with_the_right_statistics():
    from floppity_deprecated_b200 import retrieval as RT
    spec = np.ones((64, 300)) * 2.0
    noise = np.full(300, 0.5)
    out = RT.augment_and_scale(spec, noise, 50, generator=torch.Generator(device="cuda").manual_seed(3))
    assert tuple(out.shape) == (3200, 300)
    assert abs(float(out.mean()) - 2.0) < 5e-3 and abs(float(out.std()) - 0.5) < 5e-3
    assert RT.augment_and_scale(np.zeros((0, 300)), noise, 4).shape == (0, 300)       # empty round


@pytest.mark.gpu
def test_likelihood_is_evidence_match_reference_outputs():
    from floppity_deprecated_b200 import retrieval as RT

    class Fixed:
        def __init__(self, v):
            self.v = torch.tensor(v)

        def log_prob(self, Y, x=None):
            return self.v.clone()

    lik = RT.likelihood(G["obs"], G["noise"], G["lik_x"])
    assert np.abs(lik - G["lik"]).max() < 1e-9 * np.abs(G["lik"]).max()
    assert RT.likelihood(G["obs"], G["noise"], G["lik_x"][3]) == pytest.approx(G["lik"][3], rel=1e-11)
    w, eff = RT.IS(Fixed(G["is_logq"]), G["obs"], G["noise"], G["lik_x"], np.zeros((N, 6)))
    # log-weights of O(1e3) in fp64: a different summation order moves them by ~1e-10, the weights by the same (relative)
    assert np.abs(w - G["is_w"]).max() < 1e-7 and eff == pytest.approx(float(G["is_eff"]), rel=1e-7)
    logZ, logZ_err = RT.evidence_from_IS(w, eff)
    assert logZ == pytest.approx(float(G["is_logZ"]), abs=1e-7)
    ident = RT.do_nothing()
    ev = RT.evidence(Fixed(G["is_logq"]), Fixed(G["ev_logpi"]), G["lik_x"], np.zeros((N, 6)), G["obs"], G["noise"],
                     False, False, ident, ident, ident)
    assert np.abs(ev - G["ev_logZ"]).max() < 1e-6 * np.abs(G["ev_logZ"]).max()
