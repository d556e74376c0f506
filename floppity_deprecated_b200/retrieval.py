"""The data-side neighbours of the flow in a FlopPITy round, on the GPU.

Host-side mirror of the helper functions ``floppity.py`` calls right before and right after the flow -- same
names, argument order and return values as ``/root/reference/floppityFUN.py`` so the round loop keeps working:

    Normalizer(prior_bounds, lims).transform / inverse_transform      floppityFUN.py:177-194   (floppity.py:123,255)
    preprocess(np_theta, arcis_spec, r, samples_per_round, obs_spec, noise_spec, naug, do_pca, n_pca, xnorm, nwvl,
               rem_mean, output, device, args) -> theta_aug, x_f, xscaler, pca      floppityFUN.py:383-428 (floppity.py:192)
    likelihood(obs, err, x)                                            floppityFUN.py:24-28
    IS(q, obs, err, x, Y) -> w_i, eff ; evidence_from_IS(w_i, eff)     floppityFUN.py:70-91     (floppity.py:204-212)
    evidence(posterior, prior, samples, Y, obs, err, do_pca, xnorm, rem_mean, xscaler, pca)     floppityFUN.py:30-41

The per-round fits (StandardScaler / PCA statistics over the N simulated spectra of round 0) stay on the host in
fp64 exactly as sklearn computes them -- they are one-off reductions that the reference pickles to ``xscaler.p`` /
``pca.p`` -- while the part that scales with ``naug`` (noise augmentation of every spectrum, mean removal / residual
scaling, per-bin affine scaling, the fp32 cast) and the row-wise likelihoods run in the CUDA kernels of
``csrc/retrieval_kernels.cu``.  No CPU fallback: the transforms raise without a CUDA device.
"""
from __future__ import annotations

import os
import pickle
from typing import Optional

import numpy as np
import torch

from . import _lib


class Normalizer:
    """Parameter transformer: prior box -> [-lims, lims]^D (floppityFUN.py:177-194).  N x D numpy on the host, as the
    reference uses it (floppity.py:123,255; the flow works in the normalised space)."""

    def __init__(self, prior_bounds, lims):
        self.bounds = prior_bounds
        self.lims = lims

    def transform(self, Y):
        assert len(self.bounds) == Y.shape[1], "Dimensionality of prior and parameters doesn't match!"
        b = np.asarray(self.bounds, dtype=np.float64)
        return (2 * self.lims) * (Y - b[:, 0]) / (b[:, 1] - b[:, 0]) - self.lims

    def inverse_transform(self, Y):
        assert len(self.bounds) == Y.shape[1], "Dimensionality of prior and parameters doesn't match!"
        b = np.asarray(self.bounds, dtype=np.float64)
        return (Y + self.lims) * (b[:, 1] - b[:, 0]) / (2 * self.lims) + b[:, 0]


class do_nothing:
    """floppityFUN.py:217-232: identity stand-in for a disabled preprocessing stage."""

    def __init__(self, **kwargs):
        pass

    def fit(self, data):
        return None

    def transform(self, data):
        return data

    def inverse_transform(self, data):
        return data


class ColumnScaler:
    """What the reference pickles as ``xscaler.p`` (sklearn StandardScaler, floppityFUN.py:393,413): per-feature mean
    and population standard deviation (zero variance -> scale 1), fp64."""

    def fit(self, X):
        X = np.asarray(X, dtype=np.float64)
        self.mean_ = X.mean(axis=0)
        self.scale_ = X.std(axis=0)
        self.scale_[self.scale_ == 0.0] = 1.0
        self.n_features_in_ = X.shape[1]
        return self

    def transform(self, X):
        return (np.asarray(X, dtype=np.float64) - self.mean_) / self.scale_

    def inverse_transform(self, X):
        return np.asarray(X, dtype=np.float64) * self.scale_ + self.mean_


class ColumnPCA:
    """What the reference pickles as ``pca.p`` (sklearn PCA(n_components), floppityFUN.py:388,410): mean and the
    leading right singular vectors with sklearn's sign convention, fp64."""

    def __init__(self, n_components):
        self.n_components = int(n_components)

    def fit(self, X):
        X = np.asarray(X, dtype=np.float64)
        self.mean_ = X.mean(axis=0)
        _, _, Vt = np.linalg.svd(X - self.mean_, full_matrices=False)
        idx = np.argmax(np.abs(Vt), axis=1)
        signs = np.sign(Vt[np.arange(Vt.shape[0]), idx])
        self.components_ = (Vt * signs[:, None])[: self.n_components]
        return self

    def transform(self, X):
        return (np.asarray(X, dtype=np.float64) - self.mean_) @ self.components_.T

    def inverse_transform(self, T):
        return np.asarray(T, dtype=np.float64) @ self.components_ + self.mean_


class rm_mean:
    """floppityFUN.py:229-251: subtract each spectrum's mean and append it as an extra feature."""

    def transform(self, arcis_spec):
        s = np.asarray(arcis_spec, dtype=np.float64)
        xbar = s.mean(axis=1, keepdims=True)
        return np.concatenate([s - xbar, xbar], axis=1)

    def inverse_transform(self, normed):
        n = np.asarray(normed, dtype=np.float64)
        return n[:, :-1] + n[:, -1:]


def _dev(device) -> torch.device:
    d = _lib.require_cuda(None if device in (None, "cpu", "cuda") else device)
    return d


def _d64(a, dev) -> torch.Tensor:
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64))).to(dev)


def augment_and_scale(arcis_spec, noise_spec, naug: int, mode: int = 0, obs_spec=None, shift=None, scale=None,
                      z: Optional[torch.Tensor] = None, device=None, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """One fused pass: ``np.repeat(spec, naug) + noise * randn`` -> mean removal (mode 1) / residual scaling (mode 2)
    -> ``(v - shift) / scale`` -> fp32 [N * naug, C (+1)] on the device.  ``z`` ([N * naug, C] standard normal)
    replaces the draw so tests do not depend on RNG streams (the reference draws from numpy's global RNG)."""
    dev = _dev(device)
    spec = _d64(arcis_spec, dev)
    n, C = spec.shape
    rows = n * int(naug)
    if z is None:
        z = torch.randn(rows, C, device=dev, dtype=torch.float32, generator=generator)
    else:
        z = torch.as_tensor(z).to(dev, torch.float32).contiguous()
        if z.shape != (rows, C):
            raise ValueError(f"z must be [{rows}, {C}]")
    width = C + (1 if mode == 1 else 0)
    out = torch.empty(rows, width, dtype=torch.float32, device=dev)
    sh = None if shift is None else _d64(shift, dev)
    sc = None if scale is None else _d64(scale, dev)
    if sh is not None and sh.numel() != width or sc is not None and sc.numel() != width:
        raise ValueError("shift / scale must have one entry per output column")
    ob = None if obs_spec is None else _d64(obs_spec, dev).reshape(-1)
    nz = _d64(noise_spec, dev).reshape(-1)          # (named: a temporary could be recycled before the launch)
    if nz.numel() != C or (ob is not None and ob.numel() != C):
        raise ValueError("noise_spec / obs_spec must have one entry per spectral bin")
    _lib.check(_lib.lib().fp_preprocess_rows(_lib.ptr(spec), _lib.ptr(z), _lib.ptr(nz), _lib.ptr(ob), _lib.ptr(sh),
                                             _lib.ptr(sc), _lib.ptr(out), n, C, int(naug), int(mode), width,
                                             _lib.stream()), "fp_preprocess_rows")
    return out


def preprocess(np_theta, arcis_spec, r, samples_per_round, obs_spec, noise_spec, naug, do_pca, n_pca, xnorm, nwvl,
               rem_mean, output, device, args, z: Optional[torch.Tensor] = None):
    """``floppityFUN.preprocess`` (floppityFUN.py:383-428) with the same arguments and return values
    ``(theta_aug, x_f, xscaler, pca)``; ``x_f`` / ``theta_aug`` are fp32 tensors on the CUDA device.

    Round 0 fits the PCA / scaler on the un-augmented spectra and pickles them to ``output`` (pca.p, xscaler.p);
    every round re-loads them, as the reference does.  Reference quirk kept: with ``args.res_scale`` the statistics are
    still fitted on the raw spectra (floppityFUN.py:407-413) although they are applied to the residuals."""
    dev = _dev(device)
    spec = np.asarray(arcis_spec, dtype=np.float64)
    res = bool(getattr(args, "res_scale", False))
    pca = ColumnPCA(n_pca) if do_pca else do_nothing()
    xscaler = ColumnScaler() if xnorm else do_nothing()
    rem = rm_mean() if rem_mean else do_nothing()
    if r == 0:
        pca.fit(rem.transform(spec))
        with open(os.path.join(output, "pca.p"), "wb") as f:
            pickle.dump(pca, f)
        xscaler.fit(pca.transform(rem.transform(spec)))
        with open(os.path.join(output, "xscaler.p"), "wb") as f:
            pickle.dump(xscaler, f)
    with open(os.path.join(output, "pca.p"), "rb") as f:
        pca = pickle.load(f)
    with open(os.path.join(output, "xscaler.p"), "rb") as f:
        xscaler = pickle.load(f)

    theta_aug = torch.as_tensor(np.repeat(np.asarray(np_theta), naug, axis=0)).to(dev, torch.float32)
    mode = 2 if res else (1 if rem_mean else 0)
    has_pca, has_scaler = hasattr(pca, "components_"), hasattr(xscaler, "scale_")
    if has_pca:
        # centre with the PCA mean in the fused pass (fp64 arithmetic, fp32 store), project with one fp32 GEMM, then the
        # scaler on the n_pca features.  The reference projects and scales in fp64 (sklearn) and casts once at the end; the
        # fp32 projection differs from that by < 2e-5 relative on the reference's own outputs
        # (tests/test_retrieval.py::test_preprocess_matches_reference_outputs, PCA combinations)
        centred = augment_and_scale(spec, noise_spec, naug, mode, obs_spec, shift=pca.mean_, z=z, device=dev)
        comp = torch.as_tensor(pca.components_).to(dev, torch.float32).contiguous()      # [n_pca, C']
        x_f = _project(centred, comp)
        if has_scaler:
            x_f = _affine32(x_f, xscaler.mean_, xscaler.scale_)
    else:
        x_f = augment_and_scale(spec, noise_spec, naug, mode, obs_spec,
                                shift=xscaler.mean_ if has_scaler else None,
                                scale=xscaler.scale_ if has_scaler else None, z=z, device=dev)
    # floppityFUN.py:424-425: the diagnostic the reference prints every round (on the spectra the PCA was fitted on; the
    # reference applies it to the raw spectra, which only differs -- it crashes there -- when rem_mean and do_pca are combined)
    max_err = float(np.max(pca_error(pca, rem.transform(spec))))
    print(f"Maximum PCA reconstruction error: {max_err}")
    return theta_aug, x_f, xscaler, pca


def _project(x: torch.Tensor, comp: torch.Tensor) -> torch.Tensor:
    """x [R, C] @ comp[n, C]^T through the library's fp32 GEMM (C-ABI fp_gemm)."""
    R, C = x.shape
    n = comp.shape[0]
    out = torch.empty(R, n, dtype=torch.float32, device=x.device)
    a = _lib.GemmArgs()
    a.A, a.lda, a.a_kmajor = x.data_ptr(), C, 1
    a.B, a.ldb, a.b_kmajor = comp.data_ptr(), C, 1
    a.C, a.ldc, a.M, a.N, a.K, a.flags, a.split_k = out.data_ptr(), n, R, n, C, 0, 1
    import ctypes
    _lib.check(_lib.lib().fp_gemm(ctypes.byref(a), _lib.stream()), "fp_gemm")
    return out


def _affine32(x: torch.Tensor, mean, scale) -> torch.Tensor:
    out = torch.empty_like(x)
    m = torch.as_tensor(np.asarray(mean)).to(x.device, torch.float32).contiguous()
    s = torch.as_tensor(np.asarray(scale)).to(x.device, torch.float32).contiguous()
    _lib.check(_lib.lib().fp_standardize(_lib.ptr(x), _lib.ptr(m), _lib.ptr(s), _lib.ptr(out), x.shape[0], x.shape[1],
                                         _lib.stream()), "fp_standardize")
    return out


def pca_error(pca, spectra):
    """floppityFUN.py:430-438"""
    return abs(spectra - pca.inverse_transform(pca.transform(spectra)))


# ---------------------------------------------------------------------------------------------------------
def likelihood(obs, err, x, device=None):
    """Gaussian log-likelihood (floppityFUN.py:24-28).  ``x`` may be one spectrum [C] (returns a float, as the
    reference) or a batch [N, C] (returns an fp64 numpy vector): the batch form is one kernel launch."""
    dev = _dev(device)
    xa = np.asarray(x, dtype=np.float64)
    single = xa.ndim == 1
    xd = _d64(xa.reshape(1, -1) if single else xa, dev)
    out = torch.empty(xd.shape[0], dtype=torch.float64, device=dev)
    ob, er = _d64(obs, dev).reshape(-1), _d64(err, dev).reshape(-1)
    if ob.numel() != xd.shape[1] or er.numel() != xd.shape[1]:
        raise ValueError("obs / err must have one entry per spectral bin")
    _lib.check(_lib.lib().fp_gauss_loglik(_lib.ptr(ob), _lib.ptr(er), _lib.ptr(xd), _lib.ptr(out), xd.shape[0],
                                          xd.shape[1], _lib.stream()), "fp_gauss_loglik")
    o = out.cpu().numpy()
    return float(o[0]) if single else o


def IS(q, obs, err, x, Y):
    """Importance weights of the simulated spectra under proposal ``q`` (floppityFUN.py:70-84) -> (w_i, eff)."""
    log_L = likelihood(obs, err, np.asarray(x))
    log_q = q.log_prob(torch.tensor(np.asarray(Y)).float()).detach().cpu().numpy().astype(np.float64)
    log_w = log_L - log_q
    w_i = np.exp(log_w - log_w.max())
    eff = (1 / len(w_i)) * (np.sum(w_i)) ** 2 / (np.sum(w_i ** 2))
    return w_i, eff


def evidence_from_IS(w_i, eff):
    """floppityFUN.py:86-91"""
    Z = (1 / len(w_i)) * np.sum(w_i)
    Z_err = np.sqrt((1 - eff) / (len(w_i) * eff))
    return np.log(Z), Z_err / Z


def evidence(posterior, prior, samples, Y, obs, err, do_pca, xnorm, rem_mean, xscaler, pca):
    """floppityFUN.py:30-41: median and 16/84 percentiles of -(log q(theta|x_o) - log p(theta) - log L)."""
    L = likelihood(obs, err, np.asarray(samples))
    default_x = xscaler.transform(pca.transform(rem_mean.transform(np.asarray(obs).reshape(1, -1))))
    P = posterior.log_prob(torch.tensor(np.asarray(Y)), x=default_x)
    pi = prior.log_prob(torch.tensor(np.asarray(Y)))
    v = -(P.detach().cpu().numpy().astype(np.float64) - pi.detach().cpu().numpy().astype(np.float64) - L)
    logZ = np.empty(3)
    logZ[0] = np.median(v)
    logZ[2] = np.percentile(v, 84)
    logZ[1] = np.percentile(v, 16)
    return logZ
