"""Synthetic stand-in for the ARCiS forward model (external Fortran, not available).

Produces (theta, spectrum) pairs with the shapes and the input distribution of the five BASELINE.json
configs, mirroring what reaches the flow in the reference:
  * theta: scrambled Sobol in [-3, 3]^D in round 0 (modules.py:106; the ``-ynorm`` space, floppityFUN.py:186);
  * spectrum: smooth baseline + Gaussian absorption features on a fixed grid whose amplitudes / centres /
    widths are fixed seeded maps of theta (fp64 on the host);
  * two-component mixing with ``frac`` and an additive ``offset`` on the second observation for config 4
    (simulator.py:359-369,385-393) with the "hotter component first" swap of modules.py:110-116;
  * noise: + sigma_j N(0,1), sigma_j = 1 % of the median flux (floppityFUN.py:385), then per-bin StandardScaler
    (floppityFUN.py:393,413,423) so the flow sees ~zero-mean / unit-variance fp32 x.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np

try:  # scipy is in the image; Sobol as in modules.py:106
    from scipy.stats.qmc import Sobol
except Exception:  # pragma: no cover
    Sobol = None


@dataclass(frozen=True)
class RetrievalConfig:
    name: str
    D: int
    C: int
    hidden: int
    transforms: int
    bins: int
    blocks: int
    atoms: int
    n_pairs: int
    mixing: bool = False
    segments: Tuple[int, ...] = ()


CONFIGS: Dict[str, RetrievalConfig] = {
    # BASELINE.json configs[0..3]; n_pairs rounded to the power of two Sobol needs
    "c1": RetrievalConfig("c1", D=10, C=100, hidden=50, transforms=5, bins=8, blocks=2, atoms=10, n_pairs=4096),
    "c2": RetrievalConfig("c2", D=20, C=500, hidden=128, transforms=10, bins=10, blocks=2, atoms=10, n_pairs=65536),
    "c3": RetrievalConfig("c3", D=30, C=2000, hidden=512, transforms=15, bins=16, blocks=4, atoms=10, n_pairs=524288,
                          segments=(700, 700, 600)),
    "c4": RetrievalConfig("c4", D=40, C=3000, hidden=128, transforms=10, bins=10, blocks=2, atoms=20, n_pairs=65536,
                          mixing=True, segments=(1500, 1500)),
}


def sobol_theta(n: int, D: int, tail_bound: float = 3.0, seed: int = 0) -> np.ndarray:
    """Round-0 proposal of the reference (modules.py:106)."""
    m = int(np.log2(n))
    if Sobol is None:
        rng = np.random.default_rng(seed)
        u = rng.random((2 ** m, D))
    else:
        u = Sobol(D, seed=seed).random_base2(m)
    return (-tail_bound + 2 * tail_bound * u).astype(np.float64)


class SpectrumGenerator:
    """g_C(theta): fixed seeded map from parameters to a C-bin transmission-like spectrum."""

    def __init__(self, n_par: int, n_bins: int, n_features: int = 12, seed: int = 1234):
        rng = np.random.default_rng(seed)
        self.n_par, self.n_bins, self.n_feat = n_par, n_bins, n_features
        self.wl = np.linspace(0.0, 1.0, n_bins)
        self.A_amp = rng.normal(0, 0.5, (n_par, n_features))
        self.A_cen = rng.normal(0, 0.15, (n_par, n_features))
        self.A_wid = rng.normal(0, 0.3, (n_par, n_features))
        self.c0 = np.linspace(0.05, 0.95, n_features)
        self.b_slope = rng.normal(0, 0.2, n_par)
        self.b_level = rng.normal(0, 0.2, n_par)

    def __call__(self, theta: np.ndarray, chunk: int = 8192) -> np.ndarray:
        theta = np.asarray(theta, dtype=np.float64) / 3.0
        out = np.empty((theta.shape[0], self.n_bins))
        for s in range(0, theta.shape[0], chunk):
            t = theta[s:s + chunk]
            amp = 0.3 * (1 + np.tanh(t @ self.A_amp))                 # [n, F]
            cen = self.c0 + 0.05 * np.tanh(t @ self.A_cen)
            wid = 0.02 * np.exp(0.5 * np.tanh(t @ self.A_wid))
            base = 1.0 + 0.1 * (t @ self.b_level)[:, None] + 0.1 * (t @ self.b_slope)[:, None] * (self.wl - 0.5)
            z = (self.wl[None, None, :] - cen[:, :, None]) / wid[:, :, None]
            out[s:s + chunk] = base - (amp[:, :, None] * np.exp(-0.5 * z * z)).sum(axis=1)
        return out


def simulate(cfg: RetrievalConfig, theta: np.ndarray) -> np.ndarray:
    """Noise-free spectra [n, C] for parameters in the [-3, 3]^D normalised space."""
    if not cfg.mixing:
        if cfg.segments:
            parts = [SpectrumGenerator(cfg.D, c, seed=1234 + i)(theta) for i, c in enumerate(cfg.segments)]
            return np.concatenate(parts, axis=1)
        return SpectrumGenerator(cfg.D, cfg.C)(theta)
    # config 4: 2 global + 18 + 18 component parameters + frac + offset_1 = 40; two observations concatenated
    n_glob, n_comp = 2, (cfg.D - 4) // 2
    glob = theta[:, :n_glob]
    comp1 = theta[:, n_glob:n_glob + n_comp]
    comp2 = theta[:, n_glob + n_comp:n_glob + 2 * n_comp]
    frac = (theta[:, -2] + 3.0) / 6.0           # [-3,3] -> [0,1]
    offset = 0.01 * theta[:, -1]
    specs = []
    for i, c in enumerate(cfg.segments):
        g = SpectrumGenerator(n_glob + n_comp, c, seed=1234 + i)
        s1 = g(np.concatenate([glob, comp1], axis=1))
        s2 = g(np.concatenate([glob, comp2], axis=1))
        s = frac[:, None] * s1 + (1 - frac[:, None]) * s2   # simulator.py:366-369
        if i == 1:
            s = s + offset[:, None]                          # simulator.py:385-393
        specs.append(s)
    return np.concatenate(specs, axis=1)


def swap_hotter_first(cfg: RetrievalConfig, theta: np.ndarray) -> np.ndarray:
    """modules.py:110-116: column n_global must be >= column init2 when two inputs are mixed."""
    if not cfg.mixing:
        return theta
    n_glob, n_comp = 2, (cfg.D - 4) // 2
    a, b = n_glob, n_glob + n_comp
    theta = theta.copy()
    sw = theta[:, a] < theta[:, b]
    theta[sw, a], theta[sw, b] = theta[sw, b], theta[sw, a]
    return theta


def make_dataset(cfg_name: str, n: Optional[int] = None, theta: Optional[np.ndarray] = None, noise_seed: int = 4321,
                 scaler: Optional[Tuple[np.ndarray, np.ndarray]] = None):
    """-> theta fp32 [n, D], x fp32 [n, C] (noisy, per-bin standard-scaled), (scaler mean, scale), noise sigma."""
    cfg = CONFIGS[cfg_name]
    n = cfg.n_pairs if n is None else n
    if theta is None:
        theta = sobol_theta(n, cfg.D)
    theta = swap_hotter_first(cfg, np.asarray(theta, dtype=np.float64))
    spec = simulate(cfg, theta)
    sigma = 0.01 * np.median(spec, axis=0)
    rng = np.random.default_rng(noise_seed)
    noisy = spec + sigma[None, :] * rng.standard_normal(spec.shape)
    if scaler is None:
        mean, scale = noisy.mean(axis=0), noisy.std(axis=0)
        scale[scale == 0] = 1.0
        scaler = (mean, scale)
    x = (noisy - scaler[0]) / scaler[1]
    return theta.astype(np.float32), x.astype(np.float32), scaler, sigma


def observation(cfg_name: str, scaler, theta_true: Optional[np.ndarray] = None, seed: int = 99):
    """A single 'observed' spectrum x_o [1, C] (default_x of floppity.py:223) and its true parameters."""
    cfg = CONFIGS[cfg_name]
    if theta_true is None:
        theta_true = np.random.default_rng(seed).uniform(-1.5, 1.5, (1, cfg.D))
    theta_true = swap_hotter_first(cfg, theta_true)
    spec = simulate(cfg, theta_true)
    return ((spec - scaler[0]) / scaler[1]).astype(np.float32), theta_true.astype(np.float32)
