"""CPU restatement of the retrieval helpers around the flow (TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's
cpu_baseline leg may import this).

These functions live in the reference repository itself (plain numpy / sklearn), so -- unlike the flow oracle -- this
restatement IS pinned: tests/golden/golden_reference_helpers.npz holds outputs of the reference's own code
(tests/golden/make_reference_golden.py imports /root/reference/floppityFUN.py unmodified).

    Normalizer                 /root/reference/floppityFUN.py:177-194
    rm_mean.transform          /root/reference/floppityFUN.py:229-244
    sigma_res_scale.transform  /root/reference/floppityFUN.py:196-204
    preprocess (transform leg) /root/reference/floppityFUN.py:383-428
    likelihood                 /root/reference/floppityFUN.py:24-28
    IS / evidence_from_IS      /root/reference/floppityFUN.py:70-91
    evidence                   /root/reference/floppityFUN.py:30-41
"""
import numpy as np


def normalizer_transform(bounds, lims, Y):
    lo, hi = bounds[:, 0], bounds[:, 1]
    return (2 * lims) * (Y - lo) / (hi - lo) - lims


def normalizer_inverse(bounds, lims, Y):
    lo, hi = bounds[:, 0], bounds[:, 1]
    return (Y + lims) * (hi - lo) / (2 * lims) + lo


def rm_mean(spec):
    xbar = spec.mean(axis=1, keepdims=True)
    return np.concatenate([spec - xbar, xbar], axis=1)


def sigma_res(data, obs, noise):
    return (data - obs.reshape(1, -1)) / noise.reshape(1, -1)


def fit_scaler(X):
    """sklearn StandardScaler.fit: population std, zero-variance features scale 1."""
    mean = X.mean(axis=0)
    scale = X.std(axis=0)
    scale = np.where(scale == 0.0, 1.0, scale)
    return mean, scale


def fit_pca(X, n):
    """sklearn PCA(n_components=n).fit (full SVD, svd_flip on the right singular vectors' largest-|.| entries)."""
    mean = X.mean(axis=0)
    U, S, Vt = np.linalg.svd(X - mean, full_matrices=False)
    # sklearn svd_flip(u_based_decision=False since 1.5): sign so that the largest |entry| of each ROW of Vt is positive
    idx = np.argmax(np.abs(Vt), axis=1)
    signs = np.sign(Vt[np.arange(Vt.shape[0]), idx])
    return mean, (Vt * signs[:, None])[:n]


def preprocess_transform(spec, z, noise, naug, rem_mean_on, res_mode, obs, pca, scaler):
    """The transform leg of preprocess(): spec [N, C] fp64, z [N naug, C] the standard-normal draw,
    pca = (mean, components) or None, scaler = (mean, scale) or None -> fp32 [N naug, C']"""
    aug = np.repeat(spec, naug, axis=0) + noise * z
    if res_mode:
        v = sigma_res(aug, obs, noise)
    else:
        v = rm_mean(aug) if rem_mean_on else aug
    if pca is not None:
        v = (v - pca[0]) @ pca[1].T
    if scaler is not None:
        v = (v - scaler[0]) / scaler[1]
    return v.astype(np.float32)


def likelihood(obs, err, x):
    """Row-wise Gaussian log-likelihood; x [N, C] -> [N]"""
    x = np.atleast_2d(x)
    return (-np.log(np.sqrt(2 * np.pi) * err)[None, :] - (obs[None, :] - x) ** 2 / (2 * err[None, :] ** 2)).sum(axis=1)


def importance_weights(log_L, log_q):
    log_w = log_L - log_q
    w = np.exp(log_w - log_w.max())
    eff = (1 / len(w)) * w.sum() ** 2 / (w ** 2).sum()
    return w, eff


def evidence_from_IS(w, eff):
    Z = w.sum() / len(w)
    Z_err = np.sqrt((1 - eff) / (len(w) * eff))
    return np.log(Z), Z_err / Z


def evidence(P, pi, L):
    v = -(P - pi - L)
    return np.array([np.median(v), np.percentile(v, 16), np.percentile(v, 84)])
