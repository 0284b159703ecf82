"""Convergence diagnostics for the built-in sampler (ArviZ is not installable offline).

Rank-normalised split R-hat and bulk effective sample size as defined by Vehtari, Gelman,
Simpson, Carpenter & Buerkner (2021) -- the definitions behind ``az.rhat`` / ``az.ess`` that
the reference's users read off the InferenceData written by abdpymc-infer (abd.py:922-924).
Input everywhere: array of shape (chains, draws).
"""

from __future__ import annotations

import numpy as np
from scipy.special import ndtri
from scipy.stats import rankdata


def _split(x):
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[None]
    half = x.shape[1] // 2
    return np.concatenate([x[:, :half], x[:, x.shape[1] - half:]], axis=0)


def _z_scale(x):
    r = rankdata(x, method="average").reshape(x.shape)
    return ndtri((r - 0.375) / (x.size + 0.25))


def _autocov(x):
    """Autocovariance of each row by FFT (biased estimator, as Stan / ArviZ use)."""
    n = x.shape[1]
    m = 1 << int(np.ceil(np.log2(2 * n)))
    xc = x - x.mean(axis=1, keepdims=True)
    f = np.fft.rfft(xc, n=m, axis=1)
    return np.fft.irfft(f * np.conj(f), n=m, axis=1)[:, :n] / n


def _rhat_raw(x):
    n = x.shape[1]
    w = x.var(axis=1, ddof=1).mean()
    b = n * x.mean(axis=1).var(ddof=1)
    if w == 0:
        return 1.0 if b == 0 else np.inf
    return float(np.sqrt(((n - 1) / n * w + b / n) / w))


def _ess_raw(x):
    m, n = x.shape
    if n < 4:
        return float("nan")
    acov = _autocov(x)
    chain_var = acov[:, 0] * n / (n - 1.0)
    w = chain_var.mean()
    var_plus = w * (n - 1.0) / n
    if m > 1:
        var_plus += x.mean(axis=1).var(ddof=1)
    if var_plus == 0:
        return float(m * n)
    rho = 1.0 - (w - acov.mean(axis=0)) / var_plus
    rho[0] = 1.0
    # Geyer's initial monotone sequence over pairs
    tau, t, prev = -1.0, 0, np.inf
    while t + 1 < n:
        pair = rho[t] + rho[t + 1]
        if pair < 0:
            break
        pair = min(pair, prev)
        tau += 2.0 * pair
        prev = pair
        t += 2
    tau = max(tau, 1.0 / np.log10(m * n))
    return float(m * n / tau)


def rhat(x) -> float:
    """Rank-normalised split R-hat (max of bulk and folded), shape (chains, draws)."""
    s = _split(x)
    folded = np.abs(s - np.median(s))
    return max(_rhat_raw(_z_scale(s)), _rhat_raw(_z_scale(folded)))


def ess_bulk(x) -> float:
    """Bulk effective sample size: ESS of the rank-normalised split chains."""
    return _ess_raw(_z_scale(_split(x)))


def ess_mean(x) -> float:
    """ESS for the posterior mean (no rank normalisation)."""
    return _ess_raw(_split(x))


def summary(draws: dict) -> dict:
    """{name: (chains, draws)} -> {name: dict(mean, sd, ess_bulk, rhat)}."""
    out = {}
    for name, x in draws.items():
        x = np.asarray(x, dtype=np.float64)
        out[name] = dict(mean=float(x.mean()), sd=float(x.std(ddof=1)), ess_bulk=ess_bulk(x), rhat=rhat(x))
    return out
