"""Convergence diagnostics for the throughput metric "genes/hour to bulk-ESS >= 400".

The reference has no ESS / R-hat code (CmdStan's stansummary is never called), so
these follow Vehtari, Gelman, Simpson, Carpenter, Buerkner (2021): rank-normalised split-R-hat
and bulk-ESS with Geyer's initial monotone sequence.  numpy only; input shape (chains, draws).
"""
from __future__ import annotations

import numpy as np
from scipy import stats as _st


def _split(x):
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[1] // 2
    return np.concatenate([x[:, :n], x[:, x.shape[1] - n:]], axis=0)


def _rank_normalise(x):
    r = _st.rankdata(x.reshape(-1), method="average").reshape(x.shape)
    return _st.norm.ppf((r - 0.375) / (x.size + 0.25))


def _rhat(x):
    m, n = x.shape
    cm = x.mean(1)
    B = n * cm.var(ddof=1)
    W = x.var(1, ddof=1).mean()
    if W == 0:
        return 1.0
    return float(np.sqrt(((n - 1) / n * W + B / n) / W))


def split_rhat(x) -> float:
    """max of rank-normalised split-R-hat on x and on the folded draws |x - median|."""
    x = np.asarray(x, dtype=np.float64)
    z = _rank_normalise(_split(x))
    zf = _rank_normalise(_split(np.abs(x - np.median(x))))
    return max(_rhat(z), _rhat(zf))


def _autocov(x):
    n = x.shape[-1]
    m = 1 << int(np.ceil(np.log2(2 * n)))
    xc = x - x.mean(-1, keepdims=True)
    f = np.fft.rfft(xc, m, axis=-1)
    ac = np.fft.irfft(f * np.conj(f), m, axis=-1)[..., :n]
    return ac / n


def _ess(x) -> float:
    m, n = x.shape
    if n < 4:
        return float(m * n)
    acov = _autocov(x)
    cm = x.mean(1)
    W = (acov[:, 0] * n / (n - 1.0)).mean()
    var_plus = W * (n - 1.0) / n
    if m > 1:
        var_plus += cm.var(ddof=1)
    if var_plus <= 0:
        return float(m * n)
    rho = 1.0 - (W - acov.mean(0) * n / (n - 1.0)) / var_plus
    rho[0] = 1.0
    # Geyer: sums of adjacent pairs, truncated at the first negative pair, made monotone
    tau = -1.0
    prev = np.inf
    t = 0
    while t + 1 < n:
        p = rho[t] + rho[t + 1]
        if p < 0:
            break
        p = min(p, prev)
        tau += 2.0 * p
        prev = p
        t += 2
    tau = max(tau, 1.0 / np.log10(max(m * n, 10)))
    return float(m * n / tau)


def ess_bulk(x) -> float:
    x = np.asarray(x, dtype=np.float64)
    return _ess(_rank_normalise(_split(x)))


def summarize(draws: np.ndarray):
    """draws: (chains, draws, P) -> (min bulk-ESS, max split-R-hat) over the P columns."""
    ess = [ess_bulk(draws[:, :, j]) for j in range(draws.shape[2])]
    rh = [split_rhat(draws[:, :, j]) for j in range(draws.shape[2])]
    return float(np.min(ess)), float(np.max(rh))
