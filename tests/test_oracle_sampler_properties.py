"""What the sampler restatement must satisfy whatever Stan release it is compared with (the Stan boundary is unpinned,
SURVEY.md §8c): the oracle's NUTS — which the product's control logic reproduces draw for draw
(tests/test_nuts_core_vs_oracle.py) — leaves known targets invariant, and its adaptation lands where Stan's is designed
to land (accept_stat near delta, the metric near the marginal variances, healthy energy)."""
import numpy as np
import pytest

import oracle
from csplotch_b200 import diagnostics as dg


def _run(fn, D, nw, ns, chains=4, seed=5, **kw):
    outs = []
    for c in range(chains):
        cfg = oracle.nuts_cfg(nw, ns, seed=seed, **kw)
        r = oracle.nuts_generic(fn, D, cfg, gene_id=17, chain_id=c + 1)
        assert r["status"] == 0
        outs.append(r)
    return outs, np.stack([r["draws"][:, 7:] for r in outs]), np.stack([r["draws"][:, :7] for r in outs])


def _z(draws, truth):
    """(mean - truth) in units of the Monte Carlo standard error, chains pooled."""
    se = draws.std() / np.sqrt(dg.ess_bulk(draws))
    return abs(draws.mean() - truth) / se


def test_anisotropic_gaussian_is_invariant_and_adaptation_lands():
    sd = np.array([0.05, 0.3, 1.0, 1.0, 2.0, 7.0, 30.0])

    def fn(q):
        return -0.5 * np.sum((q / sd) ** 2), -q / sd ** 2

    outs, th, diag = _run(fn, len(sd), 300, 1200)
    for j in range(len(sd)):
        assert _z(th[:, :, j], 0.0) < 4.5, j
        assert _z(th[:, :, j] ** 2, sd[j] ** 2) < 4.5, j
        assert dg.split_rhat(th[:, :, j]) < 1.02
    for r in outs:
        assert np.all(np.abs(np.log(r["inv_metric"] / sd ** 2)) < np.log(2.0))     # windowed variance estimate
    acc = diag[:, :, 1].mean()
    assert 0.72 < acc < 0.95                                                      # delta = 0.8, exp(x_bar) step size
    assert diag[:, :, 5].sum() == 0 and diag[:, :, 3].max() <= 4                   # no divergences, short trees
    e = diag[:, :, 6]
    bfmi = (np.diff(e, axis=1) ** 2).mean(axis=1) / e.var(axis=1)
    assert bfmi.min() > 0.5
    # with the adapted metric every coordinate mixes almost independently: ESS of the order of the draw count
    assert min(dg.ess_bulk(th[:, :, j]) for j in range(len(sd))) > 1500


def test_higher_delta_means_smaller_steps_and_higher_acceptance():
    def fn(q):
        return -0.5 * np.sum(q * q), -q

    res = {}
    for delta in (0.6, 0.95):
        outs, th, diag = _run(fn, 20, 200, 400, chains=2, delta=delta)
        res[delta] = (np.mean([r["stepsize"] for r in outs]), diag[:, :, 1].mean())
        # the averaged iterate exp(x_bar) is a little smaller than the last step size: acceptance ends above delta
        assert -0.05 < res[delta][1] - delta < 0.2
    assert res[0.95][0] < res[0.6][0] and res[0.95][1] > res[0.6][1]


def test_log_gamma_target_with_jacobian():
    """x ~ Gamma(3, 1) sampled on u = log x: density exp(3u - e^u) (the Jacobian included, as Stan does for a
    lower-bounded parameter such as tau or sigma)."""
    def fn(q):
        return float(np.sum(3.0 * q - np.exp(q))), 3.0 - np.exp(q)

    outs, th, diag = _run(fn, 3, 200, 2000)
    x = np.exp(th)
    for j in range(3):
        assert _z(x[:, :, j], 3.0) < 4.5 and _z((x[:, :, j] - 3.0) ** 2, 3.0) < 4.5
        assert _z(th[:, :, j], 0.9227843350984671) < 4.5                           # E[log x] = digamma(3)


def test_correlated_gaussian_with_a_diagonal_metric():
    rho = 0.9
    P = np.linalg.inv(np.array([[1.0, rho], [rho, 1.0]]))

    def fn(q):
        return -0.5 * q @ P @ q, -P @ q

    outs, th, diag = _run(fn, 2, 200, 2500)
    assert _z(th[:, :, 0] * th[:, :, 1], rho) < 4.5
    assert _z(th[:, :, 0] ** 2, 1.0) < 4.5 and _z(th[:, :, 1], 0.0) < 4.5
    assert dg.split_rhat(th[:, :, 0]) < 1.02


def test_divergences_are_flagged_and_max_depth_is_respected():
    """Neal's funnel at a fixed large step size: divergent__ must fire (energy error > 1000); a badly scaled target must
    stop at max_depth with 2^depth - 1 leapfrogs."""
    def funnel(q):
        v, x = q[0], q[1:]
        lp = -0.5 * v * v / 9.0 - 0.5 * np.sum(x * x) * np.exp(-v) - 0.5 * len(x) * v
        g = np.empty_like(q)
        g[0] = -v / 9.0 + 0.5 * np.sum(x * x) * np.exp(-v) - 0.5 * len(x)
        g[1:] = -x * np.exp(-v)
        return lp, g

    cfg = oracle.nuts_cfg(0, 300, seed=3, stepsize=2.0)
    r = oracle.nuts_generic(funnel, 6, cfg, q_init=np.zeros(6))
    assert r["status"] == 0 and r["draws"][:, 5].sum() > 0

    # scales 1 and 1e4 with a unit metric: the step size found by init_stepsize (run even with num_warmup=0, as
    # CmdStan does) is set by the narrowest direction, the U-turn of the widest needs 1e4 steps: trees that are not cut short by a sub-tree turning in the unit
    # directions end at max_depth
    sd = np.array([1.0, 1.0, 1e4])

    def wide(q):
        return -0.5 * np.sum((q / sd) ** 2), -q / sd ** 2

    cfg = oracle.nuts_cfg(0, 20, seed=3, max_depth=5)
    r = oracle.nuts_generic(wide, 3, cfg, q_init=0.5 * sd)
    depth, nleap = r["draws"][:, 3], r["draws"][:, 4]
    assert depth.max() == 5 and (depth == 5).sum() >= 3 and (nleap[depth == 5] == 31).all()    # 2^5 - 1, never more
    assert (nleap < 2.0 ** (depth + 1)).all() and (nleap >= 2.0 ** depth - 1).all() and r["draws"][:, 5].sum() == 0
    assert 0.1 < r["draws"][0, 2] < 4.0 and len(set(r["draws"][:, 2])) == 1             # no adaptation: one step size
