"""Bulk-ESS / split-R-hat behind the BASELINE metric "genes/hour to bulk-ESS >= 400" (csplotch_b200/diagnostics.py;
the reference has no such code): checked against closed forms — AR(1) chains have
ESS = N (1 - rho) / (1 + rho), i.i.d. draws ESS ~ N and R-hat ~ 1 — and against the failure modes they must flag."""
import numpy as np

from csplotch_b200 import diagnostics as dg


def _ar1(rng, m, n, rho):
    x = np.empty((m, n))
    x[:, 0] = rng.normal(size=m)
    e = rng.normal(size=(m, n)) * np.sqrt(1 - rho * rho)
    for t in range(1, n):
        x[:, t] = rho * x[:, t - 1] + e[:, t]
    return x


def test_ess_of_ar1_chains_matches_the_closed_form():
    rng = np.random.default_rng(0)
    m, n = 4, 4000
    for rho in (0.0, 0.5, 0.9):
        got = np.mean([dg.ess_bulk(_ar1(rng, m, n, rho)) for _ in range(8)])
        want = m * n * (1 - rho) / (1 + rho)
        assert abs(got - want) < 0.12 * want, (rho, got, want)


def test_antithetic_chains_exceed_n_and_constant_chains_do_not_crash():
    rng = np.random.default_rng(1)
    x = _ar1(rng, 4, 2000, -0.5)
    assert dg.ess_bulk(x) > 1.5 * x.size                        # NUTS draws are often antithetic: ESS > N is real
    assert np.isfinite(dg.ess_bulk(np.ones((4, 100)))) and dg.split_rhat(np.ones((4, 100))) == 1.0


def test_rhat_flags_what_it_must():
    rng = np.random.default_rng(2)
    good = rng.normal(size=(4, 1000))
    assert dg.split_rhat(good) < 1.01
    shifted = good + np.array([0.0, 0.0, 0.0, 1.0])[:, None]    # one chain elsewhere
    assert dg.split_rhat(shifted) > 1.05 and dg.ess_bulk(shifted) < 100
    scaled = good * np.array([1.0, 1.0, 1.0, 3.0])[:, None]     # same location, different scale: the folded R-hat
    assert dg.split_rhat(scaled) > 1.05
    drift = good + np.linspace(0, 2, 1000)[None, :]             # non-stationary within chains: the split halves differ
    assert dg.split_rhat(drift) > 1.05
    heavy = rng.standard_cauchy(size=(4, 1000))                 # rank normalisation: infinite variance is no obstacle
    assert dg.split_rhat(heavy) < 1.02 and dg.ess_bulk(heavy) > 2000


def test_summarize_takes_the_worst_column():
    rng = np.random.default_rng(3)
    d = rng.normal(size=(4, 500, 3))
    d[:, :, 1] = _ar1(rng, 4, 500, 0.95)
    ess, rhat = dg.summarize(d)
    assert ess == dg.ess_bulk(d[:, :, 1]) and ess < 300
    assert rhat == max(dg.split_rhat(d[:, :, j]) for j in range(3))
