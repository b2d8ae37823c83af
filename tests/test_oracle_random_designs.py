"""Random tiny designs (SURVEY.md App. C tier 1, "random tiny designs via hypothesis"): the C oracle against the
independent torch-autograd transcription of the .stan files on shapes no fixed fixture covers — single-tissue and
single-covariate designs, 1 x n strips, K = 1 compositions, all-zero and very large counts, every level depth —
and the edge-loop (#1) and CSR (#3) code paths of the CAR / ZIP terms against each other."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle
from csplotch_b200 import synth
from tests.stan_transcript import stan_log_prob_grad

FAMS = ["splotch_stan_model", "comp_splotch_stan_model", "splotch_stan_model_csr"]

design = st.fixed_dictionaries(dict(
    family=st.sampled_from(FAMS),
    n_levels=st.integers(1, 3), zi=st.integers(0, 1), car=st.integers(0, 1),
    K=st.integers(1, 4), C=st.integers(1, 5),
    tissues=st.lists(st.tuples(st.integers(1, 4), st.integers(2, 4)), min_size=1, max_size=4),
    seed=st.integers(0, 10_000),
    counts=st.sampled_from(["model", "zeros", "large"]),
    scale=st.sampled_from([0.3, 1.0, 2.5]),
))


def _gene(cfg):
    shared = synth.shared_tiny(seed=cfg["seed"], family=cfg["family"], n_levels=cfg["n_levels"], zi=cfg["zi"],
                               car=cfg["car"], K=cfg["K"], C=cfg["C"], tissues=tuple(cfg["tissues"]))
    genes = synth.simulate_counts(shared, 1, cfg["seed"] + 1)
    N = int(np.sum(shared["N_spots"]))
    rng = np.random.default_rng(cfg["seed"] + 2)
    if cfg["counts"] == "zeros":
        genes["counts"][:] = 0
    elif cfg["counts"] == "large":
        genes["counts"][0] = rng.integers(0, 50_000, N)
    if cfg["family"] == "comp_splotch_stan_model":
        genes["beta_prior_mean"] = rng.normal(0, 1, genes["beta_prior_mean"].shape)
        genes["beta_prior_std"] = rng.uniform(0.3, 3, genes["beta_prior_std"].shape)
    return synth.gene_dict(shared, genes, 0), rng


@settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(design)
def test_oracle_matches_transcription_on_random_designs(cfg):
    d, rng = _gene(cfg)
    o = oracle.Oracle(d, cfg["family"])
    th = rng.normal(0, cfg["scale"], o.dim)
    for jac in (True, False):
        lp, g = o.log_prob_grad(th, jacobian=jac)
        lp2, g2 = stan_log_prob_grad(d, cfg["family"], th, jacobian=jac)
        if not (np.isfinite(lp2) and np.all(np.isfinite(g2))):
            assert not (np.isfinite(lp) and np.all(np.isfinite(g)))
            continue
        assert abs(lp - lp2) <= 1e-10 * max(1.0, abs(lp2))
        assert np.max(np.abs(g - g2) / np.maximum(1.0, np.abs(g2))) < 1e-10


@settings(max_examples=25, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(design)
def test_edge_loop_and_csr_models_agree_on_random_designs(cfg):
    """#1 == #3: same parameters, same lp; only the CAR and ZIP summation order differs."""
    cfg = dict(cfg, family="splotch_stan_model")
    d, rng = _gene(cfg)
    a, b = oracle.Oracle(d, "splotch_stan_model"), oracle.Oracle(d, "splotch_stan_model_csr")
    assert a.dim == b.dim
    th = rng.normal(0, cfg["scale"], a.dim)
    la, ga = a.log_prob_grad(th)
    lb, gb = b.log_prob_grad(th)
    if not np.isfinite(la):
        assert not np.isfinite(lb)
        return
    # #1 is a `~` Poisson (constants dropped) without zero inflation and `target +=` with it; #3 likewise
    assert abs(la - lb) <= 1e-10 * max(1.0, abs(la))
    assert np.max(np.abs(ga - gb) / np.maximum(1.0, np.abs(ga))) < 1e-10
