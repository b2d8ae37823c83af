"""Full-size parity: the shapes bench.py times meet the oracle on the GPU.

  c3   24 x 4 992 Visium spots, comp_splotch_stan_model 3-level ZIP+CAR, K = 10 (468 tiles per evaluation, 4 992
       distinct eigenvalues x 24)                                  stan/comp_splotch_stan_model.stan:81-280
  c4   4 x 317^2 Visium-HD bins, splotch_stan_model_csr 2-level ZIP+CAR (1 571 tiles, neighbours two tiles away)
                                                                   stan/splotch_stan_model_csr.stan:4-25,254-271
  SNAP25  the whole Science/human/data_SNAP25.R (80 sections, N = 61 031, the file's own eigenvalues incl. 88 at
       exactly +1), real counts                                   Science/human/splotch_human.stan == family 0, 2 levels

Per shape: lp and grad at two points (1e-10 relative, fp64) and a 16-step fixed-step leapfrog trajectory (1e-8)
against the C oracle.  One oracle evaluation at N = 4e5 costs milliseconds."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
from csplotch_b200 import api, synth


def _rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))


def _check(shared, genes, family, seed, scale=0.7, eps=1e-3, steps=16):
    model = api.Model(shared, family)
    b = api.Batch(model, genes["counts"], genes.get("beta_prior_mean"), genes.get("beta_prior_std"), gene_id0=5)
    G = genes["counts"].shape[0]
    rng = np.random.default_rng(seed)
    theta = rng.normal(0, scale, (2, model.dim))
    theta[1] *= 1.5
    gene_of = np.arange(2) % G
    lp, grad = b.log_prob_grad(theta, gene_of=gene_of)
    for i in range(2):
        o = oracle.Oracle(synth.gene_dict(shared, genes, int(gene_of[i])), family)
        lp0, g0 = o.log_prob_grad(theta[i])
        assert abs(lp[i] - lp0) <= 1e-10 * max(1.0, abs(lp0)), (i, lp[i], lp0)
        assert _rel(grad[i], g0) < 1e-10, i
        if i == 0:
            p0 = rng.normal(0, 1, model.dim)
            mi = np.exp(rng.normal(0, 0.3, model.dim))
            th, pp, H = b.leapfrog_fixed(theta[i] * 0.5, p0, mi, eps, steps, gene_of=[int(gene_of[i])])
            th0, pp0, H0 = o.leapfrog_fixed(theta[i] * 0.5, p0, mi, eps, steps)
            assert _rel(th[0], th0) < 1e-8 and _rel(pp[0], pp0) < 1e-8
            assert abs(H[0] - H0) < 1e-8 * max(1.0, abs(H0))
    return model, b


def test_c3_full_size():
    shared = synth.shared_c3()
    genes = synth.simulate_counts(shared, 2, 77, chunk=2)
    model, b = _check(shared, genes, "comp_splotch_stan_model", seed=1)
    assert model.N == 24 * 4992 and model.dim == 242142


def test_c4_full_size():
    shared = synth.shared_c4()
    genes = synth.simulate_counts(shared, 2, 78, chunk=2)
    model, b = _check(shared, genes, "splotch_stan_model_csr", seed=2)
    assert model.N == 4 * 317 * 317 and model.dim == 803989


def _snap25_full(golden_dir):
    z = np.load(os.path.join(golden_dir, "snap25_full.npz"))
    ns = z["N_spots"].astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(ns)])
    W0 = z["W_sparse"].astype(np.int64) - 1
    t_of = np.searchsorted(offs, W0[:, 0], side="right") - 1
    pairs = [W0[t_of == t] - offs[t] for t in range(len(ns))]
    # the file's own eigenvalues, as one sorted list (build_shared only concatenates and sorts)
    shared = synth.build_shared(ns, pairs, [z["eig_values"]], z["D"].astype(np.int64), z["tissue_mapping"],
                                (int(z["N_onsets_tissue_locations"]), int(z["N_donors"])), z["donor_mapping"], [],
                                z["size_factors"], int(z["N_covariates"]), zi=1, car=1)
    assert np.array_equal(shared["D_sparse"], z["D_sparse"])
    assert int(np.sum(shared["eig_values"] == 1.0)) == 88
    return shared, {"counts": z["counts"][None, :].astype(np.int32)}


@pytest.mark.parametrize("family", ["splotch_stan_model", "splotch_stan_model_csr"])
def test_snap25_full_size(golden_dir, family):
    shared, genes = _snap25_full(golden_dir)
    model, b = _check(shared, genes, family, seed=3)
    assert model.N == 61031
    # a = 0.9999 with eigenvalues exactly +1: log1m(a lambda) stays finite and matches
    o = oracle.Oracle(synth.gene_dict(shared, genes, 0), family)
    th = np.random.default_rng(4).normal(0, 0.3, model.dim)
    th[model.N + model.n_beta + 1] = np.log(0.9999 / 0.0001)
    lp, grad = b.log_prob_grad(th[None], gene_of=[0])
    lp0, g0 = o.log_prob_grad(th)
    assert abs(lp[0] - lp0) <= 1e-10 * abs(lp0) and _rel(grad[0], g0) < 1e-10
