"""Oracle self-consistency (SURVEY.md App. C tier 1, CPU part): the C restatement vs an
independent torch-autograd transcription of the .stan files, finite differences, mpmath,
and agreement of the three model code paths."""
import itertools

import numpy as np
import pytest

import oracle
from csplotch_b200 import synth
from tests.stan_transcript import stan_log_prob_grad

FAMS = ["splotch_stan_model", "comp_splotch_stan_model", "splotch_stan_model_csr"]


def _case(family, n_levels, zi, car, seed):
    shared = synth.shared_tiny(seed=seed, family=family, n_levels=n_levels, zi=zi, car=car)
    genes = synth.simulate_counts(shared, 2, seed + 1)
    if family == "comp_splotch_stan_model":
        rng = np.random.default_rng(seed + 2)
        genes["beta_prior_mean"] = rng.normal(0, 1, genes["beta_prior_mean"].shape)
        genes["beta_prior_std"] = rng.uniform(0.5, 3, genes["beta_prior_std"].shape)
    return shared, genes


@pytest.mark.parametrize("family,n_levels,zi,car",
                         list(itertools.product(FAMS, [1, 2, 3], [0, 1], [0, 1])))
def test_oracle_matches_autograd_transcription(family, n_levels, zi, car):
    shared, genes = _case(family, n_levels, zi, car, seed=11 * n_levels + 3 * zi + car)
    d = synth.gene_dict(shared, genes, 0)
    o = oracle.Oracle(d, family)
    rng = np.random.default_rng(5)
    for scale in (1.0, 3.0):
        th = rng.normal(0, scale, o.dim)
        for jac in (True, False):
            lp, g = o.log_prob_grad(th, jacobian=jac)
            lp2, g2 = stan_log_prob_grad(d, family, th, jacobian=jac)
            if not (np.isfinite(lp2) and np.all(np.isfinite(g2))):
                # exp overflow at an N(0,3^2) draw: both sides must agree the point is unusable
                assert not (np.isfinite(lp) and np.all(np.isfinite(g)))
                continue
            assert abs(lp - lp2) <= 1e-11 * max(1.0, abs(lp2))
            assert np.max(np.abs(g - g2) / np.maximum(1.0, np.abs(g2))) < 1e-11


@pytest.mark.parametrize("family", FAMS)
def test_oracle_gradient_finite_differences(family):
    shared, genes = _case(family, 3, 1, 1, seed=7)
    d = synth.gene_dict(shared, genes, 1)
    o = oracle.Oracle(d, family)
    rng = np.random.default_rng(1)
    th = rng.normal(0, 1, o.dim)
    lp, g = o.log_prob_grad(th)
    h = 1e-6
    for i in rng.choice(o.dim, size=min(o.dim, 40), replace=False):
        e = np.zeros(o.dim); e[i] = h
        fd = (o.log_prob_grad(th + e, want_grad=False) - o.log_prob_grad(th - e, want_grad=False)) / (2 * h)
        assert abs(fd - g[i]) < 2e-6 * max(1.0, abs(g[i])), (i, fd, g[i])


def test_three_code_paths_agree():
    """#1 == #3 (edge loop vs CSR, per-spot ZIP vs zero/non-zero split); #2 with K=1, E=1,
    prior (0,2) == #1."""
    shared, genes = _case("splotch_stan_model", 3, 1, 1, seed=21)
    d = synth.gene_dict(shared, genes, 0)
    o1 = oracle.Oracle(d, "splotch_stan_model")
    o3 = oracle.Oracle(d, "splotch_stan_model_csr")
    d2 = dict(d)
    N = int(np.sum(d["N_spots"]))
    d2["E"] = np.ones((N, 1)); d2["N_celltypes"] = 1
    d2["beta_prior_mean"] = np.zeros(1); d2["beta_prior_std"] = np.full(1, 2.0)
    o2 = oracle.Oracle(d2, "comp_splotch_stan_model")
    assert o1.dim == o3.dim == o2.dim
    rng = np.random.default_rng(3)
    L, C = d["N_level_1"] + d["N_level_2"] + d["N_level_3"], d["N_covariates"]
    for _ in range(3):
        th = rng.normal(0, 1.5, o1.dim)
        lp1, g1 = o1.log_prob_grad(th)
        lp3, g3 = o3.log_prob_grad(th)
        assert abs(lp1 - lp3) < 1e-12 * abs(lp1)
        assert np.allclose(g1, g3, rtol=1e-12, atol=1e-12)
        # family 1 stores beta_raw (i*C+j), families 0/2 column-major i+L*j
        th2 = th.copy()
        blk = th[N:N + L * C].reshape(C, L).T          # [i, j]
        th2[N:N + L * C] = blk.reshape(-1)
        lp2, g2 = o2.log_prob_grad(th2)
        assert abs(lp1 - lp2) < 1e-12 * abs(lp1)
        g2b = g2.copy()
        g2b[N:N + L * C] = g2[N:N + L * C].reshape(L, C).T.reshape(-1)
        assert np.allclose(g1, g2b, rtol=1e-12, atol=1e-12)


def test_poisson_glm_limit():
    """car=0, zi=0, one level: lp = sum(y*eta - exp(eta)) - 0.5|raw|^2 - 0.5|eps|^2 - 0.5 sigma^2 + log-Jacobian."""
    shared, genes = _case("splotch_stan_model", 1, 0, 0, seed=4)
    d = synth.gene_dict(shared, genes, 0)
    o = oracle.Oracle(d, "splotch_stan_model")
    rng = np.random.default_rng(0)
    th = rng.normal(0, 1, o.dim)
    L, C, N = d["N_level_1"], d["N_covariates"], int(np.sum(d["N_spots"]))
    raw = th[:L * C].reshape(C, L).T
    u_sigma = th[L * C]; eps = th[L * C + 1:]
    row = np.repeat(np.asarray(d["tissue_mapping"]) - 1, d["N_spots"])
    eta = 2 * raw[row, np.asarray(d["D"]) - 1] + 0.3 * np.exp(u_sigma) * eps + np.log(d["size_factors"])
    y = d["counts"]
    want = np.sum(y * eta - np.exp(eta)) - 0.5 * np.sum(raw ** 2) - 0.5 * np.sum(eps ** 2) \
        - 0.5 * np.exp(2 * u_sigma) + u_sigma
    assert abs(o.log_prob_grad(th, want_grad=False) - want) < 1e-10 * abs(want)


def test_oracle_vs_mpmath_50_digits():
    import mpmath as mp
    mp.mp.dps = 50
    shared, genes = _case("comp_splotch_stan_model", 2, 1, 1, seed=9)
    d = synth.gene_dict(shared, genes, 0)
    o = oracle.Oracle(d, "comp_splotch_stan_model")
    rng = np.random.default_rng(2)
    th = rng.normal(0, 2.0, o.dim)
    N = int(np.sum(d["N_spots"])); C = d["N_covariates"]; K = d["N_celltypes"]
    L1, L2 = d["N_level_1"], d["N_level_2"]; L = L1 + L2
    t = [mp.mpf(float(x)) for x in th]
    psi = t[:N]; raw = t[N:N + L * C * K]; pos = N + L * C * K
    tau = mp.e ** t[pos]; a = 1 / (1 + mp.e ** (-t[pos + 1])); sigma = mp.e ** t[pos + 2]
    s2 = mp.e ** t[pos + 3]; thz = 1 / (1 + mp.e ** (-t[pos + 4])); eps = t[pos + 5:]
    lp = t[pos] + mp.log(a) + mp.log(1 - a) + t[pos + 2] + t[pos + 3] + mp.log(thz) + mp.log(1 - thz)
    R = lambda i, j, k: raw[(i * C + j) * K + k]
    pm = [mp.mpf(float(x)) for x in d["beta_prior_mean"]]; ps = [mp.mpf(float(x)) for x in d["beta_prior_std"]]
    b1 = lambda i, j, k: R(i, j, k) * ps[k] + pm[k]
    b2 = lambda i, j, k: b1(int(d["level_2_mapping"][i]) - 1, j, k) + s2 * R(L1 + i, j, k)
    row = np.repeat(np.asarray(d["tissue_mapping"]) - 1, d["N_spots"])
    E = np.asarray(d["E"])
    lp += -2 * mp.log(tau) - 1 / tau
    W = np.asarray(d["W_sparse"]) - 1
    quad = sum(psi[i] * mp.mpf(float(d["D_sparse"][i])) * psi[i] for i in range(N))
    quad -= a * 2 * sum(psi[i] * psi[j] for i, j in W)
    ldet = sum(mp.log(1 - a * mp.mpf(float(l))) for l in d["eig_values"])
    lp += (N * mp.log(tau) + ldet - tau * quad) / 2
    lp += mp.log(1 - thz) - sigma ** 2 / 2 - s2 ** 2 / 2 - sum(x * x for x in eps) / 2 - sum(x * x for x in raw) / 2
    for j in range(N):
        ll = sum(b2(int(row[j]), int(d["D"][j]) - 1, k) * mp.mpf(float(E[j, k])) for k in range(K))
        eta = ll + psi[j] + mp.mpf("0.3") * sigma * eps[j] + mp.log(mp.mpf(float(d["size_factors"][j])))
        # 0.3 in the Stan source is the double 0.3; use the same double value
        eta = ll + psi[j] + mp.mpf(0.3) * sigma * eps[j] + mp.log(mp.mpf(float(d["size_factors"][j])))
        y = int(d["counts"][j])
        if y == 0:
            lp += mp.log(thz + (1 - thz) * mp.e ** (-mp.e ** eta))
        else:
            lp += mp.log(1 - thz) + y * eta - mp.e ** eta - mp.loggamma(y + 1)
    got = o.log_prob_grad(th, want_grad=False)
    assert abs(got - float(lp)) < 1e-12 * abs(float(lp))


@pytest.mark.parametrize("zi,nb", [(0, 1), (1, 1), (1, 2)])
def test_negative_binomial_branch(zi, nb):
    """NB / ZINB likelihood of comp_splotch_stan_model.stan:248-265 incl. the phi_arr broadcast of the ZINB
    branch as written (nb=1) and the per-spot variant (nb=2)."""
    shared, genes = _case("comp_splotch_stan_model", 2, zi, 1, seed=31 + zi)
    shared = dict(shared); shared["nb"] = nb
    d = synth.gene_dict(shared, genes, 0)
    o = oracle.Oracle(d, "comp_splotch_stan_model")
    N = int(np.sum(d["N_spots"]))
    assert o.dim == 2 * N + o.dim - 2 * N and o.n_out == o.dim + 2 * N + (o.dim - 2 * N - (4 + zi + 1))
    rng = np.random.default_rng(3)
    for scale in (0.5, 1.5):
        th = rng.normal(0, scale, o.dim)
        for jac in (True, False):
            lp, g = o.log_prob_grad(th, jacobian=jac)
            lp2, g2 = stan_log_prob_grad(d, "comp_splotch_stan_model", th, jacobian=jac)
            assert abs(lp - lp2) <= 1e-11 * max(1.0, abs(lp2))
            assert np.max(np.abs(g - g2) / np.maximum(1.0, np.abs(g2))) < 1e-10
    wa = o.write_array(th)
    assert np.allclose(o.unconstrain(wa[:o.dim]), th, rtol=1e-12, atol=1e-12)
    phi = 1e-6 + np.exp(th[N + (o.dim - 2 * N - (4 + zi + 1)) + 4 + zi])
    assert np.allclose(wa[o.dim + N:o.dim + 2 * N], phi)           # phi_arr follows log_lambda


def test_oracle_reproduces_committed_known_answers():
    """tests/golden/c1_lpgrad.npz (made by tests/golden/make_lpgrad_golden.py; oracle output accepted only where the
    autograd transcription agreed): a change of the oracle that moves lp or the gradient on the reference's c1 genes
    fails here, and the GPU path is held to the same file (tests/test_gpu_parity.py)."""
    import os
    from csplotch_b200 import rdump
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z = np.load(os.path.join(gdir, "c1_lpgrad.npz"))
    assert z["theta"].shape == (7, 3, 94)
    for g in range(7):
        o = oracle.Oracle(rdump.read_rdump(os.path.join(gdir, "c1", "data_%d.R" % (g + 1))), "comp_splotch_stan_model")
        for i in range(3):
            lp, grad = o.log_prob_grad(z["theta"][g, i])
            assert abs(lp - z["lp_jacobian"][g, i]) <= 1e-13 * max(1.0, abs(lp))
            assert np.max(np.abs(grad - z["grad_jacobian"][g, i]) / np.maximum(1.0, np.abs(grad))) < 1e-13
            lp0 = o.log_prob_grad(z["theta"][g, i], jacobian=False, want_grad=False)
            assert abs(lp0 - z["lp_no_jacobian"][g, i]) <= 1e-13 * max(1.0, abs(lp0))
