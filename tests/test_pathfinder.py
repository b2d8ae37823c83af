"""Host logic of the batched multi-path Pathfinder (/root/reference/bin/splotch_vinit:137-182).

No Stan here (§8c), so the restatement is pinned by what the published algorithm defines: the compact BFGS inverse
Hessian equals the recursive BFGS updates, log q equals the multivariate-normal density, a Gaussian target is
recovered with ELBO = log Z, PSIS recovers a known Pareto tail, and on the reference's own c1 files (the CPU oracle
standing in for the GPU density) the draws are usable starting points for NUTS.  The GPU binding is `BatchTarget`.
"""
import json
import os

import numpy as np
import pytest
from scipy import stats

from csplotch_b200 import pathfinder as pf

HERE = os.path.dirname(os.path.abspath(__file__))


class GaussTarget:
    """G independent D-variate normals N(mean_g, P_g^-1)."""

    def __init__(self, G, D, seed=0, gene_id0=0, cond=1.0):
        rng = np.random.default_rng(seed)
        self.G, self.dim, self.gene_id0 = G, D, gene_id0
        self.mean = rng.normal(size=(G, D))
        A = rng.normal(size=(G, D, D)) / np.sqrt(D)
        self.P = cond * np.einsum("gij,gkj->gik", A, A) + 0.5 * np.eye(D)

    def sub(self, g):
        out = GaussTarget(1, self.dim, gene_id0=self.gene_id0 + g)
        out.mean, out.P = self.mean[g:g + 1], self.P[g:g + 1]
        return out

    def lp_grad(self, th, go):
        d = th - self.mean[go]
        Pd = np.einsum("rij,rj->ri", self.P[go], d)
        return -0.5 * np.einsum("ri,ri->r", d, Pd), -Pd

    def lp(self, th, go):
        return self.lp_grad(th, go)[0]


def _sigma_of(ap, r):
    """Dense Sigma of row r of an _Approx from its sampling factors."""
    D = ap.D
    Th = np.eye(D)
    for rows, Q, T in ap.groups:
        hit = np.nonzero(rows == r)[0]
        if len(hit):
            q, t = Q[hit[0]], T[hit[0]]
            Th = np.eye(D) + q @ t @ q.T
    Th = ap.sqrt_alpha[r][:, None] * Th
    return Th @ Th.T


def test_compact_inverse_hessian_equals_recursive_bfgs():
    rng = np.random.default_rng(3)
    D, J = 9, 5
    for m in (0, 1, 3, 5):
        x, g = rng.normal(size=(2, D)), rng.normal(size=(2, D))
        alpha = rng.uniform(0.3, 2.0, size=(2, D))
        Hs = rng.normal(size=(2, D, D))
        Hs = np.einsum("rij,rkj->rik", Hs, Hs) + np.eye(D)           # a positive definite "true" Hessian: y = H s
        S = np.zeros((2, J, D))
        Y = np.zeros((2, J, D))
        S[:, :m] = rng.normal(size=(2, m, D))
        Y[:, :m] = np.einsum("rij,rmj->rmi", Hs, S[:, :m])
        ap = pf._Approx(x, g, alpha, S, Y, np.array([m, m]))
        assert ap.ok.all()
        for r in range(2):
            H = np.diag(alpha[r])
            for k in range(m):                                        # BFGS update of the inverse Hessian, oldest first
                s, y = S[r, k], Y[r, k]
                rho = 1.0 / (s @ y)
                V = np.eye(D) - rho * np.outer(s, y)
                H = V @ H @ V.T + rho * np.outer(s, s)
            Sig = _sigma_of(ap, r)
            assert np.abs(Sig - H).max() < 1e-10 * np.abs(H).max()
            assert np.abs(ap.mu[r] - (x[r] - H @ g[r])).max() < 1e-10
            assert abs(ap.logdet[r] - np.linalg.slogdet(H)[1]) < 1e-10
            u = rng.normal(size=(6, D))
            phi, logq = ap.take(np.array([r])).draw(u)
            ref = stats.multivariate_normal(ap.mu[r], H).logpdf(phi[0])
            assert np.abs(logq[0] - ref).max() < 1e-9


def test_two_loop_is_the_bfgs_inverse_hessian_with_scaled_identity():
    rng = np.random.default_rng(4)
    D, J = 7, 5
    g = rng.normal(size=(3, D))
    S = np.zeros((3, J, D))
    Y = np.zeros((3, J, D))
    m = np.array([0, 2, 5])
    for r in range(3):
        Hh = rng.normal(size=(D, D))
        Hh = Hh @ Hh.T + np.eye(D)
        S[r, :m[r]] = rng.normal(size=(m[r], D))
        Y[r, :m[r]] = S[r, :m[r]] @ Hh
    d = pf._two_loop(g, S, Y, m)
    for r in range(3):
        if m[r] == 0:
            H = np.eye(D)
        else:
            s, y = S[r, m[r] - 1], Y[r, m[r] - 1]
            H = (s @ y) / (y @ y) * np.eye(D)
        for k in range(m[r]):
            s, y = S[r, k], Y[r, k]
            rho = 1.0 / (s @ y)
            V = np.eye(D) - rho * np.outer(s, y)
            H = V @ H @ V.T + rho * np.outer(s, s)
        assert np.abs(d[r] + H @ g[r]).max() < 1e-10 * max(1.0, np.abs(H @ g[r]).max())


def test_history_ring_keeps_the_newest_pairs_in_order():
    S = np.zeros((2, 3, 2))
    Y = np.zeros((2, 3, 2))
    m = np.zeros(2, dtype=np.int64)
    for k in range(1, 6):
        rows = np.array([0]) if k % 2 else np.array([0, 1])
        pf._push(S, Y, m, rows, np.full((len(rows), 2), float(k)), np.full((len(rows), 2), -float(k)))
    assert m.tolist() == [3, 2]
    assert S[0, :, 0].tolist() == [3.0, 4.0, 5.0] and Y[0, :, 0].tolist() == [-3.0, -4.0, -5.0]
    assert S[1, :, 0].tolist() == [2.0, 4.0, 0.0]


def test_gaussian_targets_are_recovered():
    """D <= history size: BFGS learns the exact covariance, so ELBO -> log Z and the draws have the target's moments."""
    t = GaussTarget(3, 4, seed=1)
    r = pf.pathfinder(t, num_paths=4, num_draws=4000, seed=11)
    assert (r.info["return_code"] > 0).all() and (r.lbfgs_iters < 60).all()
    for g in range(3):
        cov = np.linalg.inv(t.P[g])
        logZ = 0.5 * np.linalg.slogdet(2 * np.pi * cov)[1]
        assert abs(r.elbo[g].max() - logZ) < 0.25                      # 25-draw estimates, best of all iterates
        assert np.abs(r.draws[g].mean(0) - t.mean[g]).max() < 0.12
        assert np.abs(np.cov(r.draws[g].T) - cov).max() < 0.15 * np.abs(cov).max()
        assert r.pareto_k[g] < 0.7
        lp = t.lp(r.draws[g], np.full(4000, g))
        assert np.abs(lp - r.lp[g]).max() < 1e-12                      # lp__ belongs to the returned draw
        assert set(np.unique(r.path[g])) <= {1, 2, 3, 4}


def test_result_of_a_gene_does_not_depend_on_the_batch():
    t = GaussTarget(3, 6, seed=2, gene_id0=40)
    full = pf.pathfinder(t, num_paths=2, num_draws=50, seed=9)
    one = pf.pathfinder(t.sub(2), num_paths=2, num_draws=50, seed=9)
    assert np.array_equal(full.draws[2], one.draws[0]) and np.array_equal(full.lp[2], one.lp[0])
    assert np.array_equal(full.lbfgs_iters[2], one.lbfgs_iters[0])
    other = pf.pathfinder(t.sub(2), num_paths=2, num_draws=50, seed=10)
    assert not np.array_equal(other.draws[0], one.draws[0])


def test_single_path_returns_its_draws_unweighted():
    t = GaussTarget(1, 5, seed=4)
    r = pf.pathfinder(t, num_paths=1, num_draws=300, seed=1)
    assert (r.path == 1).all() and not np.isfinite(r.pareto_k[0])            # no importance resampling happened
    assert len(np.unique(r.lp[0])) == 300                                       # every draw once, none duplicated
    r2 = pf.pathfinder(t, num_paths=2, num_draws=300, seed=1, psis_resample=False)
    assert sorted(np.unique(r2.path[0])) == [1, 2] and len(np.unique(r2.lp[0])) == 300


def test_failed_paths_and_non_finite_densities():
    class Hostile(GaussTarget):
        def lp_grad(self, th, go):
            lp, g = super().lp_grad(th, go)
            bad = th[:, 0] > 1.0                                         # outside the support: Stan would throw
            lp[bad], g[bad] = np.nan, np.nan
            dead = go == 1                                               # a gene whose density never evaluates
            lp[dead] = -np.inf
            return lp, g

    t = Hostile(2, 3, seed=5)
    t.mean[0, 0] = -1.0
    r = pf.pathfinder(t, num_paths=3, num_draws=100, seed=2)
    assert np.isfinite(r.lp[0]).all() and (r.draws[0][:, 0] <= 1.0).all()
    assert not np.isfinite(r.lp[1]).any() and (r.elbo[1] == -np.inf).all()


def test_psis_recovers_a_pareto_tail_and_normalises():
    rng = np.random.default_rng(0)
    for k_true in (0.2, 0.6):
        ratios = (1 - rng.uniform(size=20000)) ** (-k_true)              # Pareto tail with shape k
        logw, k = pf.psis_log_weights(np.log(ratios))
        assert abs(k - k_true) < 0.12
        assert abs(np.exp(logw).sum() - 1) < 1e-12
        assert logw.max() <= np.log(ratios.max() / ratios.sum()) + 1e-12  # smoothing never raises the largest weight
    logw, k = pf.psis_log_weights(np.zeros(100))
    assert np.allclose(logw, -np.log(100))
    logw, k = pf.psis_log_weights(np.array([0.0, -np.inf, np.nan, 1.0]))
    assert np.isfinite(logw[[0, 3]]).all() and abs(np.exp(logw[[0, 3]]).sum() - 1) < 1e-12


def test_gpdfit_on_generalised_pareto_samples():
    rng = np.random.default_rng(1)
    for k_true, sigma in ((0.1, 1.0), (0.5, 2.0)):
        x = np.sort(stats.genpareto(k_true, scale=sigma).rvs(5000, random_state=rng))
        k, s = pf._gpdfit(x)
        assert abs(k - k_true) < 0.06 and abs(s - sigma) < 0.15 * sigma


# ------------------------------------------------------------------------------------------- cSplotch target
class OracleTarget:
    """The CPU oracle as the density of the reference's c1 genes (tests only; the product uses BatchTarget)."""

    def __init__(self, gene_dicts, family, gene_id0=1):
        import oracle
        self.o = [oracle.Oracle(d, family) for d in gene_dicts]
        self.G, self.dim, self.gene_id0 = len(self.o), self.o[0].dim, gene_id0

    def lp_grad(self, th, go):
        lp, g = np.empty(len(th)), np.empty_like(th)
        for i in range(len(th)):
            lp[i], g[i] = self.o[go[i]].log_prob_grad(th[i])
        return lp, g

    def lp(self, th, go):
        return np.array([self.o[go[i]].log_prob_grad(th[i], want_grad=False) for i in range(len(th))])


@pytest.fixture(scope="module")
def c1_target():
    from csplotch_b200 import rdump
    ds = [rdump.read_rdump(os.path.join(HERE, "golden", "c1", "data_%d.R" % g)) for g in (1, 2)]
    return OracleTarget(ds, "comp_splotch_stan_model")


def test_c1_genes_give_usable_starting_points(c1_target):
    """bin/splotch_vinit on the reference's example: Pathfinder -> NUM_CHAINS draws -> `sample num_warmup=150 init=`."""
    import oracle
    t = c1_target
    r = pf.pathfinder(t, num_paths=4, num_draws=200, seed=3, max_lbfgs_iters=300)
    assert np.isfinite(r.elbo).any(axis=1).all() and np.isfinite(r.draws).all()
    rng = np.random.default_rng(0)
    for g in range(t.G):
        lp_rand = t.lp(rng.uniform(-2, 2, (50, t.dim)), np.full(50, g))
        assert np.median(r.lp[g]) > np.max(lp_rand)                      # far inside the typical set vs uniform(-2, 2)
        assert np.abs(t.lp(r.draws[g], np.full(200, g)) - r.lp[g]).max() < 1e-9
        # PSIS weights of a 94-dimensional hierarchical posterior are heavy-tailed; k is reported, not asserted
        assert np.isfinite(r.lp_approx[g]).all()
    # the sampler accepts the draws as explicit starting points (what init=FILE.json becomes)
    cfg = oracle.nuts_cfg(150, 50, seed=4)
    pick = np.random.default_rng(1).choice(200, size=2, replace=False)       # bin/splotch_vinit:156
    for c, i in enumerate(pick):
        out = t.o[0].nuts(cfg, gene_id=1, chain_id=c + 1, q_init=r.draws[0, i])
        assert out["status"] == 0 and np.isfinite(out["draws"]).all()


def test_pathfinder_csv_feeds_the_reference_init_extractor(tmp_path, c1_target):
    """The CSV the shim's `pathfinder` method writes, read by the reference's own extraction logic
    (bin/splotch_vinit:146-182, restated): header -> JSON of constrained values -> `unconstrain` -> the draw."""
    from types import SimpleNamespace as NS
    from csplotch_b200 import api, stancsv
    t = c1_target
    o = t.o[0]
    r = pf.pathfinder(t, num_paths=2, num_draws=20, seed=8, max_lbfgs_iters=100)
    vals = np.stack([o.write_array(th) for th in r.draws[0]])
    d = o._keep
    model = NS(levels=(1, 0, 0), N=o.N, C=14, K=o.K, car=1, zi=1, nb=0, family=1, dim=o.dim)
    names = [n for n, _, _ in sorted(api.param_layout(model), key=lambda e: e[1])]
    # columns: parameters in declaration order, then transformed parameters (a11); only the names matter here
    cols = _c1_columns(model, vals.shape[1])
    lead = np.stack([r.lp_approx[0], r.lp[0], r.path[0].astype(float)], axis=1)
    p = str(tmp_path / "output_1.csv")
    stancsv.write_combined_csv(p, cols, lead, vals, lead_cols=stancsv.PATHFINDER_COLS)
    with open(p) as fh:
        lines = [ln for ln in fh if not ln.startswith("#")]
    header = lines[0].strip().split(",")
    assert header[:3] == ["lp_approx__", "lp__", "path__"] and len(lines) == 21
    values = lines[1 + 7].strip().split(",")
    init = {k: float(v) for k, v in zip(header, values) if k not in ["lp_approx__", "lp__", "path__"]}
    json.dump(init, open(str(tmp_path / "init_1_1.json"), "w"), indent=2)
    th = api.unconstrain(model, json.load(open(str(tmp_path / "init_1_1.json"))))
    assert set(names) <= set(init)
    assert np.abs(th - r.draws[0, 7]).max() < 1e-9 * max(1.0, np.abs(r.draws[0, 7]).max())


def _c1_columns(model, n_out):
    """Dotted column names of comp_splotch_stan_model's write_array for the c1 design."""
    N, C, K = model.N, model.C, model.K
    cols = ["psi.%d" % (i + 1) for i in range(N)]
    cols += ["beta_raw.%d.%d.%d" % (1, j + 1, k + 1) for k in range(K) for j in range(C)]
    cols += ["tau.1", "a.1", "sigma", "theta.1"]
    cols += ["noise_raw.%d" % (i + 1) for i in range(N)]
    cols += ["log_lambda.%d" % (i + 1) for i in range(N)]
    cols += ["beta_level_1.%d.%d.%d" % (1, j + 1, k + 1) for k in range(K) for j in range(C)]
    assert len(cols) == n_out, (len(cols), n_out)
    return cols


def test_vinit_chains_sample_the_same_posterior(c1_target):
    """bin/splotch_vinit end to end on the CPU oracle: Pathfinder starts + 150 warm-up transitions give the same
    posterior as the plain run (N warm-up transitions from uniform(-2, 2)) — means within 5 Monte Carlo standard
    errors for every beta_level_1 entry and scalar of the reference's c1 gene."""
    import oracle
    from csplotch_b200 import diagnostics as dg
    t, o = c1_target, c1_target.o[0]
    n, chains = 400, 4
    r = pf.pathfinder(t, num_paths=chains, num_draws=1000, num_out=chains, seed=12, max_lbfgs_iters=200)
    runs = {}
    for name, nw, inits in (("plain", n, [None] * chains), ("vinit", 150, list(r.draws[0]))):
        cfg = oracle.nuts_cfg(nw, n, seed=31)
        outs = [o.nuts(cfg, gene_id=1, chain_id=c + 1, q_init=inits[c]) for c in range(chains)]
        assert all(x["status"] == 0 for x in outs)
        th = np.stack([x["draws"][:, 7:] for x in outs])                        # (chains, n, D) unconstrained
        wa = np.stack([[o.write_array(q) for q in ch] for ch in th])            # constrained, CSV column order
        runs[name] = wa[:, :, 10:10 + 70 + 4]                                    # beta_raw (70) | tau a sigma theta
    worst = 0.0
    for j in range(runs["plain"].shape[2]):
        a, b = runs["plain"][:, :, j], runs["vinit"][:, :, j]
        se = np.sqrt(a.var() / dg.ess_bulk(a) + b.var() / dg.ess_bulk(b))
        worst = max(worst, abs(a.mean() - b.mean()) / se)
    assert worst < 5.0, worst
