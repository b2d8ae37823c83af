"""One gene-chain per thread-block cluster (csrc/device_ops.cuh, DevOps<KT, CL = true>): the sampler and the trajectory
kernel of long designs split a chain's tiles over 2 / 4 CTAs (halo cells behind the psi ring, reductions through
distributed shared memory in rank order).  Same parity bars as the one-CTA path: fixed-step leapfrog trajectories 1e-8
against the oracle, the first adaptive NUTS transitions draw for draw, identical draws whatever the order of the genes.
CSPLOTCH_CLUSTER forces the cluster size at csplotch_model_create (designs this small default to 1)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
from csplotch_b200 import api, synth


def _rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))


def _design(kind):
    if kind == "comp_hex":       # Visium hex rows of 32, K = 10, 3 levels: the c3 kernel (KT = 10, LOOK = 1), 20 tiles
        return "comp_splotch_stan_model", synth.shared_c3(n_arrays=4, rows=40, cols=32, levels=(1, 2, 4))
    if kind == "grid_2level":    # 4-neighbour grids, 2 levels, Poisson-free ZIP + CAR: KT = 1, LOOK = 1, 15 tiles
        return "splotch_stan_model", synth.shared_c4(n_arrays=3, side=36, levels=(1, 3))
    if kind == "hd_wide":        # 300-wide rows: neighbours two tiles away (LOOK = 2), 14.1 -> 15 tiles
        rng = np.random.default_rng(2)
        h, w = 6, 300
        p = synth.grid4_pairs(h, w)
        n = h * w
        return "splotch_stan_model_csr", synth.build_shared(
            [n, n], [p, p], [synth.car_eigenvalues(p, n)] * 2, rng.integers(1, 4, size=2 * n), [1, 2], (1, 2), [1, 1], [],
            np.clip(rng.lognormal(0, 0.5, 2 * n), 0.05, 8), 3, zi=1, car=1)
    raise ValueError(kind)


@pytest.mark.parametrize("kind,cl", [("comp_hex", 2), ("comp_hex", 4), ("grid_2level", 2), ("grid_2level", 4), ("hd_wide", 2)])
def test_cluster_leapfrog_trajectory_matches_oracle(kind, cl, monkeypatch):
    monkeypatch.setenv("CSPLOTCH_CLUSTER", str(cl))
    family, shared = _design(kind)
    genes = synth.simulate_counts(shared, 3, 31)
    model = api.Model(shared, family)
    b = api.Batch(model, genes["counts"], genes.get("beta_prior_mean"), genes.get("beta_prior_std"))
    rng = np.random.default_rng(cl)
    B = 5                                          # more trajectories than genes: clusters loop over them
    gene_of = np.arange(B) % 3
    theta0 = rng.normal(0, 0.4, (B, model.dim))
    p0 = rng.normal(0, 1, (B, model.dim))
    mi = np.exp(rng.normal(0, 0.3, (B, model.dim)))
    for eps, n in ((1e-3, 16), (5e-3, 0)):
        th, p, H = b.leapfrog_fixed(theta0, p0, mi, eps, n, gene_of=gene_of)
        for i in range(B):
            o = oracle.Oracle(synth.gene_dict(shared, genes, int(gene_of[i])), family)
            th0, pp0, H0 = o.leapfrog_fixed(theta0[i], p0[i], mi[i], eps, n)
            assert _rel(th[i], th0) < 1e-8 and _rel(p[i], pp0) < 1e-8, (kind, cl, i)
            assert abs(H[i] - H0) < 1e-8 * max(1.0, abs(H0)), (kind, cl, i)


@pytest.mark.parametrize("kind,cl", [("comp_hex", 4), ("grid_2level", 2), ("hd_wide", 2)])
def test_cluster_nuts_transitions_match_oracle(kind, cl, monkeypatch):
    monkeypatch.setenv("CSPLOTCH_CLUSTER", str(cl))
    family, shared = _design(kind)
    genes = synth.simulate_counts(shared, 2, 32)
    model = api.Model(shared, family)
    b = api.Batch(model, genes["counts"], genes.get("beta_prior_mean"), genes.get("beta_prior_std"), gene_id0=7)
    nw, ns, nch = 12, 4, 2
    s = api.Sampler(b, api.nuts_cfg(nw, ns, nch, seed=5, keep_draws=True))
    ocfg = oracle.nuts_cfg(nw, ns, seed=5)
    want = {}
    for g in range(2):
        o = oracle.Oracle(synth.gene_dict(shared, genes, g), family)
        for c in range(nch):
            want[g, c] = o.nuts(ocfg, gene_id=7 + g, chain_id=c + 1, save_warmup=True)
    n_leap = 0
    for it in range(6):
        s.advance(1)
        th, eps, _ = s.state()
        for (g, c), w in want.items():
            assert np.max(np.abs(th[g, c] - w["draws"][it, 7:])) < 1e-7, (kind, cl, it, g, c)
            assert abs(eps[g, c] - w["draws"][it + 1, 2]) < 1e-7 * eps[g, c]
        n_leap += sum(int(w["draws"][it, 4]) for w in want.values())
        assert s.counters()["n_leapfrog"] == n_leap
    s.advance(nw + ns)
    st, it_done = s.status()
    assert (st == 0).all() and (it_done == nw + ns).all()
    diag, small, nd = s.draws()
    assert (nd == ns).all() and np.all(diag[..., 4] >= 1)
    # the streaming spot summaries (rank 0 of the cluster) equal the moments of the kept draws
    summ = s.spot_summaries()
    theta = s.theta_draws()
    cols = model.column_names()
    i0 = cols.index("log_lambda.1")
    wa = b.write_array(theta[0].reshape(-1, model.dim), gene_of=np.zeros(nch * ns, dtype=np.int32))
    ll = wa[:, i0:i0 + model.N]
    assert np.allclose(summ["log_lambda/mean"][0], ll.mean(0), rtol=1e-10, atol=1e-12)
    assert np.allclose(summ["lambda/std"][0], np.exp(ll).std(0), rtol=1e-8, atol=1e-12)
    assert np.allclose(summ["psi/mean"][0], wa[:, :model.N].mean(0), rtol=1e-10, atol=1e-12)


def test_cluster_draws_do_not_depend_on_gene_order(monkeypatch):
    """The RNG key travels with the gene in cluster launches too.  Compared after three transitions to 1e-9: on designs
    with many tiles the bottom-level beta-gradient bins are summed by shared-memory fp64 atomics whose order varies from
    launch to launch (last-bit differences, amplified ~3x per transition), so runs are reproducible to rounding, bit for
    bit only when no bin receives more than two warp contributions."""
    monkeypatch.setenv("CSPLOTCH_CLUSTER", "2")
    family, shared = _design("grid_2level")
    genes = synth.simulate_counts(shared, 4, 33)
    model = api.Model(shared, family)
    cfg = api.nuts_cfg(10, 4, 2, seed=3)
    full = api.Sampler(api.Batch(model, genes["counts"], gene_id0=50), cfg)
    assert full.layout()[1] == 2
    full.advance(2); full.advance(1)
    th_full, eps_full, _ = full.state()
    perm = np.array([3, 1, 0, 2])
    part = api.Sampler(api.Batch(model, genes["counts"][perm], gene_ids=50 + perm), cfg)
    part.advance(3)
    th_part, eps_part, _ = part.state()
    assert np.allclose(th_full[perm], th_part, rtol=1e-9, atol=1e-9)
    assert np.allclose(eps_full[perm], eps_part, rtol=1e-9)


def test_cluster_size_follows_the_number_of_chains():
    """choose_cluster: few gene-chains of a long design run on 4-CTA clusters, thousands on plain CTAs"""
    family, shared = "comp_splotch_stan_model", synth.shared_c3(n_arrays=9)     # 176 tiles: clusters of up to 4
    genes = synth.simulate_counts(shared, 2, 5, chunk=2)
    model = api.Model(shared, family)
    b = api.Batch(model, genes["counts"], genes["beta_prior_mean"], genes["beta_prior_std"])
    s = api.Sampler(b, api.nuts_cfg(2, 0, 4, seed=1))
    resident, cl = s.layout()
    assert cl == 4 and resident >= 8
    s.advance(1)
    st, it = s.status()
    assert (st == 0).all() and (it == 1).all()
    tiny_shared = synth.shared_tiny(seed=1)
    tm = api.Model(tiny_shared, family)
    tg = synth.simulate_counts(tiny_shared, 2, 6)
    ts = api.Sampler(api.Batch(tm, tg["counts"], tg["beta_prior_mean"], tg["beta_prior_std"]), api.nuts_cfg(2, 0, 2))
    assert ts.layout()[1] == 1
