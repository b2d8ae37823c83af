"""One gene-chain per WARP (csrc/warp_ops.cuh): the path of small designs with many chains (the reference's ST examples,
c2).  CSPLOTCH_SCOPE=warp at csplotch_model_create forces it for any number of chains.  Same bars as the CTA path:
energies and fixed-step trajectories against the oracle (1e-10 / 1e-8), adaptive NUTS transitions draw for draw, streaming
summaries equal to the moments of the draws — and, because a warp is alone on its reduction tables (no atomics), draws
that are BITWISE independent of batch composition and queue order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
from csplotch_b200 import api, synth


def _rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))


def _design(kind):
    if kind == "c2":             # the ST-v1 example geometry, 2-level Poisson + CAR, K = 1 (N = 2 594)
        return "splotch_stan_model", synth.shared_c2()
    if kind == "comp":           # compositional, K = 10, 3 levels, ZIP + CAR, hex arrays of 20 x 16
        return "comp_splotch_stan_model", synth.shared_c3(n_arrays=3, rows=20, cols=16, levels=(1, 2, 3), C=4)
    if kind == "csr_nocar":      # no CAR, ZIP, 1 level
        return "splotch_stan_model_csr", synth.shared_tiny(seed=4, family="splotch_stan_model_csr", n_levels=1, zi=1, car=0)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["c2", "comp", "csr_nocar"])
def test_warp_energy_and_trajectory_match_oracle(kind, monkeypatch):
    monkeypatch.setenv("CSPLOTCH_SCOPE", "warp")
    family, shared = _design(kind)
    genes = synth.simulate_counts(shared, 3, 41)
    model = api.Model(shared, family)
    b = api.Batch(model, genes["counts"], genes.get("beta_prior_mean"), genes.get("beta_prior_std"))
    rng = np.random.default_rng(5)
    B = 11                                         # more trajectories than warps of one CTA
    gene_of = np.arange(B) % 3
    theta0 = rng.normal(0, 0.5, (B, model.dim))
    p0 = rng.normal(0, 1, (B, model.dim))
    mi = np.exp(rng.normal(0, 0.3, (B, model.dim)))
    for eps, n in ((1e-3, 16), (5e-3, 0)):
        th, p, H = b.leapfrog_fixed(theta0, p0, mi, eps, n, gene_of=gene_of)
        for i in range(B):
            o = oracle.Oracle(synth.gene_dict(shared, genes, int(gene_of[i])), family)
            th0, pp0, H0 = o.leapfrog_fixed(theta0[i], p0[i], mi[i], eps, n)
            assert _rel(th[i], th0) < 1e-8 and _rel(p[i], pp0) < 1e-8, (kind, i)
            assert abs(H[i] - H0) < 1e-10 * max(1.0, abs(H0)) if n == 0 else abs(H[i] - H0) < 1e-8 * max(1.0, abs(H0))


@pytest.mark.parametrize("kind", ["c2", "comp"])
def test_warp_nuts_transitions_match_oracle(kind, monkeypatch):
    monkeypatch.setenv("CSPLOTCH_SCOPE", "warp")
    family, shared = _design(kind)
    genes = synth.simulate_counts(shared, 2, 42)
    model = api.Model(shared, family)
    b = api.Batch(model, genes["counts"], genes.get("beta_prior_mean"), genes.get("beta_prior_std"), gene_id0=3)
    nw, ns, nch = 25, 6, 3
    s = api.Sampler(b, api.nuts_cfg(nw, ns, nch, seed=8, keep_draws=True))
    assert s.layout()[1] == 0                       # 0 = one gene-chain per warp
    ocfg = oracle.nuts_cfg(nw, ns, seed=8)
    want = {}
    for g in range(2):
        o = oracle.Oracle(synth.gene_dict(shared, genes, g), family)
        for c in range(nch):
            want[g, c] = o.nuts(ocfg, gene_id=3 + g, chain_id=c + 1, save_warmup=True)
    n_leap = 0
    for it in range(8):
        s.advance(1)
        th, eps, _ = s.state()
        for (g, c), w in want.items():
            assert np.max(np.abs(th[g, c] - w["draws"][it, 7:])) < 1e-7, (kind, it, g, c)
            assert abs(eps[g, c] - w["draws"][it + 1, 2]) < 1e-7 * eps[g, c]
        n_leap += sum(int(w["draws"][it, 4]) for w in want.values())
        assert s.counters()["n_leapfrog"] == n_leap
    s.advance(nw + ns)
    st, it_done = s.status()
    assert (st == 0).all() and (it_done == nw + ns).all()
    diag, small, nd = s.draws()
    assert (nd == ns).all() and np.all(diag[..., 4] >= 1)
    summ = s.spot_summaries()
    theta = s.theta_draws()
    cols = model.column_names()
    i0 = cols.index("log_lambda.1")
    wa = b.write_array(theta[1].reshape(-1, model.dim), gene_of=np.ones(nch * ns, dtype=np.int32))
    ll = wa[:, i0:i0 + model.N]
    assert np.allclose(summ["log_lambda/mean"][1], ll.mean(0), rtol=1e-10, atol=1e-12)
    assert np.allclose(summ["lambda/std"][1], np.exp(ll).std(0), rtol=1e-8, atol=1e-12)


def test_warp_draws_are_bitwise_independent_of_batch_and_order(monkeypatch):
    monkeypatch.setenv("CSPLOTCH_SCOPE", "warp")
    family, shared = _design("c2")
    genes = synth.simulate_counts(shared, 5, 43)
    model = api.Model(shared, family)
    cfg = api.nuts_cfg(12, 4, 2, seed=3)
    full = api.Sampler(api.Batch(model, genes["counts"], gene_id0=50), cfg).run(chunk=5)
    d_full, s_full, _ = full.draws()
    perm = np.array([3, 1, 4, 0])
    monkeypatch.setenv("CSPLOTCH_QUEUE_ORDER", "index")
    part = api.Sampler(api.Batch(model, genes["counts"][perm], gene_ids=50 + perm), cfg).run()
    d_part, s_part, _ = part.draws()
    assert np.array_equal(d_full[perm], d_part) and np.array_equal(s_full[perm], s_part)


def test_warp_scope_is_opt_in():
    family, shared = _design("c2")
    genes = synth.simulate_counts(shared, 2, 44)
    model = api.Model(shared, family)
    few = api.Sampler(api.Batch(model, genes["counts"]), api.nuts_cfg(2, 0, 4))
    assert few.layout()[1] == 1                     # default: a CTA per chain
