"""The product's sampler control logic (csplotch_b200/csrc/nuts_core.h: iterative tree with a
vector pool, reservoir proposal, windowed adaptation) run on the host against the oracle's
recursive restatement of Stan's base_nuts: same Philox stream => the same draws."""
import ctypes as C

import numpy as np
import pytest

import oracle
from csplotch_b200 import synth
from tests import emul


def _gauss_ctx(sd):
    arr = np.concatenate([[len(sd)], sd]).astype(np.float64)
    return arr, arr.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("nw,ns", [(0, 30), (10, 20), (60, 40), (150, 50), (200, 100)])
def test_gaussian_draw_for_draw(nw, ns):
    D = 7
    sd = np.linspace(0.3, 4.0, D)
    arr, ctx = _gauss_ctx(sd)
    cfg = oracle.nuts_cfg(nw, ns, seed=11)
    fn = C.cast(oracle.lib().orc_gauss_density, C.c_void_p)
    want = oracle.lib()
    draws_o = np.zeros((ns, 7 + D)); inv_o = np.empty(D); eps_o = C.c_double(0); ng = C.c_int64(0)
    st = want.orc_nuts_generic(fn, ctx, C.c_int32(D), C.byref(cfg), C.c_uint32(3), C.c_uint32(2), None,
                               C.c_int32(0), draws_o.ctypes.data_as(C.c_void_p),
                               inv_o.ctypes.data_as(C.c_void_p), C.byref(eps_o), C.byref(ng))
    assert st == 0
    got = emul.nuts_run(fn, ctx, D, cfg, gene_id=3, chain_id=2, chunk=7)
    assert got["status"] == 0
    assert np.array_equal(got["draws"][:, 3:6], draws_o[:, 3:6])          # treedepth, n_leapfrog, divergent
    assert np.allclose(got["draws"], draws_o, rtol=1e-9, atol=1e-9)
    assert np.allclose(got["inv_metric"], inv_o, rtol=1e-12)
    assert abs(got["stepsize"] - eps_o.value) < 1e-12 * eps_o.value


@pytest.mark.parametrize("family,n_levels,zi,car", [
    ("comp_splotch_stan_model", 3, 1, 1),
    ("splotch_stan_model", 2, 0, 1),
    ("splotch_stan_model_csr", 2, 1, 1),
    ("splotch_stan_model", 1, 1, 0),
])
def test_splotch_draw_for_draw(family, n_levels, zi, car):
    shared = synth.shared_tiny(seed=5, family=family, n_levels=n_levels, zi=zi, car=car)
    genes = synth.simulate_counts(shared, 1, 8)
    o = oracle.Oracle(synth.gene_dict(shared, genes, 0), family)
    cfg = oracle.nuts_cfg(40, 25, seed=4)
    want = o.nuts(cfg, gene_id=9, chain_id=1)
    fn = C.cast(oracle.lib().orc_splotch_density, C.c_void_p)
    got = emul.nuts_run(fn, C.cast(C.pointer(o.c), C.c_void_p), o.dim, cfg, gene_id=9, chain_id=1, chunk=13)
    assert want["status"] == 0 and got["status"] == 0
    assert np.array_equal(got["draws"][:, 3:6], want["draws"][:, 3:6])
    assert np.allclose(got["draws"], want["draws"], rtol=1e-8, atol=1e-8)
    assert abs(got["stepsize"] - want["stepsize"]) < 1e-10 * want["stepsize"]


def test_max_depth_and_divergence_paths():
    """tiny step size forces max_depth; a huge one forces divergences; both must agree with the oracle."""
    D = 5
    sd = np.array([1e-2, 1.0, 1.0, 1.0, 50.0])
    arr, ctx = _gauss_ctx(sd)
    fn = C.cast(oracle.lib().orc_gauss_density, C.c_void_p)
    for stepsize, md in ((1e-3, 6), (1.0, 10)):
        cfg = oracle.nuts_cfg(0, 40, seed=2, max_depth=md, stepsize=stepsize)
        draws_o = np.zeros((40, 7 + D)); ng = C.c_int64(0)
        q0 = np.full(D, 0.1)
        oracle.lib().orc_nuts_generic(fn, ctx, C.c_int32(D), C.byref(cfg), C.c_uint32(0), C.c_uint32(1),
                                      q0.ctypes.data_as(C.c_void_p), C.c_int32(0),
                                      draws_o.ctypes.data_as(C.c_void_p), None, None, C.byref(ng))
        got = emul.nuts_run(fn, ctx, D, cfg, gene_id=0, chain_id=1, q_init=q0)
        assert np.array_equal(got["draws"][:, 3:6], draws_o[:, 3:6])
        assert np.allclose(got["draws"], draws_o, rtol=1e-9, atol=1e-9)


from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=40, deadline=None, derandomize=True)
@given(D=st.integers(1, 12), log_sd=st.lists(st.floats(-2, 2), min_size=12, max_size=12),
       nw=st.sampled_from([0, 1, 5, 19, 20, 21, 33, 99, 149, 150, 151, 200]), seed=st.integers(0, 2**31 - 1),
       max_depth=st.sampled_from([1, 2, 5, 10]), delta=st.sampled_from([0.6, 0.8, 0.95]),
       chunk=st.integers(0, 17), gene=st.integers(0, 40000), chain=st.integers(1, 8))
def test_random_gaussians_windows_and_depths(D, log_sd, nw, seed, max_depth, delta, chunk, gene, chain):
    """Warm-up lengths around every branch of Stan's window arithmetic (< 20: step size only; < 150: the 15 % / 10 %
    rule; the stretched last window), tree depths from 1 to 10, any launch chunking: the iterative tree / vector pool /
    reservoir proposal of nuts_core.h and the oracle's recursive restatement give the same draws, step size and metric."""
    sd = 10.0 ** np.array(log_sd[:D])
    arr, ctx = _gauss_ctx(sd)
    ns = 12
    cfg = oracle.nuts_cfg(nw, ns, seed=seed, max_depth=max_depth, delta=delta)
    fn = C.cast(oracle.lib().orc_gauss_density, C.c_void_p)
    draws_o = np.zeros((ns, 7 + D)); inv_o = np.empty(D); eps_o = C.c_double(0); ng = C.c_int64(0)
    st_o = oracle.lib().orc_nuts_generic(fn, ctx, C.c_int32(D), C.byref(cfg), C.c_uint32(gene), C.c_uint32(chain), None,
                                         C.c_int32(0), draws_o.ctypes.data_as(C.c_void_p),
                                         inv_o.ctypes.data_as(C.c_void_p), C.byref(eps_o), C.byref(ng))
    got = emul.nuts_run(fn, ctx, D, cfg, gene_id=gene, chain_id=chain, chunk=chunk)
    assert got["status"] == st_o
    if st_o != 0:
        return
    assert np.array_equal(got["draws"][:, 3:6], draws_o[:, 3:6])
    assert np.allclose(got["draws"], draws_o, rtol=1e-8, atol=1e-8)
    assert np.allclose(got["inv_metric"], inv_o, rtol=1e-10)
    assert abs(got["stepsize"] - eps_o.value) <= 1e-10 * eps_o.value
    # work accounting: the product counts the gradient evaluations it performs — one at the current point per step-size
    # search, reused by every attempt — while the oracle, like upstream, re-evaluates it per attempt; so the oracle's
    # count can only be the larger one (the rate the product reports is not inflated)
    assert 0 <= ng.value - got["n_grad"] <= 8 + ng.value // 5


def test_two_merges_per_pass_equal_one_by_one_bit_for_bit():
    """The merge2 cascade step (nuts_core.h: a leaf that closes two levels above level 0 runs both U-turn merges in one
    pass, the inner rho sum never stored) makes the same decisions and leaves the same vectors as two single merges:
    identical draws, to the last bit, on targets whose trees reach depth 6-10."""
    D = 9
    sd = np.geomspace(0.05, 6.0, D)
    arr, ctx = _gauss_ctx(sd)
    fn = C.cast(oracle.lib().orc_gauss_density, C.c_void_p)
    for seed in (1, 2, 3):
        cfg = oracle.nuts_cfg(60, 40, seed=seed)
        a = emul.nuts_run(fn, ctx, D, cfg, gene_id=5, chain_id=1, chunk=11, merge2=True)
        b = emul.nuts_run(fn, ctx, D, cfg, gene_id=5, chain_id=1, chunk=11, merge2=False)
        assert a["status"] == 0 and b["status"] == 0
        assert a["draws"][:, 3].max() >= 3          # deep enough for a two-level cascade
        assert np.array_equal(a["draws"], b["draws"])
        assert a["n_leapfrog"] == b["n_leapfrog"] and a["stepsize"] == b["stepsize"]


def test_time_budget_bounds_a_cpu_sample():
    """bench.py's CPU legs: with a wall-clock budget every chain of orc_nuts_batch stops at its next leapfrog once the
    budget is spent (gradients done so far are counted), and the budget is cleared afterwards."""
    import time
    shared = synth.shared_tiny(seed=5, family="comp_splotch_stan_model", n_levels=3, zi=1, car=1)
    genes = synth.simulate_counts(shared, 2, 8)
    o = oracle.Oracle(synth.gene_dict(shared, genes, 0), "comp_splotch_stan_model")
    t0 = time.time()
    res = o.nuts_batch(genes["counts"], oracle.nuts_cfg(100000, 0, seed=1), n_chains=2, prior_mean=genes["beta_prior_mean"],
                       prior_std=genes["beta_prior_std"], n_threads=4, want_draws=False, time_budget=0.5)
    dt = time.time() - t0
    assert 0.4 < dt < 5.0 and res["n_grad"] > 100 and res["busy_seconds"] > 0
    full = o.nuts_batch(genes["counts"], oracle.nuts_cfg(5, 3, seed=1), n_chains=2, prior_mean=genes["beta_prior_mean"],
                        prior_std=genes["beta_prior_std"], n_threads=4)
    assert (full["status"] == 0).all() and full["draws"].shape[2] == 3 and np.isfinite(full["draws"]).all()
