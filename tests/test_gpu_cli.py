"""End-to-end drop-in tests on the GPU: the `splotch` CLI on data_<g>.R files (the reference's own c1
example through its own input pipeline), both output modes, and the real SNAP25 fixture."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
from csplotch_b200 import api, cli, rdump, stancsv, synth


def _c1_data_dir(golden_dir, tmp_path):
    d = tmp_path / "data" / "0"
    d.mkdir(parents=True)
    for g in range(1, 8):
        (d / ("data_%d.R" % g)).write_text(open(os.path.join(golden_dir, "c1", "data_%d.R" % g)).read())
    return str(tmp_path / "data")


def test_cli_csv_and_summary_agree(golden_dir, tmp_path):
    data_dir = _c1_data_dir(golden_dir, tmp_path)
    out_csv, out_sum = str(tmp_path / "out_csv"), str(tmp_path / "out_sum")
    common = ["-d", data_dir, "-b", "/some/where/comp_splotch_stan_model", "-n", "40", "-c", "2", "--seed", "5"]
    assert cli.main(["-g", "1-7", "-o", out_csv] + common) == 0
    assert cli.main(["-g", "1-7", "-o", out_sum, "-s"] + common) == 0
    variables = stancsv.SUMMARY_VARS
    for g in range(1, 8):
        p = os.path.join(out_csv, "0", "combined_%d.csv" % g)
        lines = open(p).read().splitlines()
        assert lines[0].startswith("lp__,accept_stat__,stepsize__,treedepth__,n_leapfrog__,divergent__,energy__,psi.1")
        assert len(lines) == 1 + 2 * 40 and not any(l.startswith("#") for l in lines)
        post = stancsv.read_stan_csv(p, variables + ["lp__", "treedepth__"])
        assert post["beta_level_1"].shape == (80, 1, 14, 5) and post["log_lambda"].shape == (80, 10)
        assert np.all(np.isfinite(post["lp__"])) and post["treedepth__"].max() <= 10
        want = stancsv.summarize({k: v for k, v in post.items() if k in variables})
        got = stancsv.read_summary(os.path.join(out_sum, "0", "combined_%d.hdf5" % g))
        assert set(got) == set(want)            # same groups as the reference's HDF5 (Tutorial2.ipynb cell 9)
        for key in want:
            for stat in want[key]:
                assert got[key][stat].shape == want[key][stat].shape, (key, stat)
                assert np.allclose(got[key][stat], want[key][stat], rtol=1e-9, atol=1e-12), (g, key, stat)
    # the reference's consumers index these exactly like this (bin/splotch_compile_betas:70-71, _lambdas:49)
    assert got["beta_level_1"]["mean"].flatten().shape == (70,) and got["lambda"]["mean"].shape == (10,)


def test_pystan_like_interface(golden_dir):
    data = rdump.read_rdump(os.path.join(golden_dir, "c1", "data_2.R"))
    fit = api.SplotchModel(file="stan/comp_splotch_stan_model.stan").sampling(data=data, iter=120, chains=4, seed=3)
    ex = fit.extract(permuted=True)
    assert ex["beta_level_1"].shape == (4 * 60, 1, 14, 5)       # simulation/fitting.py:51 reads exactly this
    assert ex["sigma"].shape == (240,) and np.all(ex["sigma"] > 0) and np.all((ex["theta"] > 0) & (ex["theta"] < 1))
    diag = fit.sampler_diagnostics()
    assert diag["treedepth__"].max() <= 10 and np.all(diag["n_leapfrog__"] >= 1)


def test_snap25_real_counts(golden_dir):
    """Science/human/data_SNAP25.R (first 6 tissue sections, real counts): variable names of
    splotch_human.stan map 1:1 onto splotch_stan_model with 2 levels."""
    z = np.load(os.path.join(golden_dir, "snap25_head.npz"))
    ns = z["N_spots"]
    offs = np.concatenate([[0], np.cumsum(ns)])
    W0 = z["W_sparse"] - 1
    pairs, eigs = [], []
    for t in range(len(ns)):
        m = (W0[:, 0] >= offs[t]) & (W0[:, 0] < offs[t + 1])
        p = W0[m] - offs[t]
        pairs.append(p)
        eigs.append(synth.car_eigenvalues(p, int(ns[t])))
    shared = synth.build_shared(ns, pairs, eigs, z["D"], z["tissue_mapping"], (int(z["N_onsets_tissue_locations"]),
                                int(z["N_donors"])), z["donor_mapping"], [], z["size_factors"],
                                int(z["N_covariates"]), zi=1, car=1)
    assert np.array_equal(shared["D_sparse"], z["D_sparse"])
    genes = {"counts": z["counts"][None, :].astype(np.int32)}
    for family in ("splotch_stan_model", "splotch_stan_model_csr"):
        model = api.Model(shared, family)
        b = api.Batch(model, genes["counts"])
        o = oracle.Oracle(synth.gene_dict(shared, genes, 0), family)
        rng = np.random.default_rng(1)
        th = rng.normal(0, 0.7, (2, model.dim))
        lp, grad = b.log_prob_grad(th, gene_of=[0, 0])
        for i in range(2):
            lp0, g0 = o.log_prob_grad(th[i])
            assert abs(lp[i] - lp0) < 1e-10 * abs(lp0)
            assert np.max(np.abs(grad[i] - g0) / np.maximum(1, np.abs(g0))) < 1e-10
    s = api.Sampler(b, api.nuts_cfg(30, 10, 4, seed=2)).run()
    st, it = s.status()
    assert (st == 0).all() and (it == 40).all()
    summ = s.spot_summaries()
    assert np.all(np.isfinite(summ["lambda/mean"])) and np.all(summ["lambda/mean"] > 0)
