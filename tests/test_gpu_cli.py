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


def test_cmdstan_argv_shim(golden_dir, tmp_path):
    """`BINARY sample num_samples=N num_warmup=N random id=CHAIN data file=IN.R output file=OUT.csv refresh=10`
    (bin/splotch:131): the shim honours the executable's argv so the unmodified reference wrapper can call it."""
    out = str(tmp_path / "output_3_2.csv")
    argv = ["/opt/bin/comp_splotch_stan_model", "sample", "num_samples=25", "num_warmup=25", "random", "id=2", "data",
            "file=" + os.path.join(golden_dir, "c1", "data_3.R"), "output", "file=" + out, "refresh=10"]
    assert cli.cmdstan_shim(argv) == 0
    lines = open(out).read().splitlines()
    assert lines[0].startswith("lp__,") and len(lines) == 26
    post = stancsv.read_stan_csv(out, ["beta_level_1", "log_lambda", "tau", "a", "theta"])
    assert post["beta_level_1"].shape == (25, 1, 14, 5) and np.all(post["tau"] > 0) and np.all(post["a"] < 1)
    assert cli.cmdstan_shim(["/opt/bin/comp_splotch_stan_model", "optimize"]) == 1       # a method this path does not have
    assert cli.cmdstan_shim(["/opt/bin/comp_splotch_stan_model", "pathfinder"]) == 1     # no data file
    assert cli.cmdstan_shim(["/opt/bin/comp_splotch_stan_model", "sample", "data", "file=/no/such/data_1.R"]) == 1


def test_error_behaviour(golden_dir):
    """Bad inputs are refused with a message, like Stan's data-block validation, never silently."""
    from csplotch_b200._lib import CsplotchError
    shared = synth.shared_tiny(seed=1, family="comp_splotch_stan_model")
    genes = synth.simulate_counts(shared, 2, 2)
    bad = dict(shared); bad["nb"] = 7
    with pytest.raises(CsplotchError, match="nb must be"):
        api.Model(bad, "comp_splotch_stan_model")
    bad = dict(shared); bad["D"] = np.asarray(shared["D"]).copy(); bad["D"][0] = 99
    with pytest.raises(CsplotchError, match="D out of range"):
        api.Model(bad, "comp_splotch_stan_model")
    bad = dict(shared); bad["size_factors"] = np.asarray(shared["size_factors"]).copy(); bad["size_factors"][3] = 0.0
    with pytest.raises(CsplotchError, match="size_factors"):
        api.Model(bad, "comp_splotch_stan_model")
    bad = dict(shared); bad["D_sparse"] = np.asarray(shared["D_sparse"]).copy(); bad["D_sparse"][0] += 1
    with pytest.raises(CsplotchError, match="D_sparse"):
        api.Model(bad, "comp_splotch_stan_model")
    big = synth.shared_tiny(seed=1, family="comp_splotch_stan_model", K=17)
    with pytest.raises(CsplotchError, match="N_celltypes"):
        api.Model(big, "comp_splotch_stan_model")
    model = api.Model(shared, "comp_splotch_stan_model")
    c = genes["counts"].copy(); c[0, 0] = -1
    with pytest.raises(CsplotchError, match="counts"):
        api.Batch(model, c, genes["beta_prior_mean"], genes["beta_prior_std"])
    with pytest.raises(CsplotchError, match="beta_prior"):
        api.Batch(model, genes["counts"])
    b = api.Batch(model, genes["counts"], genes["beta_prior_mean"], genes["beta_prior_std"])
    with pytest.raises(CsplotchError, match="max_depth"):
        api.Sampler(b, api.nuts_cfg(10, 10, 2, max_depth=20))
    with pytest.raises(ValueError):
        api.family_from_binary("/usr/bin/not_a_splotch_model")
    # an initial point with non-finite density is reported per chain (status 1), other chains run
    th0 = np.zeros((2, 2, model.dim)); th0[0, 0, :] = 800.0
    s = api.Sampler(b, api.nuts_cfg(5, 5, 2, seed=1), theta_init=th0)
    s.run()
    st, it = s.status()
    assert st[0, 0] == 1 and it[0, 0] == 0 and (st.reshape(-1)[1:] == 0).all() and (it.reshape(-1)[1:] == 10).all()


def test_pinned_buffers_give_the_same_results():
    """Page-locked input / output arrays (csplotch_host_alloc) are a transfer optimisation only."""
    shared = synth.shared_tiny(seed=3)
    genes = synth.simulate_counts(shared, 5, 11)
    model = api.Model(shared, "comp_splotch_stan_model")
    cfg = api.nuts_cfg(15, 10, 2, seed=4)
    b0 = api.Batch(model, genes["counts"], genes["beta_prior_mean"], genes["beta_prior_std"])
    s0 = api.Sampler(b0, cfg).run()
    want_draws, want_summ = s0.draws(), s0.spot_summaries()
    counts = api.pinned_empty(genes["counts"].shape, np.int32)
    counts[...] = genes["counts"]
    b1 = api.Batch(model, counts, genes["beta_prior_mean"], genes["beta_prior_std"])
    s1 = api.Sampler(b1, cfg).run()
    out = (api.pinned_empty(want_draws[0].shape), api.pinned_empty(want_draws[1].shape),
           api.pinned_empty(want_draws[2].shape, np.int32))
    got_draws = s1.draws(out=out)
    summ_buf = api.pinned_empty((6, 5, model.N))
    got_summ = s1.spot_summaries(out=summ_buf)
    for a, b in zip(want_draws, got_draws):
        assert np.array_equal(a, b)
    assert got_draws[0] is out[0]
    for k in want_summ:
        assert np.array_equal(want_summ[k], got_summ[k])
    del out, got_draws, summ_buf, got_summ, counts          # frees the page-locked memory with the last view


def test_init_file_round_trip(golden_dir, tmp_path):
    """CmdStan `init=FILE.json` (bin/splotch_vinit:177-194): constrained values keyed by CSV column name -> theta."""
    import json
    for fam, path in (("comp_splotch_stan_model", os.path.join(golden_dir, "c1", "data_3.R")), ("splotch_stan_model", None)):
        if path:
            data = rdump.read_rdump(path)
        else:
            shared0 = synth.shared_tiny(seed=2, family=fam, n_levels=2, zi=0)
            data = dict(shared0, counts=synth.simulate_counts(shared0, 1, 3)["counts"][0])
        shared = {k: v for k, v in data.items() if k not in rdump.GENE_KEYS}
        model = api.Model(shared, fam)
        comp = fam.startswith("comp")
        batch = api.Batch(model, np.asarray(data["counts"])[None], np.asarray(data["beta_prior_mean"])[None] if comp else None,
                          np.asarray(data["beta_prior_std"])[None] if comp else None)
        rng = np.random.default_rng(8)
        theta = rng.normal(0, 0.7, model.dim)
        cols, vals = model.column_names(), batch.write_array(theta[None])[0]
        init = dict(zip(cols, vals.tolist()))           # a Pathfinder / CSV row: parameters and transformed parameters
        init.update({"lp__": -1.0, "lp_approx__": 0.0})
        got = api.unconstrain(model, init)
        assert np.max(np.abs(got - theta)) < 1e-11
        s = api.Sampler(batch, api.nuts_cfg(5, 5, 1, seed=1), theta_init=got[None, None, :])
        assert np.max(np.abs(s.state()[0][0, 0] - theta)) < 1e-11
        # absent parameters start uniform(-2, 2); structured (CmdStan JSON) arrays are accepted as well
        part = {"sigma": float(init["sigma"]), "noise_raw": [init["noise_raw.%d" % (i + 1)] for i in range(model.N)]}
        th2 = api.unconstrain(model, part, seed=3)
        lay = {n: i for n, i, _ in api.param_layout(model)}
        assert abs(th2[lay["sigma"]] - theta[lay["sigma"]]) < 1e-12
        assert np.allclose(th2[[lay["noise_raw.%d" % (i + 1)] for i in range(model.N)]],
                           theta[[lay["noise_raw.%d" % (i + 1)] for i in range(model.N)]], atol=1e-12)
        assert np.all(np.abs(th2) <= np.maximum(2.0, np.abs(theta) + 1e-9))
        if comp:
            p = tmp_path / "init_3_1.json"
            p.write_text(json.dumps(init))
            out = str(tmp_path / "o.csv")
            argv = ["/x/comp_splotch_stan_model", "sample", "num_samples=5", "num_warmup=5", "random", "id=1", "init=" + str(p),
                    "data", "file=" + path, "output", "file=" + out]
            assert cli.cmdstan_shim(argv) == 0 and len(open(out).read().splitlines()) == 6
            bad = tmp_path / "bad.json"
            bad.write_text(json.dumps({"tau.1": -3.0}))
            argv[6] = "init=" + str(bad)
            assert cli.cmdstan_shim(argv) == 1


def test_resume_and_report(golden_dir, tmp_path):
    """`--resume` skips genes whose output exists (file-level resume); `--report` appends one JSON line
    per device batch with the work done, sampler health and ESS / R-hat."""
    import json
    data_dir = _c1_data_dir(golden_dir, tmp_path)
    out, rep = str(tmp_path / "out"), str(tmp_path / "report.jsonl")
    common = ["-d", data_dir, "-o", out, "-b", "comp_splotch_stan_model", "-n", "30", "-c", "2", "--seed", "5", "-s"]
    assert cli.main(["-g", "3-4"] + common) == 0
    stamp = {g: os.path.getmtime(os.path.join(out, "0", n)) for g in (3, 4) for n in os.listdir(os.path.join(out, "0"))
             if n.startswith("combined_%d." % g)}
    assert cli.main(["-g", "1-7", "--resume", "--report", rep] + common) == 0
    lines = [json.loads(l) for l in open(rep)]
    assert [l["genes"] for l in lines] == [[1, 2], [5, 7]]                       # 3 and 4 were not sampled again
    for g in (3, 4):
        assert stamp[g] == [os.path.getmtime(os.path.join(out, "0", n)) for n in os.listdir(os.path.join(out, "0"))
                            if n.startswith("combined_%d." % g)][0]
    for l in lines:
        assert l["device_ms"] > 0 and l["grad_evals"] >= l["leapfrogs"] > 0 and l["failed_genes"] == []
        assert sum(l["treedepth_hist"]) == l["n_genes"] * 2 * 30 and l["median_min_bulk_ess"] > 0
    assert len([n for n in os.listdir(os.path.join(out, "0")) if n.startswith("combined_")]) == 7
