"""I/O parity against vectors produced by the reference's own Python (tests/golden/make_golden.py ran
/root/reference/splotch/utils.py to_rdump / read_rdump / read_stan_csv and
bin/splotch_generate_input_files on examples_csplotch)."""
import os
import re

import numpy as np
import pytest

import oracle
from csplotch_b200 import rdump, stancsv, synth


def test_read_rdump_matches_reference_parse(golden_dir):
    mine = rdump.read_rdump(os.path.join(golden_dir, "io", "ref_written.R"))
    ref = np.load(os.path.join(golden_dir, "io", "ref_written_parsed.npz"))
    for k in ref.files:
        assert np.array_equal(np.asarray(mine[k]), ref[k]), k
    assert np.asarray(mine["empty"]).size == 0
    assert np.asarray(mine["Wempty"]).size == 0


def test_to_rdump_writes_what_the_reference_writes(golden_dir, tmp_path):
    src = os.path.join(golden_dir, "io", "ref_written.R")
    d = rdump.read_rdump(src)
    d["empty"] = []
    d["Wempty"] = np.zeros((0, 0))
    out = tmp_path / "mine.R"
    rdump.to_rdump(d, str(out))
    assert out.read_text() == open(src).read()


def test_rdump_whitespace_variants(tmp_path):
    # the Science/human/data_SNAP25.R dialect: ", " separators and ".Dim = c("
    p = tmp_path / "x.R"
    p.write_text("W_sparse <-\nstructure(c(1, 1, 2, 3, 4, 5), .Dim = c(3, 2))\nN_tissues <- 80\ncounts <-\nc(0, 2, 1)\n"
                 "sf <-\nc(0.5, 1.25e-1, 3)\n")
    d = rdump.read_rdump(str(p))
    assert np.array_equal(d["W_sparse"], [[1, 3], [1, 4], [2, 5]])
    assert d["N_tissues"] == 80 and np.array_equal(d["counts"], [0, 2, 1])
    assert np.allclose(d["sf"], [0.5, 0.125, 3])
    g = rdump.read_rdump_gene_only(str(p))
    assert list(g) == ["counts"]


def test_read_stan_csv_matches_reference(golden_dir):
    variables = ["log_lambda", "beta_level_1", "psi", "sigma", "theta"]
    mine = stancsv.read_stan_csv(os.path.join(golden_dir, "io", "combined_ref.csv"), variables)
    ref = np.load(os.path.join(golden_dir, "io", "combined_ref_parsed.npz"))
    assert sorted(mine) == sorted(ref.files)
    for k in ref.files:
        assert mine[k].shape == ref[k].shape, k
        assert np.array_equal(mine[k], ref[k]), k


def test_summary_layout(golden_dir, tmp_path):
    variables = stancsv.SUMMARY_VARS
    post = stancsv.read_stan_csv(os.path.join(golden_dir, "io", "combined_ref.csv"), variables)
    summ = stancsv.summarize(post)
    assert set(summ["beta_level_1"]) == {"mean", "std", "samples"}
    assert set(summ["log_lambda"]) == {"mean", "std"}
    assert summ["sigma"]["mean"].shape == (1,)                       # scalars stored with shape (1,)
    assert np.allclose(summ["lambda"]["mean"], np.exp(post["log_lambda"]).mean(0))
    assert np.allclose(summ["psi"]["std"], post["psi"].std(0))      # ddof = 0
    path = stancsv.write_summary(str(tmp_path / "combined_1.hdf5"), summ)
    back = stancsv.read_summary(path)
    for k in summ:
        for st in summ[k]:
            assert np.array_equal(back[k][st], summ[k][st])


def test_combined_csv_round_trip(tmp_path):
    cols = ["psi.1", "psi.2", "sigma", "beta_level_1.1.1", "beta_level_1.2.1"]
    rng = np.random.default_rng(0)
    diag = np.abs(rng.normal(size=(5, 7)))
    diag[:, 3:6] = [[3, 7, 0]] * 5
    vals = rng.normal(size=(5, 5))
    p = str(tmp_path / "combined_3.csv")
    stancsv.write_combined_csv(p, cols, diag, vals)
    lines = open(p).read().splitlines()
    assert lines[0].startswith("lp__,accept_stat__,stepsize__,treedepth__,n_leapfrog__,divergent__,energy__,psi.1")
    assert not any(l.startswith("#") for l in lines) and len(lines) == 6
    post = stancsv.read_stan_csv(p, ["psi", "sigma", "beta_level_1", "lp__"])
    assert np.array_equal(post["psi"], vals[:, :2])
    assert np.array_equal(post["beta_level_1"][:, :, 0], vals[:, 3:5])
    assert np.array_equal(post["lp__"][:, 0], diag[:, 0])


def test_c1_fixture_is_what_the_survey_says(golden_dir):
    """examples_csplotch through the reference pipeline: 1 tissue section, N=10 spots, 12 pairs, K=5,
    C=14, D=94."""
    for g in range(1, 8):
        d = rdump.read_rdump(os.path.join(golden_dir, "c1", "data_%d.R" % g))
        assert int(np.sum(d["N_spots"])) == 10 and int(d["W_n"][0]) == 12
        assert d["N_celltypes"] == 5 and d["N_covariates"] == 14 and d["zi"] == 1 and d["car"] == 1
        o = oracle.Oracle(d, "comp_splotch_stan_model")
        assert o.dim == 94
        lp, grad = o.log_prob_grad(np.zeros(94))
        assert np.isfinite(lp) and np.all(np.isfinite(grad))
        # CSR arrays in the file describe the same graph as W_sparse (utils.py:592-614)
        V, U = synth.pairs_to_csr(np.asarray(d["W_sparse"]) - 1, 10)
        assert np.array_equal(V, d["V"]) and np.array_equal(U, d["U"])
        assert np.allclose(np.asarray(d["E"]).sum(1), 1.0)


def test_c2_geometry_fixture(golden_dir):
    shared = synth.shared_c2()
    assert shared["N_level_1"] == 2 and shared["N_level_2"] == 4 and shared["N_covariates"] == 11
    N = int(np.sum(shared["N_spots"]))
    assert N == 2594 and shared["zi"] == 0 and shared["car"] == 1
    z = np.load(os.path.join(golden_dir, "c2_geometry.npz"))
    assert np.array_equal(shared["W_sparse"], z["W_sparse"])
    assert np.array_equal(shared["U"], z["U"]) and np.array_equal(shared["V"], z["V"])
    assert np.allclose(shared["eig_values"], z["eig_values"])


def test_synthetic_c3_matches_survey_counts():
    pairs = synth.hex_pairs()
    assert len(pairs) == 14693                                        # pairs per full 78x64 array
    deg = np.bincount(pairs.ravel(), minlength=4992)
    assert dict(zip(*np.unique(deg, return_counts=True))) == {2: 2, 3: 78, 4: 124, 5: 76, 6: 4712}


def test_cli_gene_spec():
    from csplotch_b200 import cli
    assert cli.parse_genes("7") == [7]
    assert cli.parse_genes("1-3,9,11-12") == [1, 2, 3, 9, 11, 12]
    assert cli.data_path("DATA", 3661).endswith(os.path.join("DATA", "36", "data_3661.R"))


def test_param_layout_and_unconstrain_without_a_device():
    """CmdStan `init=` files (bin/splotch_vinit:177-194): names -> positions in the unconstrained vector
    (include/csplotch.h layout) and the constraining transforms; pure host logic, checked on a stand-in model."""
    from types import SimpleNamespace as NS
    from csplotch_b200 import api
    comp = NS(levels=(1, 2, 0), N=5, C=3, K=2, car=1, zi=1, nb=1, family=1, dim=5 + 3 * 3 * 2 + 2 + 1 + 1 + 1 + 1 + 5)
    lay = api.param_layout(comp)
    assert sorted(i for _, i, _ in lay) == list(range(comp.dim))
    pos = {n: i for n, i, _ in lay}
    # psi | beta_raw[(i*C + j)*K + k] | tau a sigma sigma_level_2 theta phi | noise_raw
    assert pos["psi.1"] == 0 and pos["beta_raw.1.1.1"] == 5 and pos["beta_raw.1.1.2"] == 6 and pos["beta_raw.2.1.1"] == 5 + 6
    assert [pos[n] for n in ("tau.1", "a.1", "sigma", "sigma_level_2.1", "theta.1", "phi.1")] == list(range(23, 29))
    assert pos["noise_raw.5"] == comp.dim - 1
    plain = NS(levels=(2, 0, 0), N=4, C=3, K=1, car=0, zi=0, nb=0, family=0, dim=6 + 1 + 4)
    pos = {n: i for n, i, _ in api.param_layout(plain)}
    assert pos["beta_raw.2.1"] == 1 and pos["beta_raw.1.2"] == 2 and pos["sigma"] == 6     # column-major i + L*j
    th = api.unconstrain(plain, {"sigma": 2.0, "beta_raw": [[1, 2, 3], [4, 5, 6]], "log_lambda.1": 9.0, "lp__": -3}, seed=1)
    assert th[:6].tolist() == [1, 4, 2, 5, 3, 6] and abs(th[6] - np.log(2.0)) < 1e-15
    assert np.all(np.abs(th[7:]) <= 2.0)                                                     # absent: uniform(-2, 2)
    th = api.unconstrain(comp, {"a.1": 0.25, "theta": [0.5], "phi.1": 1.0 + 1e-6, "tau": 3.0})
    pos = {n: i for n, i, _ in lay}
    assert abs(th[pos["a.1"]] - np.log(1 / 3)) < 1e-15 and th[pos["theta.1"]] == 0.0
    assert abs(th[pos["phi.1"]]) < 1e-9 and abs(th[pos["tau.1"]] - np.log(3.0)) < 1e-15
    for bad in ({"tau.1": -1.0}, {"a.1": 1.0}, {"phi.1": 1e-7}):
        with pytest.raises(api.CsplotchError):
            api.unconstrain(comp, bad)
