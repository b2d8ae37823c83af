"""The native per-gene R-dump reader (csplotch_rdump_read_genes, csrc/ingest.cpp) against the
Python reader that follows the reference's read_rdump (/root/reference/splotch/utils.py:78-111), on the files
the reference's own splotch_generate_input_files wrote (tests/golden/c1) and on edge cases.  Host code only: these
tests call the library without a GPU."""
import os

import numpy as np
import pytest

from csplotch_b200 import rdump
from csplotch_b200._lib import CsplotchError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C1 = os.path.join(ROOT, "tests", "golden", "c1")


def test_c1_files_match_python_reader():
    paths = [os.path.join(C1, "data_%d.R" % g) for g in range(1, 8)]
    want = [rdump.read_rdump(p) for p in paths]
    N, K = len(want[0]["counts"]), len(want[0]["beta_prior_mean"])
    got = rdump.read_genes_native(paths, N, K)
    assert got["counts"].dtype == np.int32 and got["counts"].shape == (7, N)
    for g in range(7):
        assert np.array_equal(got["counts"][g], want[g]["counts"])
        assert np.array_equal(got["beta_prior_mean"][g], want[g]["beta_prior_mean"])   # bit-exact: same strtod
        assert np.array_equal(got["beta_prior_std"][g], want[g]["beta_prior_std"])
    # counts only (non-compositional callers pass K = 0)
    only = rdump.read_genes_native(paths, N, 0, n_threads=3)
    assert set(only) == {"counts"} and np.array_equal(only["counts"], got["counts"])


def _write(path, data):
    rdump.to_rdump(data, str(path))


def test_round_trip_many_files_and_threads(tmp_path):
    rng = np.random.default_rng(5)
    N, K, G = 5000, 7, 37
    shared = {"N_spots": np.array([N]), "E": rng.dirichlet(np.ones(K), N), "eig_values": rng.normal(size=N)}
    counts = rng.poisson(rng.gamma(0.5, 40.0, (G, 1)), (G, N)).astype(np.int64)
    counts[3] = 0
    counts[4, 17] = 2147483647
    pm, ps = rng.normal(size=(G, K)), rng.gamma(2.0, 1.0, (G, K))
    pm[2, 1] = 1e-300
    pm[2, 2] = -123456789.123456789
    paths = []
    for g in range(G):
        d = dict(shared)
        d["counts"] = counts[g]
        d["beta_prior_mean"], d["beta_prior_std"] = pm[g], ps[g]
        p = tmp_path / ("data_%d.R" % (g + 1))
        _write(p, d)
        paths.append(str(p))
    for nt in (1, 4, 0):
        got = rdump.read_genes_native(paths, N, K, n_threads=nt)
        assert np.array_equal(got["counts"], counts)
        assert np.array_equal(got["beta_prior_mean"], pm) and np.array_equal(got["beta_prior_std"], ps)


def test_any_key_order_and_foreign_formatting(tmp_path):
    # gene keys FIRST, followed by megabytes of shared covariates (the tail window has to grow), `, ` separators,
    # spaces around `<-`, a structure(...) wrapper and integral values written as doubles
    N = 300
    big = ",".join("%r" % x for x in np.linspace(0, 1, 400000))
    cnt = np.arange(N) % 7
    txt = ("counts_total <- 5\n"
           "counts  <-\n structure(c(" + ", ".join("%d.0" % c for c in cnt) + "), .Dim = c(%d))\n" % N
           + "beta_prior_mean <-\nc(0.5, -1e3,\n 2)\n"
           + "beta_prior_std <- c(1,2,3)\n"
           + "size_factors <-\n c(" + big + ")\n"
           + "xcounts <- c(1,2)\n")
    p = tmp_path / "data_1.R"
    p.write_text(txt)
    got = rdump.read_genes_native([str(p)], N, 3)
    assert np.array_equal(got["counts"][0], cnt)
    assert got["beta_prior_mean"][0].tolist() == [0.5, -1000.0, 2.0] and got["beta_prior_std"][0].tolist() == [1, 2, 3]
    py = rdump.read_rdump_gene_only(str(p))
    assert np.array_equal(py["counts"], cnt)


def test_errors_name_the_file(tmp_path):
    N = 20
    good = tmp_path / "data_1.R"
    _write(good, {"N_spots": np.array([N]), "counts": np.arange(N)})
    with pytest.raises(CsplotchError, match="data_2.R"):
        rdump.read_genes_native([str(good), str(tmp_path / "data_2.R")], N)          # missing file
    with pytest.raises(CsplotchError, match="expected 21"):
        rdump.read_genes_native([str(good)], N + 1)                                  # wrong length
    with pytest.raises(CsplotchError, match="beta_prior_mean"):
        rdump.read_genes_native([str(good)], N, 4)                                   # priors absent
    frac = tmp_path / "data_3.R"
    frac.write_text("counts <-\n c(1,2.5,3)\n")
    with pytest.raises(CsplotchError, match="non-integer"):
        rdump.read_genes_native([str(frac)], 3)
    nokey = tmp_path / "data_4.R"
    nokey.write_text("N_spots <-\n c(3)\nmycounts <-\n c(1,2,3)\n")
    with pytest.raises(CsplotchError, match="not found"):
        rdump.read_genes_native([str(nokey)], 3)
    assert rdump.read_genes_native([], N)["counts"].shape == (0, N)


# ------------------------------------------------------------------------------------------ native CSV writer
def test_native_csv_writer_round_trips_every_double(tmp_path):
    """`csplotch_csv_write` (bin/splotch:131,138-140 stand-in): what read_stan_csv's float() reads back is the
    double that was written, bit for bit; integral values have no decimal point; many files in one call."""
    from csplotch_b200 import stancsv
    rng = np.random.default_rng(0)
    F, S, n = 5, 40, 300
    cols = ["psi.%d" % (i + 1) for i in range(n - 1)] + ["sigma"]
    diag = np.abs(rng.normal(size=(F, S, 7)))
    diag[..., 3:6] = rng.integers(0, 1023, (F, S, 3))
    vals = rng.normal(size=(F, S, n)) * 10.0 ** rng.integers(-300, 300, (F, S, n))
    vals[0, 0, :8] = [0.0, -0.0, 5e-324, 1.7976931348623157e308, 1e15, 123456789012345678.0, 0.1, -2.5]
    vals[1, 1, :3] = [np.nan, np.inf, -np.inf]
    paths = [str(tmp_path / ("combined_%d.csv" % f)) for f in range(F)]
    stancsv.write_combined_csvs(paths, cols, diag, vals, n_threads=3)
    for f in range(F):
        lines = open(paths[f]).read().split("\n")
        assert lines[-1] == "" and len(lines) == S + 2 and not any(l.startswith("#") for l in lines)
        assert lines[0] == ",".join(stancsv.SAMPLER_COLS + cols)
        got = np.array([[float(t) for t in l.split(",")] for l in lines[1:-1]])
        want = np.concatenate([diag[f], vals[f]], axis=1)
        assert np.array_equal(got, want, equal_nan=True)                      # exact: shortest round-trip text
        toks = lines[1].split(",")
        assert all("." not in t and "e" not in t for t in toks[3:6])           # treedepth__, n_leapfrog__, divergent__
        post = stancsv.read_stan_csv(paths[f], ["psi", "sigma", "lp__"])
        assert np.array_equal(post["sigma"][:, 0], vals[f, :, -1], equal_nan=True)
    assert open(paths[0]).read().split("\n")[1].split(",")[7:15] == [
        "0", "0", "5e-324", "1.7976931348623157e+308", "1e+15", "123456789012345680", "0.1", "-2.5"]
    assert not any(p.endswith(".tmp") for p in os.listdir(tmp_path))


def test_native_csv_writer_sig_figs_and_errors(tmp_path):
    from csplotch_b200 import stancsv
    from csplotch_b200._lib import CsplotchError
    vals = np.array([[[np.pi * 1e5, -np.e * 1e-7, 2.0, 1234567.0]]])
    lead = np.array([[[-1234.56789, 1.0, 3.0]]])
    p = str(tmp_path / "pf.csv")
    stancsv.write_combined_csvs([p], ["a", "b", "c", "d"], lead, vals, lead_cols=stancsv.PATHFINDER_COLS, sig_figs=6)
    hdr, row = open(p).read().splitlines()
    assert hdr == "lp_approx__,lp__,path__,a,b,c,d"
    assert row == "-1234.57,1,3,314159,-2.71828e-07,2,1234567"          # CmdStan's default precision
    with pytest.raises(CsplotchError, match="no_such_dir"):
        stancsv.write_combined_csvs([str(tmp_path / "no_such_dir" / "x.csv")], ["a", "b", "c", "d"], lead, vals,
                                    lead_cols=stancsv.PATHFINDER_COLS)
    # zero rows: header only (a chain file of a run with num_samples=0)
    stancsv.write_combined_csvs([p], ["a"], np.zeros((1, 0, 7)), np.zeros((1, 0, 1)))
    assert open(p).read() == ",".join(stancsv.SAMPLER_COLS + ["a"]) + "\n"
