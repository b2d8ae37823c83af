"""`splotch_vinit` on the GPU (/root/reference/bin/splotch_vinit:137-194): Pathfinder with the
device density, the CmdStan `pathfinder` argv, and the reference's own CSV -> JSON -> `init=` hand-over."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
from csplotch_b200 import api, cli, pathfinder, rdump, stancsv


def _c1_data_dir(golden_dir, tmp_path):
    d = tmp_path / "data" / "0"
    d.mkdir(parents=True)
    for g in range(1, 4):
        (d / ("data_%d.R" % g)).write_text(open(os.path.join(golden_dir, "c1", "data_%d.R" % g)).read())
    return str(tmp_path / "data")


def test_gpu_target_gives_the_oracle_targets_pathfinder(golden_dir):
    """Same control logic, same variates: the L-BFGS paths on the device density follow those on the CPU oracle
    (early iterates to rounding; chaotic later, so only the first steps are compared) and the draws are good starts."""
    ds = [rdump.read_rdump(os.path.join(golden_dir, "c1", "data_%d.R" % g)) for g in (1, 2)]
    shared = {k: v for k, v in ds[0].items() if k not in rdump.GENE_KEYS}
    model = api.Model(shared, "comp_splotch_stan_model")
    batch = api.Batch(model, np.stack([d["counts"] for d in ds]), np.stack([d["beta_prior_mean"] for d in ds]),
                      np.stack([d["beta_prior_std"] for d in ds]), gene_id0=1)
    tg = pathfinder.BatchTarget(batch)

    class OracleTarget:
        G, dim, gene_id0 = 2, model.dim, 1
        o = [oracle.Oracle(d, "comp_splotch_stan_model") for d in ds]

        def lp_grad(self, th, go):
            out = [self.o[go[i]].log_prob_grad(th[i]) for i in range(len(th))]
            return np.array([a for a, _ in out]), np.stack([b for _, b in out])

        def lp(self, th, go):
            return np.array([self.o[go[i]].log_prob_grad(th[i], want_grad=False) for i in range(len(th))])

    kw = dict(num_paths=2, num_draws=64, seed=6, max_lbfgs_iters=3)
    a, b = pathfinder.pathfinder(tg, **kw), pathfinder.pathfinder(OracleTarget(), **kw)
    assert np.array_equal(a.best_iter, b.best_iter) and np.allclose(a.elbo, b.elbo, rtol=1e-8, atol=1e-8)
    assert np.allclose(a.draws, b.draws, rtol=1e-7, atol=1e-7) and np.allclose(a.lp, b.lp, rtol=1e-8, atol=1e-7)
    full = pathfinder.pathfinder(tg, num_paths=4, num_draws=200, seed=6, max_lbfgs_iters=200)
    rng = np.random.default_rng(0)
    for g in range(2):
        rand = tg.lp(rng.uniform(-2, 2, (50, model.dim)), np.full(50, g, dtype=np.int32))
        assert np.isfinite(full.lp[g]).all() and np.median(full.lp[g]) > rand.max()


def test_splotch_vinit_cli(golden_dir, tmp_path):
    data_dir = _c1_data_dir(golden_dir, tmp_path)
    out = str(tmp_path / "out")
    argv = ["-g", "1-3", "-d", data_dir, "-o", out, "-b", "comp_splotch_stan_model", "-n", "30", "-c", "2", "--seed", "4",
            "--max-lbfgs-iters", "100"]
    assert cli.main_vinit(argv) == 0
    for g in (1, 2, 3):
        lines = open(os.path.join(out, "0", "combined_%d.csv" % g)).read().splitlines()
        assert lines[0].startswith("lp__,accept_stat__") and len(lines) == 1 + 2 * 30
        post = stancsv.read_stan_csv(os.path.join(out, "0", "combined_%d.csv" % g), ["lp__", "beta_level_1", "sigma"])
        assert np.isfinite(post["lp__"]).all() and post["beta_level_1"].shape == (60, 1, 14, 5)
    assert cli.main_vinit(argv[:5] + [str(tmp_path / "sum"), "-s"] + argv[6:]) == 0
    got = stancsv.read_summary(os.path.join(str(tmp_path / "sum"), "0", "combined_2.hdf5"))
    assert got["beta_level_1"]["samples"].shape == (60, 1, 14, 5) and got["lambda"]["mean"].shape == (10,)


def test_shim_pathfinder_method_and_reference_hand_over(golden_dir, tmp_path):
    """`BINARY pathfinder num_draws=M num_paths=C data file=… output file=…` (bin/splotch_vinit:140), then the
    reference's extraction (bin/splotch_vinit:146-182, restated) and `BINARY sample … init=init_<g>_<c>.json` (:194)."""
    path = os.path.join(golden_dir, "c1", "data_2.R")
    pf_csv = str(tmp_path / "output_2.csv")
    argv = ["/x/comp_splotch_stan_model", "pathfinder", "num_draws=40", "num_paths=2", "max_lbfgs_iters=80",
            "data", "file=" + path, "output", "file=" + pf_csv]
    assert cli.cmdstan_shim(argv) == 0
    lines = [ln for ln in open(pf_csv) if not ln.startswith("#")]
    header = lines[0].strip().split(",")
    assert header[:4] == ["lp_approx__", "lp__", "path__", "psi.1"] and len(lines) == 41
    rows = np.array([[float(v) for v in ln.strip().split(",")] for ln in lines[1:]])
    assert np.isfinite(rows).all() and set(rows[:, 2]) <= {1.0, 2.0}
    for c, i in enumerate((3, 17)):
        init = {k: float(v) for k, v in zip(header, lines[1 + i].strip().split(","))
                if k not in ["lp_approx__", "lp__", "path__"]}
        p = tmp_path / ("init_2_%d.json" % (c + 1))
        p.write_text(json.dumps(init, indent=2))
        out = str(tmp_path / ("output_2_%d.csv" % (c + 1)))
        argv = ["/x/comp_splotch_stan_model", "sample", "num_samples=20", "num_warmup=150", "random", "id=%d" % (c + 1),
                "init=" + str(p), "data", "file=" + path, "output", "file=" + out, "refresh=10"]
        assert cli.cmdstan_shim(argv) == 0
        post = stancsv.read_stan_csv(out, ["lp__"])
        assert post["lp__"].shape[0] == 20 and np.isfinite(post["lp__"]).all()
