"""Control flow of `splotch_vinit` and of the shim's `pathfinder` method without a device: `api.Model / Batch / Sampler`
are replaced by stand-ins that answer from the CPU oracle (tests only — the product has no such path), so that the
host logic between the reference's files and the C ABI is exercised here: /root/reference/bin/splotch_vinit:137-194."""
import json
import os

import numpy as np
import pytest

import oracle
from csplotch_b200 import api, cli, rdump, stancsv


class _Model:
    def __init__(self, data, family):
        self.family = api.family_from_binary(family)
        self._shared, self._fam_name = dict(data), os.path.basename(str(family))
        d = dict(data, counts=np.zeros(int(np.sum(data["N_spots"])), dtype=np.int32))
        if self.family == 1:
            K = int(d["N_celltypes"])
            d.update(beta_prior_mean=np.zeros(K), beta_prior_std=np.ones(K))
        o = oracle.Oracle(d, self.family)
        self.N, self.dim, self.K, self.n_outputs = o.N, o.dim, o.K, o.n_out
        self.C = int(d["N_covariates"])
        self.levels = (int(d["N_level_1"]), int(d.get("N_level_2", 0)), int(d.get("N_level_3", 0)))
        self.car, self.zi = int(d.get("car", 0)), int(d.get("zi", 0))
        self.nb = int(d.get("nb", 0)) if self.family == 1 else 0
        L = sum(self.levels)
        self.n_beta = L * self.C * self.K
        self.n_scalar = self.dim - self.n_beta - (2 if self.car else 1) * self.N
        self.n_small = self.n_beta + self.n_scalar
        self.level_2_mapping = np.asarray(d.get("level_2_mapping", []), dtype=np.int64).reshape(-1)
        self.level_3_mapping = np.asarray(d.get("level_3_mapping", []), dtype=np.int64).reshape(-1)

    column_names = api.Model.column_names


class _Batch:
    def __init__(self, model, counts, beta_prior_mean=None, beta_prior_std=None, gene_id0=0, gene_ids=None):
        self.model, self.gene_id0 = model, gene_id0
        self.gene_ids = gene_ids
        counts = np.asarray(counts).reshape(-1, model.N)
        self.G = len(counts)
        self.beta_prior_mean = None if beta_prior_mean is None else np.asarray(beta_prior_mean, dtype=float).reshape(self.G, -1)
        self.beta_prior_std = None if beta_prior_std is None else np.asarray(beta_prior_std, dtype=float).reshape(self.G, -1)
        self.o = []
        for g in range(self.G):
            d = dict(model._shared, counts=counts[g])
            if model.family == 1:
                d.update(beta_prior_mean=np.asarray(beta_prior_mean).reshape(self.G, -1)[g],
                         beta_prior_std=np.asarray(beta_prior_std).reshape(self.G, -1)[g])
            self.o.append(oracle.Oracle(d, model.family))

    def log_prob_grad(self, theta, gene_of=None, jacobian=True, want_grad=True):
        theta = np.asarray(theta).reshape(-1, self.model.dim)
        go = np.zeros(len(theta), dtype=int) if gene_of is None else np.asarray(gene_of)
        out = [self.o[go[i]].log_prob_grad(theta[i], jacobian=jacobian) for i in range(len(theta))]
        lp = np.array([a for a, _ in out])
        return (lp, np.stack([b for _, b in out])) if want_grad else lp

    def write_array(self, theta, gene_of=None):
        theta = np.asarray(theta).reshape(-1, self.model.dim)
        go = np.zeros(len(theta), dtype=int) if gene_of is None else np.asarray(gene_of)
        return np.stack([self.o[go[i]].write_array(theta[i]) for i in range(len(theta))])


class _Sampler:
    started = []          # (num_warmup, theta_init) of every sampler made, for the assertions

    def __init__(self, batch, cfg, theta_init=None):
        self.batch, self.cfg, self.model = batch, cfg, batch.model
        self.G, self.nch = batch.G, cfg.num_chains
        self.theta_init = None if theta_init is None else np.asarray(theta_init).reshape(self.G, self.nch, -1)
        _Sampler.started.append((cfg.num_warmup, self.theta_init))

    def run(self, chunk=None):
        ocfg = oracle.nuts_cfg(self.cfg.num_warmup, self.cfg.num_samples, seed=int(self.cfg.seed))
        D = self.model.dim
        self._diag = np.zeros((self.G, self.nch, self.cfg.num_samples, 7))
        self._theta = np.zeros((self.G, self.nch, self.cfg.num_samples, D))
        self._status = np.zeros((self.G, self.nch), dtype=np.int32)
        for g in range(self.G):
            for c in range(self.nch):
                qi = None if self.theta_init is None else self.theta_init[g, c]
                out = self.batch.o[g].nuts(ocfg, gene_id=(self.batch.gene_id0 + g) if self.batch.gene_ids is None else int(self.batch.gene_ids[g]),
                                              chain_id=c + 1, q_init=qi)
                self._status[g, c] = out["status"]
                self._diag[g, c], self._theta[g, c] = out["draws"][:, :7], out["draws"][:, 7:]
        return self

    def status(self):
        return self._status, np.full((self.G, self.nch), self.cfg.num_warmup + self.cfg.num_samples)

    def draws(self, out=None):
        # the non-spot block in the library's order: beta_raw[(i*C + j)*K + k] | scalars in declaration order
        m = self.model
        lay = {n: i for n, i, _ in api.param_layout(m)}
        L = sum(m.levels)
        idx = []
        for i in range(L):
            for j in range(m.C):
                for k in range(m.K):
                    idx.append(lay["beta_raw.%d.%d.%d" % (i + 1, j + 1, k + 1)] if m.family == 1
                               else lay["beta_raw.%d.%d" % (i + 1, j + 1)])
        idx += [i for n, i, t in api.param_layout(m) if t != "id"]
        return self._diag, self._theta[..., idx], np.full((self.G, self.nch), self.cfg.num_samples, dtype=np.int32)

    def theta_draws(self, g0=0, n_genes=None):
        return self._theta[g0:(None if n_genes is None else g0 + n_genes)]

    def counters(self):
        return dict(n_leapfrog=int(self._diag[..., 4].sum()), n_grad=int(self._diag[..., 4].sum()), n_divergent=0,
                    n_launches=1)

    extract_small = api.Sampler.extract_small

    def spot_summaries(self, out=None):
        """What the device accumulates while sampling: mean / std (ddof = 0, chains pooled) of log_lambda, exp(log_lambda)
        and psi — here from write_array on the kept draws."""
        m = self.model
        cols = m.column_names()
        i0 = cols.index("log_lambda.1")
        res = {k: np.zeros((self.G, m.N)) for k in ("log_lambda/mean", "log_lambda/std", "lambda/mean", "lambda/std",
                                                    "psi/mean", "psi/std")}
        for g in range(self.G):
            wa = self.batch.write_array(self._theta[g].reshape(-1, m.dim), gene_of=np.full(self.nch * self.cfg.num_samples, g))
            ll = wa[:, i0:i0 + m.N]
            for key, arr in (("log_lambda", ll), ("lambda", np.exp(ll)), ("psi", wa[:, :m.N])):
                res[key + "/mean"][g], res[key + "/std"][g] = arr.mean(0), arr.std(0)
        if not m.car:
            res.pop("psi/mean"); res.pop("psi/std")
        return res


@pytest.fixture()
def host_only(monkeypatch):
    monkeypatch.setattr(api, "Model", _Model)
    monkeypatch.setattr(api, "Batch", _Batch)
    monkeypatch.setattr(api, "Sampler", _Sampler)
    monkeypatch.setattr(api, "check", lambda rc: None)          # csplotch_set_device needs a device
    _Sampler.started = []
    return _Sampler


def _c1_data_dir(golden_dir, tmp_path, genes):
    d = tmp_path / "data" / "0"
    d.mkdir(parents=True)
    for g in genes:
        (d / ("data_%d.R" % g)).write_text(open(os.path.join(golden_dir, "c1", "data_%d.R" % g)).read())
    return str(tmp_path / "data")


def test_splotch_vinit_control_flow(host_only, golden_dir, tmp_path):
    data_dir = _c1_data_dir(golden_dir, tmp_path, (2, 3))
    out = str(tmp_path / "out")
    argv = ["-g", "2-3", "-d", data_dir, "-o", out, "-b", "/opt/bin/comp_splotch_stan_model", "-n", "12", "-c", "2",
            "--seed", "4", "--max-lbfgs-iters", "40"]
    rep = str(tmp_path / "vinit.jsonl")
    assert cli.main_vinit(argv + ["--report", rep]) == 0
    (line,) = [json.loads(l) for l in open(rep)]
    assert line["num_warmup"] == 150 and line["pathfinder"]["lp_grad_evals"] > 0 and line["pathfinder"]["failed_genes"] == []
    assert line["pathfinder"]["median_lbfgs_iters"] <= 40
    (nw, init), = host_only.started
    assert nw == 150 and init.shape == (2, 2, 94)                              # bin/splotch_vinit:194 num_warmup=150
    # the chains start at Pathfinder draws: far better than uniform(-2, 2) starts
    b = _Batch(_Model({k: v for k, v in rdump.read_rdump(cli.data_path(data_dir, 2)).items()
                       if k not in rdump.GENE_KEYS}, "comp_splotch_stan_model"),
               *[np.stack([rdump.read_rdump(cli.data_path(data_dir, g))[k] for g in (2, 3)])
                 for k in ("counts", "beta_prior_mean", "beta_prior_std")])
    rng = np.random.default_rng(0)
    for g in range(2):
        lp0 = b.log_prob_grad(init[g], gene_of=[g, g], want_grad=False)
        rand = b.log_prob_grad(rng.uniform(-2, 2, (40, 94)), gene_of=[g] * 40, want_grad=False)
        assert lp0.min() > np.median(rand) and np.isfinite(lp0).all()
    for g in (2, 3):
        lines = open(os.path.join(out, "0", "combined_%d.csv" % g)).read().splitlines()
        assert lines[0].startswith("lp__,accept_stat__,stepsize__") and len(lines) == 1 + 2 * 12
    # plain `splotch` keeps num_warmup = num_samples and random starts (bin/splotch:131)
    assert cli.main(argv[:-2]) == 0
    assert host_only.started[-1][0] == 12 and host_only.started[-1][1] is None


def test_resume_skips_finished_genes_and_report_lines(host_only, golden_dir, tmp_path):
    """File-level resume ("skip genes whose combined_<g>.* exists") and the per-batch JSON report."""
    data_dir = _c1_data_dir(golden_dir, tmp_path, (1, 2, 3))
    out, rep = str(tmp_path / "out"), str(tmp_path / "report.jsonl")
    common = ["-d", data_dir, "-o", out, "-b", "comp_splotch_stan_model", "-n", "16", "-c", "2", "--seed", "9"]
    assert cli.main(["-g", "2"] + common) == 0
    first = open(os.path.join(out, "0", "combined_2.csv")).read()
    n0 = len(host_only.started)
    assert cli.main(["-g", "1-3", "--resume", "--report", rep] + common) == 0
    # gene 2 was not sampled again: two device batches (genes 1 and 3 are no longer contiguous), its file untouched
    assert [s[1] for s in host_only.started[n0:]] == [None, None] and len(host_only.started) == n0 + 2
    assert open(os.path.join(out, "0", "combined_2.csv")).read() == first
    assert all(os.path.exists(os.path.join(out, "0", "combined_%d.csv" % g)) for g in (1, 2, 3))
    assert cli.main(["-g", "1-3", "--resume", "--report", rep] + common) == 0 and len(host_only.started) == n0 + 2
    lines = [json.loads(l) for l in open(rep)]
    assert [l["genes"] for l in lines] == [[1, 1], [3, 3]]
    for l in lines:
        assert l["n_genes"] == 1 and l["chains"] == 2 and l["num_samples"] == 16 and l["failed_genes"] == []
        assert l["grad_evals"] > 0 and sum(l["treedepth_hist"]) == 32 and 0 < l["mean_stepsize"] < 10
        assert l["ess_genes"] == 1 and l["median_min_bulk_ess"] > 0 and l["median_max_rhat"] > 0.9
    # summary mode looks for the summary file, not the CSV
    assert not cli.output_exists(out, 2, summary=True) and cli.output_exists(out, 2, summary=False)


def test_shim_pathfinder_then_reference_extraction_then_sample(host_only, golden_dir, tmp_path):
    path = os.path.join(golden_dir, "c1", "data_1.R")
    pf_csv = str(tmp_path / "output_1.csv")
    argv = ["/x/comp_splotch_stan_model", "pathfinder", "num_draws=30", "num_paths=2", "max_lbfgs_iters=40",
            "data", "file=" + path, "output", "file=" + pf_csv]
    assert cli.cmdstan_shim(argv) == 0
    # --- the reference's extraction of NUM_CHAINS random draws (bin/splotch_vinit:146-182), restated
    first_line, draw_lines, line_count = None, [], 0
    draws = np.random.default_rng(3).choice(30, size=2, replace=False)
    with open(pf_csv) as fh:
        for line in fh:
            if not line.startswith("#"):
                if first_line is None:
                    first_line = line
                else:
                    if line_count in draws:
                        draw_lines.append(line)
                    line_count += 1
    assert line_count == 30 and len(draw_lines) == 2
    header = first_line.split(",")
    assert header[:3] == stancsv.PATHFINDER_COLS
    for i, last_line in enumerate(draw_lines):
        init = {k: float(v) for k, v in zip(header, last_line.split(",")) if k not in ["lp_approx__", "lp__", "path__"]}
        p = str(tmp_path / ("init_1_%d.json" % (i + 1)))
        json.dump(init, open(p, "w"), indent=2)
        out = str(tmp_path / ("output_1_%d.csv" % (i + 1)))
        argv = ["/x/comp_splotch_stan_model", "sample", "num_samples=8", "num_warmup=150", "random", "id=%d" % (i + 1),
                "init=" + p, "data", "file=" + path, "output", "file=" + out, "refresh=10"]
        assert cli.cmdstan_shim(argv) == 0
        nw, th0 = host_only.started[-1]
        assert nw == 150 and th0.shape == (1, 1, 94)
        # the starting point is the Pathfinder draw of that CSV line (constrained -> unconstrained round trip)
        vals = np.array([float(v) for v in last_line.split(",")[3:]])
        o = oracle.Oracle(rdump.read_rdump(path), "comp_splotch_stan_model")
        assert np.abs(o.write_array(th0[0, 0]) - vals).max() < 1e-9 * max(1.0, np.abs(vals).max())
        assert len(open(out).read().splitlines()) == 9
    # a method the executable does not have
    assert cli.cmdstan_shim(["/x/comp_splotch_stan_model", "optimize", "data", "file=" + path]) == 1


def test_cli_shards_are_disjoint_and_shard_independent(host_only, golden_dir, tmp_path, monkeypatch):
    """One process per GPU under torchrun (RANK / WORLD_SIZE): each rank samples its contiguous block of the gene list and
    writes its own files; a gene's draws do not depend on the number of ranks (RNG keyed by the gene id)."""
    data_dir = _c1_data_dir(golden_dir, tmp_path, (1, 2, 3, 4, 5))
    common = ["-g", "1-5", "-d", data_dir, "-b", "comp_splotch_stan_model", "-n", "8", "-c", "2", "--seed", "21"]
    one = str(tmp_path / "one")
    assert cli.main(common + ["-o", one]) == 0
    two = str(tmp_path / "two")
    owners = {}
    for rank in (0, 1):
        monkeypatch.setenv("WORLD_SIZE", "2")
        monkeypatch.setenv("RANK", str(rank))
        monkeypatch.setenv("LOCAL_RANK", str(rank))
        before = set(os.listdir(os.path.join(two, "0"))) if os.path.isdir(os.path.join(two, "0")) else set()
        assert cli.main(common + ["-o", two]) == 0
        for f in set(os.listdir(os.path.join(two, "0"))) - before:
            owners[f] = rank
    assert sorted(owners) == ["combined_%d.csv" % g for g in range(1, 6)]
    assert [owners["combined_%d.csv" % g] for g in range(1, 6)] == [0, 0, 0, 1, 1]        # contiguous blocks, sizes 3 + 2
    for g in range(1, 6):
        assert open(os.path.join(one, "0", "combined_%d.csv" % g)).read() == \
            open(os.path.join(two, "0", "combined_%d.csv" % g)).read()


def test_pystan_like_interface_semantics(host_only, golden_dir):
    """simulation/fitting.py:35-51: `StanModel(file).sampling(data, iter, chains)` — `iter` counts warm-up + sampling and
    warm-up defaults to iter // 2; `extract(permuted=True)['beta_level_1']` has the draw axis first."""
    data = rdump.read_rdump(os.path.join(golden_dir, "c1", "data_7.R"))
    fit = api.SplotchModel(file="/somewhere/stan/comp_splotch_stan_model.stan").sampling(data=data, iter=40, chains=3, seed=2,
                                                                                         keep_draws=True)
    nw, init = host_only.started[-1]
    assert nw == 20 and init is None
    ex = fit.extract(permuted=True)
    assert ex["beta_level_1"].shape == (3 * 20, 1, 14, 5) and ex["sigma"].shape == (60,) and ex["lp__"].shape == (60,)
    assert ex["log_lambda"].shape == (60, 10) and ex["psi"].shape == (60, 10)
    assert set(fit.extract(pars=["tau", "a"])) == {"tau", "a"}
    # the constrained values are those of write_array on the same draws (oracle): beta_level_1 = raw * std + mean
    s = fit.sampler
    wa = s.batch.write_array(s.theta_draws()[0].reshape(-1, s.model.dim))
    cols = s.model.column_names()
    j = cols.index("beta_level_1.1.3.2")
    assert np.allclose(ex["beta_level_1"][:, 0, 2, 1], wa[:, j], rtol=1e-12, atol=1e-12)
    assert np.allclose(ex["tau"], wa[:, cols.index("tau.1")], rtol=1e-12)
    assert np.allclose(ex["theta"], wa[:, cols.index("theta.1")], rtol=1e-12)
    d = fit.sampler_diagnostics()
    assert d["treedepth__"].shape == (60,) and d["treedepth__"].max() <= 10
    fit = api.SplotchModel(model_name="comp_splotch_stan_model").sampling(data=data, iter=30, warmup=25, chains=1)
    assert host_only.started[-1][0] == 25 and fit.extract()["sigma"].shape == (5,)


@pytest.mark.parametrize("family", ["splotch_stan_model", "comp_splotch_stan_model", "splotch_stan_model_csr"])
def test_extract_small_equals_write_array_on_three_levels(host_only, family):
    """The host-side constraining of the non-spot block (`Sampler.extract_small`: hierarchical betas from beta_raw,
    sigma_level_k and the level maps; stan/*.stan transformed parameters) against the oracle's write_array columns."""
    from csplotch_b200 import synth
    shared = synth.shared_tiny(seed=5, family=family, n_levels=3, zi=1, car=1)
    genes = synth.simulate_counts(shared, 2, 6)
    if family == "comp_splotch_stan_model":
        rng = np.random.default_rng(1)
        genes["beta_prior_mean"] = rng.normal(0, 1, genes["beta_prior_mean"].shape)
        genes["beta_prior_std"] = rng.uniform(0.5, 3, genes["beta_prior_std"].shape)
    model = api.Model(shared, family)
    batch = api.Batch(model, genes["counts"], genes.get("beta_prior_mean"), genes.get("beta_prior_std"))
    s = api.Sampler(batch, api.nuts_cfg(6, 5, 2, seed=3, keep_draws=True)).run()
    ex = s.extract_small()
    cols = model.column_names()
    for g in range(2):
        wa = batch.write_array(s.theta_draws()[g].reshape(-1, model.dim), gene_of=np.full(10, g))
        got = stancsv_like(ex, g)
        for name, arr in got.items():
            j = cols.index(name)
            assert np.allclose(arr, wa[:, j], rtol=1e-12, atol=1e-12), (family, name)


def stancsv_like(ex, g):
    """{dotted column name: draws} of every beta_level_* / scalar entry of extract_small for gene g."""
    out = {}
    for key, arr in ex.items():
        if key in ("diag", "n_draws"):
            continue
        a = arr[g]
        if a.ndim == 1:
            out[key if key == "sigma" else key + ".1"] = a
            continue
        for idx in np.ndindex(*a.shape[1:]):
            out[key + "." + ".".join(str(i + 1) for i in idx)] = a[(slice(None),) + idx]
    return out


def test_summary_mode_equals_summary_of_the_csv(host_only, golden_dir, tmp_path):
    """`-s` (summaries built per gene from the raw draws of the batch, no CSV) against `summarize(read_stan_csv(csv))` of
    the same run: the dictionary splotch_summarize_output would have written (utils.py:234-257), gene by gene."""
    data_dir = _c1_data_dir(golden_dir, tmp_path, (1, 2, 3))
    common = ["-g", "1-3", "-d", data_dir, "-b", "comp_splotch_stan_model", "-n", "15", "-c", "2", "--seed", "8"]
    assert cli.main(common + ["-o", str(tmp_path / "csv")]) == 0
    assert cli.main(common + ["-o", str(tmp_path / "sum"), "-s"]) == 0
    for g in (1, 2, 3):
        post = stancsv.read_stan_csv(os.path.join(str(tmp_path / "csv"), "0", "combined_%d.csv" % g), stancsv.SUMMARY_VARS)
        want = stancsv.summarize(post)
        got = stancsv.read_summary(os.path.join(str(tmp_path / "sum"), "0", "combined_%d.hdf5" % g))
        assert set(got) == set(want)
        for key in want:
            for stat in want[key]:
                assert got[key][stat].shape == want[key][stat].shape, (key, stat)
                assert np.allclose(got[key][stat], want[key][stat], rtol=1e-10, atol=1e-12), (g, key, stat)


def test_extract_small_of_selected_genes(host_only, golden_dir):
    datas = [rdump.read_rdump(os.path.join(golden_dir, "c1", "data_%d.R" % g)) for g in (1, 2, 3, 4)]
    shared = {k: v for k, v in datas[0].items() if k not in rdump.GENE_KEYS}
    model = api.Model(shared, "comp_splotch_stan_model")
    batch = api.Batch(model, *[np.stack([d[k] for d in datas]) for k in rdump.GENE_KEYS])
    s = api.Sampler(batch, api.nuts_cfg(5, 6, 2, seed=1)).run()
    full = s.extract_small()
    part = s.extract_small(genes=[3, 1], raw=s.draws())
    for key in full:
        assert np.array_equal(part[key], full[key][[3, 1]]), key
