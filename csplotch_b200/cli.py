"""`splotch` — same flags as the reference wrapper (/root/reference/bin/splotch:9-29):

    splotch -g GENE_IDX -d DATA_DIRECTORY -o OUTPUT_DIRECTORY -b BINARY -n NUM_SAMPLES -c NUM_CHAINS [-s]

`-b BINARY`: the basename (splotch_stan_model | comp_splotch_stan_model | splotch_stan_model_csr) selects
the model family; no CmdStan executable is run.  Additive: `-g` also accepts ranges / lists
("1-2000", "3,7,9-12") so one process samples thousands of genes at once on the GPU, and `--seed`.  Several GPUs:
launch one process per device with torchrun (`python -m torch.distributed.run --nproc-per-node N -m csplotch_b200.cli
...`); every rank takes its block of the gene list (WORLD_SIZE / RANK / LOCAL_RANK), no collective.

Inputs are the unchanged `DATA/<g//100>/data_<g>.R` files of splotch_generate_input_files; outputs are
`OUT/<g//100>/combined_<g>.csv` (header + draws, chain-major, no comments) or, with -s,
`combined_<g>.hdf5` (summary layout of summarize_stan_hdf5; .npz with the same keys without h5py).

`splotch_vinit` (`main_vinit`) — same flags as /root/reference/bin/splotch_vinit:15-36: multi-path Pathfinder of every
gene (`pathfinder.py`, densities on the GPU), NUM_CHAINS of its draws as starting points, 150 warm-up transitions
(bin/splotch_vinit:137-194) — without the CSV -> JSON -> init= round trip.

`cmdstan_shim` accepts the argv contract of the CmdStan executable itself
(`BINARY sample num_samples=N num_warmup=N random id=CHAIN data file=IN.R output file=OUT.csv`) so that the
unmodified reference `bin/splotch` also works with `-b <shim named like the model>`; it also accepts
`BINARY pathfinder num_draws=M num_paths=C data file=IN.R output file=OUT.csv` (bin/splotch_vinit:140).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from . import rdump, stancsv


def parse_genes(spec: str):
    out = []
    for part in str(spec).split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part[1:]:
            a, b = part.split("-", 1)
            out.extend(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


def data_path(data_dir: str, g: int) -> str:
    return os.path.join(data_dir, str(g // 100), "data_%d.R" % g)


def load_genes(data_dir: str, gene_ids):
    """Shared covariates parsed once (first file); `counts` / beta priors of every gene by the library's
    native tail reader, all host cores (csplotch_rdump_read_genes)."""
    first = rdump.read_rdump(data_path(data_dir, gene_ids[0]))
    shared = {k: v for k, v in first.items() if k not in rdump.GENE_KEYS}
    n_spots = int(np.sum(first["N_spots"]))
    K = int(np.size(first["beta_prior_mean"])) if "beta_prior_mean" in first else 0
    genes = rdump.read_genes_native([data_path(data_dir, g) for g in gene_ids], n_spots, K)
    return shared, genes


VINIT_NUM_DRAWS = 1000     # NUM_INIT, bin/splotch_vinit:137
VINIT_NUM_WARMUP = 150     # bin/splotch_vinit:194


def sample_genes(shared, genes, family, gene_ids, out_dir, num_samples, num_chains, summary, seed=0,
                 num_warmup=None, device=0, vinit=None, sig_figs=0, report=None):
    """`vinit`: None, or a dict of `pathfinder.pathfinder` options — every chain then starts at a Pathfinder draw of
    its gene (bin/splotch_vinit:137-182) instead of uniform(-2, 2)."""
    from . import api, _lib
    api.check(_lib.lib().csplotch_set_device(device))
    model = api.Model(shared, family)
    fam = api.family_from_binary(family)
    batch = api.Batch(model, genes["counts"], genes.get("beta_prior_mean") if fam == 1 else None,
                      genes.get("beta_prior_std") if fam == 1 else None, gene_id0=int(gene_ids[0]))
    nw = num_samples if num_warmup is None else num_warmup          # bin/splotch:131 num_warmup=num_samples
    cfg = api.nuts_cfg(nw, num_samples, num_chains, seed=seed, keep_draws=not summary)
    theta_init, pf_failed, pf_rep = None, np.zeros(len(gene_ids), dtype=bool), None
    if vinit is not None:
        import time
        from . import pathfinder
        opts = dict(num_paths=num_chains, num_draws=VINIT_NUM_DRAWS, num_out=num_chains, seed=seed)
        opts.update(vinit)
        t0 = time.time()
        pf = pathfinder.pathfinder(pathfinder.BatchTarget(batch), **opts)
        pf_failed = ~np.isfinite(pf.lp).all(axis=1)
        theta_init = pf.draws
        pf_rep = {"seconds": time.time() - t0, "lp_grad_evals": int(pf.n_lp_grad), "lp_evals": int(pf.n_lp),
                  "median_lbfgs_iters": float(np.median(pf.lbfgs_iters)), "median_best_iter": float(np.median(pf.best_iter)),
                  "median_pareto_k": float(np.median(pf.pareto_k)), "failed_genes": [int(g) for g, f in zip(gene_ids, pf_failed) if f]}
    s = api.Sampler(batch, cfg, theta_init=theta_init).run(chunk=50)
    status, iters = s.status()
    cols = model.column_names()
    written = []
    diag, _, _ = s.draws()
    good = []
    for i, g in enumerate(gene_ids):
        if pf_failed[i]:
            print("Error: Pathfinder run failed! (gene %d)" % g, file=sys.stderr)   # bin/splotch_vinit:140
            continue
        if (status[i] != 0).any():
            # a failed chain: like a failed CmdStan run, no output file for this gene (error_exit, bin/splotch:131)
            print("Error: sampling failed for gene %d (status %s)" % (g, status[i].tolist()), file=sys.stderr)
            continue
        os.makedirs(os.path.join(out_dir, str(g // 100)), exist_ok=True)
        good.append(i)
    if summary:
        for i in good:
            g = gene_ids[i]
            written.append(stancsv.write_summary(os.path.join(out_dir, str(g // 100), "combined_%d.hdf5" % g),
                                                 stancsv.summary_from_sampler(s, i)))
    else:
        # write_array on the device and the native CSV writer, a block of genes at a time (<= 1 GB of rows on the host)
        rows = num_chains * num_samples
        step = max(1, min(os.cpu_count() or 1, (1 << 30) // max(1, 8 * rows * model.n_outputs)))
        for a0 in range(0, len(good), step):
            blk = good[a0:a0 + step]
            # the draws of this block only (the whole batch at once can exceed host memory at c3-scale dimensions)
            th = np.concatenate([s.theta_draws(g0=i, n_genes=1) for i in blk]).reshape(-1, model.dim)
            vals = batch.write_array(th, gene_of=np.repeat(np.asarray(blk, dtype=np.int32), rows))
            paths = [os.path.join(out_dir, str(gene_ids[i] // 100), "combined_%d.csv" % gene_ids[i]) for i in blk]
            stancsv.write_combined_csvs(paths, cols, diag[blk].reshape(len(blk), rows, 7),
                                        vals.reshape(len(blk), rows, -1), sig_figs=sig_figs)
            written += paths
    if report:
        import json
        with open(report, "a") as f:
            rep = batch_report(s, gene_ids, status)
            if pf_rep is not None:
                rep["pathfinder"] = pf_rep
            f.write(json.dumps(rep) + "\n")
    # genes that wrote no combined_<g>.* file: the caller exits non-zero like the reference's error_exit (bin/splotch:131)
    s.failed_genes = [int(g) for i, g in enumerate(gene_ids) if i not in set(good)]
    return written, s


def batch_report(s, gene_ids, status, ess_genes=8):
    """One JSON-able dict per device batch (SURVEY.md §5 metrics: the reference only has CmdStan's progress lines):
    work done, sampler health, and bulk-ESS / split-R-hat of the betas and scalar parameters of the first genes."""
    from . import diagnostics as dg
    raw = s.draws()
    diag = raw[0]
    cnt = s.counters()
    nch, ns = s.nch, s.cfg.num_samples
    ok = (np.asarray(status) == 0).all(axis=1)
    d = diag[ok].reshape(-1, 7) if ok.any() else np.zeros((0, 7))
    rep = {"genes": [int(gene_ids[0]), int(gene_ids[-1])], "n_genes": len(gene_ids), "chains": int(nch),
           "num_warmup": int(s.cfg.num_warmup), "num_samples": int(ns), "device_ms": float(getattr(s, "total_ms", 0.0)),
           "grad_evals": int(cnt["n_grad"]), "leapfrogs": int(cnt["n_leapfrog"]),
           "failed_genes": [int(g) for g, o in zip(gene_ids, ok) if not o]}
    if len(d):
        rep.update(divergent_frac=float(d[:, 5].mean()), mean_treedepth=float(d[:, 3].mean()),
                   treedepth_hist=np.bincount(d[:, 3].astype(int), minlength=11).tolist(),
                   mean_stepsize=float(d[:, 2].mean()), mean_accept_stat=float(d[:, 1].mean()))
    if ess_genes and ns >= 8 and ok.any():
        pick = np.nonzero(ok)[0][:ess_genes]
        ex = s.extract_small(genes=pick, raw=raw)
        keys = [k for k in ex if k not in ("diag", "n_draws")]
        ess, rhat = [], []
        for i in range(len(pick)):
            cols = np.concatenate([ex[k][i].reshape(nch * ns, -1) for k in keys], axis=1).reshape(nch, ns, -1)
            e, r = dg.summarize(cols)
            ess.append(e)
            rhat.append(r)
        rep.update(ess_genes=len(ess), median_min_bulk_ess=float(np.median(ess)), median_max_rhat=float(np.median(rhat)))
    return rep


def output_exists(out_dir, g, summary):
    """A finished gene: outputs are written to a temporary name and renamed, so existence means complete."""
    base = os.path.join(out_dir, str(g // 100), "combined_%d" % g)
    return any(os.path.exists(base + ext) for ext in ((".hdf5", ".hdf5.npz") if summary else (".csv",)))


def main_vinit(argv=None):
    """`splotch_vinit`: Pathfinder-initialised chains and a 150-transition warm-up."""
    return main(argv, vinit=True)


def main(argv=None, vinit=False):
    ap = argparse.ArgumentParser(prog="splotch_vinit" if vinit else "splotch",
                                 description="A wrapper for running Splotch with Pathfinder initialization "
                                             "(B200-native)" if vinit else
                                             "A wrapper for running Splotch (B200-native sampler)")
    ap.add_argument("-g", required=True, help="gene index (1-based); ranges/lists allowed: 1-100,205")
    ap.add_argument("-d", required=True, help="data directory")
    ap.add_argument("-o", required=True, help="output directory")
    ap.add_argument("-b", required=True, help="Splotch binary (its basename selects the model)")
    ap.add_argument("-n", required=True, type=int, help="number of samples")
    ap.add_argument("-c", required=True, type=int, help="number of chains")
    ap.add_argument("-s", action="store_true", help="store summary rather than full output")
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--batch", type=int, default=512, help="genes resident on the GPU at once")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--sig-figs", type=int, default=0,
                    help="significant digits in the CSV (CmdStan's sig_figs; 0 = shortest text that round-trips)")
    ap.add_argument("--resume", action="store_true",
                    help="skip genes whose combined_<g>.* already exists in the output directory")
    ap.add_argument("--report", default=None, help="append one JSON line per device batch (work, sampler health, "
                                                   "ESS / R-hat of the first genes) to this file")
    if vinit:
        ap.add_argument("--max-lbfgs-iters", type=int, default=1000, help="Pathfinder: L-BFGS iterations per path")
    a = ap.parse_args(argv)
    gene_ids = parse_genes(a.g)
    seed = a.seed if a.seed is not None else (int.from_bytes(os.urandom(4), "little"))   # CmdStan seeds from the clock
    if not os.path.isdir(a.d):
        print("Error: data directory does not exist!", file=sys.stderr)
        return 1
    try:
        os.makedirs(a.o, exist_ok=True)
    except OSError:
        print("Error: Cannot create the output directory!", file=sys.stderr)
        return 1
    # one process per GPU under torchrun: this rank's contiguous block of the gene list, no collective
    from . import shard
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if world > 1:
        gene_ids = shard.my_genes(sorted(gene_ids), world, rank)
        a.device = int(os.environ.get("LOCAL_RANK", str(rank)))
        if not gene_ids:
            return 0
    if a.resume:
        gene_ids = [g for g in gene_ids if not output_exists(a.o, g, a.s)]
        if not gene_ids:
            return 0
    # genes resident at once: bounded by device memory.  Per gene: 4 D-vectors + 6 N summary doubles per
    # chain, and (CSV mode only) every sampling draw of the full unconstrained vector.
    try:
        head = rdump.read_rdump(data_path(a.d, gene_ids[0]), keys=("N_spots", "car"))
        n_spots = int(np.sum(head["N_spots"]))
        dim = (2 if int(head.get("car", 1)) else 1) * n_spots + 4096
        per_gene = a.c * 8 * (4 * dim + 6 * n_spots + (0 if a.s else a.n * dim))
        a.batch = max(1, min(a.batch, int(float(os.environ.get("CSPLOTCH_MEM_GB", "60")) * 1e9 // per_gene)))
        if vinit:
            # host side of Pathfinder: ~45 D-vectors per gene-path (iterate, gradient, 2 x 5 history pairs, the
            # factors of the current and the best approximation, line-search candidates)
            a.batch = max(1, min(a.batch, int(float(os.environ.get("CSPLOTCH_HOST_GB", "16")) * 1e9
                                              // (a.c * 8 * dim * 45))))
    except (OSError, KeyError):
        pass  # missing first file: the per-run loader reports it
    # contiguous runs keep gene_id0 + i == gene id (the RNG stream is keyed by the gene index)
    runs = shard.contiguous_runs(gene_ids, a.batch)
    rc = 0
    for run in runs:
        try:
            shared, genes = load_genes(a.d, run)
            if vinit:
                _, smp = sample_genes(shared, genes, a.b, run, a.o, a.n, a.c, a.s, seed=seed, device=a.device,
                                      num_warmup=VINIT_NUM_WARMUP, vinit=dict(max_lbfgs_iters=a.max_lbfgs_iters),
                                      sig_figs=a.sig_figs, report=a.report)
            else:
                _, smp = sample_genes(shared, genes, a.b, run, a.o, a.n, a.c, a.s, seed=seed, device=a.device,
                                      sig_figs=a.sig_figs, report=a.report)
            if getattr(smp, "failed_genes", None):
                rc = 1     # a gene without output is a failed run for job arrays and schedulers
        except Exception as e:  # noqa: BLE001
            print("Error: Stan run failed! (%s)" % e, file=sys.stderr)
            rc = 1
    return rc


def cmdstan_shim(argv=None, family=None):
    """argv contract of the CmdStan executable (bin/splotch:131, bin/splotch_vinit:194)."""
    argv = list(sys.argv if argv is None else argv)
    family = family or os.path.basename(argv[0])
    args = argv[1:]
    if not args or args[0] not in ("sample", "pathfinder"):
        print("only the `sample` and `pathfinder` methods are supported", file=sys.stderr)
        return 1
    method = args[0]
    kv = {}
    section = None
    for tok in args[1:]:
        if "=" in tok:
            k, v = tok.split("=", 1)
            kv[(section + "." + k) if section in ("data", "output", "random") and k in ("file", "id", "seed") else k] = v
        else:
            section = tok
    num_samples = int(kv.get("num_samples", 1000))
    num_warmup = int(kv.get("num_warmup", 1000))
    chain_id = int(kv.get("random.id", kv.get("id", 1)))
    seed = int(kv.get("random.seed", kv.get("seed", int.from_bytes(os.urandom(4), "little"))))
    if "data.file" not in kv:
        print("%s: `data file=<data_<g>.R>` is required" % method, file=sys.stderr)
        return 1
    try:
        data = rdump.read_rdump(kv["data.file"])
    except OSError as e:
        print("Error reading data file %s: %s" % (kv["data.file"], e), file=sys.stderr)
        return 1
    from . import api
    shared = {k: v for k, v in data.items() if k not in rdump.GENE_KEYS}
    model = api.Model(shared, family)
    fam = api.family_from_binary(family)
    batch = api.Batch(model, np.asarray(data["counts"])[None], data.get("beta_prior_mean") if fam == 1 else None,
                      data.get("beta_prior_std") if fam == 1 else None)
    out = kv.get("output.file", "output.csv")
    sig_figs = max(0, int(kv.get("sig_figs", 0)))       # CmdStan: output sig_figs=N (-1 = its default)
    if method == "pathfinder":
        # bin/splotch_vinit:140; CmdStan's output: lp_approx__, lp__, path__, then the write_array columns
        from . import pathfinder
        opt = {k: int(kv[k]) for k in ("num_paths", "num_draws", "history_size", "num_elbo_draws", "max_lbfgs_iters")
               if k in kv}
        opt.setdefault("num_paths", 4)
        opt.setdefault("num_draws", 1000)
        pf = pathfinder.pathfinder(pathfinder.BatchTarget(batch), seed=seed, **opt)
        if not np.isfinite(pf.lp[0]).all():
            print("Pathfinder failed: no path produced a usable approximation", file=sys.stderr)
            return 1
        vals = batch.write_array(pf.draws[0], gene_of=np.zeros(pf.draws.shape[1], dtype=np.int32))
        lead = np.stack([pf.lp_approx[0], pf.lp[0], pf.path[0].astype(np.float64)], axis=1)
        stancsv.write_combined_csv(out, model.column_names(), lead, vals, lead_cols=stancsv.PATHFINDER_COLS,
                                   sig_figs=sig_figs)
        return 0
    # one chain per process like CmdStan; chain ids start at 1, so run chain_id chains' worth of keying
    cfg = api.nuts_cfg(num_warmup, num_samples, 1, seed=seed + 7919 * (chain_id - 1), keep_draws=True)
    # init=FILE.json: constrained starting values (bin/splotch_vinit:194); init=R: radius of the uniform start
    theta_init = None
    if "init" in kv:
        try:
            cfg.init_radius = float(kv["init"])
        except ValueError:
            import json
            try:
                with open(kv["init"]) as f:
                    theta_init = api.unconstrain(model, json.load(f), radius=cfg.init_radius, seed=cfg.seed)[None, None, :]
            except (OSError, ValueError, api.CsplotchError) as e:
                print("Error reading init file %s: %s" % (kv["init"], e), file=sys.stderr)
                return 1
    s = api.Sampler(batch, cfg, theta_init=theta_init).run(chunk=50)
    st, _ = s.status()
    if (st != 0).any():
        return 1
    diag, _, _ = s.draws()
    th = s.theta_draws()[0, 0]
    vals = batch.write_array(th, gene_of=np.zeros(len(th), dtype=np.int32))
    stancsv.write_combined_csv(out, model.column_names(), diag[0, 0], vals, sig_figs=sig_figs)
    return 0


if __name__ == "__main__":
    sys.exit(main())
