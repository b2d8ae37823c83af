#!/usr/bin/env python
"""bench.py — gradient evaluations / second of batched NUTS over genes x chains; genes / hour to bulk-ESS >= 400
(BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c5|c2|c4] [--impl reference]

Default workload (BASELINE configs[2] / configs[4]): ONE fixed gene list on the c3 covariates (synthetic Visium
24 x 4 992 spots, comp_splotch_stan_model, 3-level, ZIP+CAR, K = 10; D = 242 142), 592 genes x 4 chains.  At N = 1
the GPU owns the whole list ("c3"); at N > 1 the same list is dealt over the ranks, longest genes first
(csplotch_b200.shard.deal; "c5": the sharded sweep) => strong scaling, no data-path collective, time = max over ranks.
`--workload c2|c4` run those shapes as per-GPU slices (weak scaling) like round 1.

One "step" advances EVERY gene-chain of the list by `--iters` NUTS transitions (one launch of the persistent kernel
per rank).  W untimed steps, then exactly K timed steps bracketed by a barrier + device synchronise; device time from
CUDA events on the sampler's stream, max over ranks.

  value      gradient evaluations / s, whole job (all ranks), inputs resident in HBM
  e2e        same metric through the host-buffer C ABI: counts H2D, sampler create + advance, draws and
             spot summaries D2H, all inside the timed region
  roofline   dominant kernel k_nuts_advance: algorithmic bytes (B_leap = 56 D + 4 N per
             in-sampler leapfrog, B_grad = 16 D + 4 N per other gradient) / CUDA-event duration vs the
             measured HBM peak (MEASURED_PEAKS.json); fp64_frac = fp64 flops per gradient / measured DFMA peak
  ess        a complete `splotch -n 200 -c 4` job of 16 genes per GPU: genes / hour, genes / hour to bulk-ESS >= 400
             (all parameters, and betas only), minutes for a 20 000-gene transcriptome at that rate
  secondary  (N = 1) the c2 shape (BASELINE configs[1]) resident-arm rate and roofline fraction
  cpu_baseline  the CPU oracle (restatement of the Stan models + Stan's NUTS; CmdStan itself is not
             installable here), -O3 -march=native timing build, all host cores, bounded sample, rank 0 at N=1

`--impl reference` times that same CPU implementation (all host threads) on this arm's config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grad_evals_per_sec"
UNIT = "grad-evals/s"

WORKLOADS = {
    # name: (config key, genes (total for the strong list, per GPU for the weak ones), iterations per step, description)
    "c3": ("c3", 592, 1, "c3: synthetic Visium 24x4992 spots, comp_splotch_stan_model 3-level ZIP+CAR K=10, "
                         "D=242142, fixed list of genes x 4 chains"),
    "c5": ("c3", 592, 1, "c5: the c3 gene list (synthetic Visium 24x4992, comp 3-level ZIP+CAR K=10, D=242142) sharded "
                         "over the GPUs, longest genes first, x 4 chains"),
    "c2": ("c2", 11313, 2, "c2: examples/ ST-v1 geometry, splotch_stan_model 2-level Poisson+CAR, N=2594 spots "
                           "(39 tissue sections of the 10 example arrays), D=5258, all 11313 genes x 4 chains"),
    "c4": ("c4", 296, 1, "c4: synthetic Visium-HD 4x317^2 bins, splotch_stan_model_csr 2-level ZIP+CAR, D=803989, "
                         "gene slice x 4 chains"),
}
STRONG = ("c3", "c5")      # one fixed gene list for every N


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(workload: str, n_genes: int, rank: int, head_only: bool = False):
    """(family, shared, genes).  Strong workloads: the same `n_genes`-gene list on every rank (seeded by the config);
    weak ones: a different slice per rank.  `head_only`: just the first generation chunk of the gene list (the
    generator is chunk-wise deterministic, so these are the same genes as the head of the full list) — all the CPU
    legs ever touch."""
    from csplotch_b200 import synth
    key = WORKLOADS[workload][0]
    cfg = synth.CONFIGS[key]
    shared = cfg["shared"]()
    chunk = max(16, (1 << 22) // int(np.sum(shared["N_spots"])))
    if head_only:
        n_genes = min(n_genes, chunk)
    seed = cfg["seed"] if workload in STRONG else cfg["seed"] + 1000 * rank
    genes = synth.simulate_counts(shared, n_genes, seed, chunk=chunk)
    return cfg["family"], shared, genes


def take_genes(genes: dict, idx):
    return {k: np.ascontiguousarray(v[idx]) for k, v in genes.items()}


def fp64_peak():
    """Measured DFMA roof of this box (tools/fp64_peak, TFLOP/s), or None when the tool is absent."""
    exe = os.path.join(ROOT, "tools", "fp64_peak")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
        return float(json.loads(out.strip().splitlines()[-1])["dfma_tflops"])
    except Exception:
        return None


def ess_leg(api, model, genes, G, n, chains=4, seed=11, ess_genes=64, gene_ids=None):
    """Second half of BASELINE.json's metric: genes / hour to bulk-ESS >= 400.  A complete `splotch -n N -c 4` job
    (N warm-up + N sampling transitions from initialisation, /root/reference/bin/splotch:131) of the first G genes;
    rate = G / device time of the launches, scaled by the fraction of (the first `ess_genes`) genes whose minimum
    rank-normalised bulk-ESS reaches 400 — over every beta_level_* entry and scalar parameter (the metric as
    stated), and over the betas alone (what the reference's downstream Bayes factors use; the variance scalars of
    these models mix slowly in any implementation)."""
    from csplotch_b200 import diagnostics as dg
    pm = genes["beta_prior_mean"][:G] if "beta_prior_mean" in genes else None
    ps = genes["beta_prior_std"][:G] if "beta_prior_std" in genes else None
    batch = api.Batch(model, genes["counts"][:G], pm, ps, gene_ids=None if gene_ids is None else gene_ids[:G])
    s = api.Sampler(batch, api.nuts_cfg(n, n, chains, seed=seed)).run(chunk=50)
    seconds = s.total_ms * 1e-3
    status, _ = s.status()
    ex = s.extract_small(genes=np.arange(min(G, ess_genes)))
    bkeys = [k for k in ex if k.startswith("beta_level")]
    skeys = [k for k in ex if k not in bkeys and k not in ("diag", "n_draws")]
    ess_all, ess_beta, rhat_beta = [], [], []
    for g in range(min(G, ess_genes)):
        if (status[g] != 0).any():
            ess_all.append(0.0); ess_beta.append(0.0); rhat_beta.append(float("inf"))
            continue
        cb = np.concatenate([ex[k][g].reshape(chains * n, -1) for k in bkeys], axis=1).reshape(chains, n, -1)
        cs = np.concatenate([ex[k][g].reshape(chains * n, -1) for k in skeys], axis=1).reshape(chains, n, -1)
        eb, rb = dg.summarize(cb)
        es, _ = dg.summarize(cs)
        ess_all.append(min(eb, es)); ess_beta.append(eb); rhat_beta.append(rb)
    ess_all, ess_beta = np.array(ess_all), np.array(ess_beta)
    gph = G / seconds * 3600.0 if seconds > 0 else 0.0
    cnt = s.counters()
    diag = ex.get("diag")
    out = {"genes": int(G), "chains": int(chains), "num_warmup": int(n), "num_samples": int(n), "device_seconds": seconds,
           "genes_per_hour": gph, "grad_evals": int(cnt["n_grad"]), "grad_evals_per_sec": cnt["n_grad"] / seconds if seconds > 0 else 0.0,
           "genes_evaluated_for_ess": int(len(ess_all)),
           "genes_per_hour_to_ess400": gph * float((ess_all >= 400).mean()), "median_min_bulk_ess": float(np.median(ess_all)),
           "beta_only": {"genes_per_hour_to_ess400": gph * float((ess_beta >= 400).mean()),
                         "median_min_bulk_ess": float(np.median(ess_beta)), "median_max_rhat": float(np.median(rhat_beta))},
           "note": "whole job incl. both phases (warm-up trees are deeper than sampling-phase trees): "
                   "grad_evals_per_sec here is the rate over warm-up + sampling"}
    if diag is not None and diag.size:
        out["sampling_phase_mean_n_leapfrog"] = float(np.mean(diag[..., 4]))
        out["sampling_phase_mean_treedepth"] = float(np.mean(diag[..., 3]))
    if hasattr(s, "close"):
        s.close(); batch.close()
    return out


def cpu_reference_run(family, shared, genes, seconds, n_threads, n_chains=4):
    """Bounded CPU sample: n_threads gene-chains (one thread each, like CmdStan's one process per chain) run Stan's
    adaptive NUTS from initialisation for `seconds` of wall clock, then stop at their next leapfrog (one max-depth
    transition of a c3 chain is ~15 s of a host core, so "the first k transitions" cannot be bounded to seconds).
    Returns (grads, seconds, description, wall seconds): `seconds` is the thread-busy time / threads, i.e.
    grads / seconds is the rate of a long job whose threads never idle.  Uses the timing build of the oracle
    (oracle/Makefile `fast`: -O3 -march=native, rebuilt on this host)."""
    if "CSPLOTCH_ORACLE_LIB" not in os.environ:
        import oracle as _o
        os.environ["CSPLOTCH_ORACLE_LIB"] = _o.build_fast()
    import oracle
    from csplotch_b200 import synth
    G = max(1, n_threads // n_chains)
    G = min(G, genes["counts"].shape[0])
    o = oracle.Oracle(synth.gene_dict(shared, genes, 0), family)
    cfg = oracle.nuts_cfg(1000, 0, seed=1)
    pm = genes["beta_prior_mean"][:G] if "beta_prior_mean" in genes else None
    ps = genes["beta_prior_std"][:G] if "beta_prior_std" in genes else None
    t0 = time.time()
    res = o.nuts_batch(genes["counts"][:G], cfg, n_chains=n_chains, prior_mean=pm, prior_std=ps,
                       n_threads=n_threads, want_draws=False, time_budget=seconds)
    wall = time.time() - t0
    workers = min(n_threads, G * n_chains)
    busy = res.get("busy_seconds", 0.0) / max(1, workers)
    dt = busy if busy > 0 else wall
    return res["n_grad"], dt, ("%d genes x %d chains, adaptive NUTS (max_depth 10) from initialisation for %.1f s of wall clock, "
                               "%d threads, gcc -O3 -march=native; rate = gradients / (thread-busy seconds / threads)"
                               % (G, n_chains, seconds, n_threads)), wall


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path on the host cores.  CmdStan is
    not installable in this image (no stanc / Stan Math / network), so this is the C restatement in
    oracle/ (kind "port"), gene-parallel over all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl_name = args.workload or ("c3" if args.gpus <= 1 else "c5")
    # the same genes as the cpu_baseline leg of the GPU arm: the head of the workload's gene list
    family, shared, genes = make_inputs(wl_name, WORKLOADS[wl_name][1], 0, head_only=True)
    nthr = os.cpu_count() or 1
    # every step is one time-bounded sample; the whole arm (warm-up + steps) takes ~2.5 minutes of host time
    n_runs = max(1, args.steps + args.warmup)
    iters = float(args.cpu_iters) if args.cpu_iters else max(2.0, 150.0 / n_runs)
    for _ in range(args.warmup):
        cpu_reference_run(family, shared, genes, iters, nthr)
    tot_g, tot_t, tot_wall, sample = 0, 0.0, 0.0, ""
    for _ in range(args.steps):
        g, dt, sample, wall = cpu_reference_run(family, shared, genes, iters, nthr)
        tot_g += g; tot_t += dt; tot_wall += wall
    val = tot_g / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_wall / max(1, args.steps), "higher_is_better": True,
        "scaling": "strong" if wl_name in STRONG else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[wl_name][3]},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthr, "kind": "port", "sample": "per step: " + sample,
                         "wall_clock_value": tot_g / tot_wall},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement of the Stan models + Stan NUTS (oracle/), not CmdStan: CmdStan cannot be built here",
    }
    emit(line)
    return 0


_JSON_FD = None


def claim_stdout():
    """stdout carries the ONE JSON line and nothing else: file descriptor 1 is pointed at stderr for the rest of the
    process (NCCL prints its version banner to stdout, libraries may chat), the line goes to the original descriptor."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: c3 at --gpus 1, c5 (the same gene list, sharded) at --gpus > 1")
    ap.add_argument("--genes", type=int, default=0, help="genes in the list (c3/c5) or per GPU (c2/c4)")
    ap.add_argument("--iters", type=int, default=0, help="NUTS transitions per chain per step")
    ap.add_argument("--chains", type=int, default=4)
    ap.add_argument("--cpu-iters", type=int, default=0, help="seconds of wall clock of one CPU sample (default: 20, or 150 / (steps + warmup) per step of --impl reference)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--e2e-genes", type=int, default=0)
    ap.add_argument("--e2e-iters", type=int, default=0, help="NUTS transitions per chain in one e2e step")
    ap.add_argument("--ess-genes", type=int, default=-1,
                    help="genes per GPU of the `splotch -n N -c 4` job behind genes/hour to bulk-ESS >= 400 "
                         "(default 16 at c3: 64 gene-chains, one resident wave of 4-CTA clusters; 0 skips the leg)")
    ap.add_argument("--ess-n", type=int, default=200, help="N of that job (the reference's README uses -n 200)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the c2 sub-record (N = 1)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 3) if os.environ.get("CSPLOTCH_BENCH_STRICT", "1") == "1" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from csplotch_b200 import api, shard, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    api.check(_lib.lib().csplotch_set_device(local))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_start = time.time()

    def note(msg):
        if rank == 0:
            print("[bench %6.1fs] %s" % (time.time() - t_start, msg), file=sys.stderr, flush=True)

    wl_name = args.workload or ("c3" if world == 1 else "c5")
    wl = WORKLOADS[wl_name]
    strong = wl_name in STRONG
    G_list = args.genes or wl[1]
    iters = args.iters or wl[2]
    family, shared, genes_all = make_inputs(wl_name, G_list, rank)
    if strong:
        # one list for every N: dealt longest-first (non-zero counts = cost estimate), ids travel with the genes
        ids_all = np.arange(G_list)
        cost = (genes_all["counts"] > 0).sum(axis=1)
        mine = np.asarray(shard.deal(ids_all, cost, world, rank), dtype=np.int64)
        genes = take_genes(genes_all, mine)
        gene_ids = mine.astype(np.uint32)
    else:
        genes = genes_all
        gene_ids = (np.arange(G_list) + rank * G_list).astype(np.uint32)
    G = genes["counts"].shape[0]
    del genes_all
    note("synthetic inputs ready (%d genes in the list, %d on this rank)" % (G_list if strong else G_list * world, G))
    model = api.Model(shared, family)
    # The resident arm runs one gene-chain per plain CTA whatever the number of chains.  Left to itself the library puts
    # launches of few chains (a strong-scaling share at N = 8: 296) on 4-CTA clusters, which is right for jobs run from
    # initialisation (the e2e and ESS legs below: 104 k against 87 k grad/s, GPU call 24) but not for a launch that
    # advances every chain by ONE transition of comparable depth (141 k against 105 k): the cluster's shorter latency buys
    # nothing there and its serial rank-0 epilogue costs 28 % of the throughput.  CSPLOTCH_CLUSTER is the library's
    # documented knob (read at model creation); a value set by the caller wins.
    model_res = model
    if "CSPLOTCH_CLUSTER" not in os.environ:
        os.environ["CSPLOTCH_CLUSTER"] = "1"
        try:
            model_res = api.Model(shared, family)
        finally:
            del os.environ["CSPLOTCH_CLUSTER"]
    N, D = model.N, model.dim
    B_leap, B_grad = 56 * D + 4 * N, 16 * D + 4 * N

    # ---- resident arm: inputs already in HBM ------------------------------------------------------
    def resident(model_, genes_, ids_, iters_, steps_, warm_, clocks_=None):
        batch = api.Batch(model_, genes_["counts"], genes_.get("beta_prior_mean"), genes_.get("beta_prior_std"),
                          gene_ids=ids_)
        total_iters = (warm_ + steps_) * iters_
        sampler = api.Sampler(batch, api.nuts_cfg(total_iters, 0, args.chains, seed=20240))  # every step in Stan's warm-up phase
        for _ in range(warm_):
            sampler.advance(iters_, sync=True)
        c0 = sampler.counters()
        barrier()
        if clocks_ is not None:
            clocks_.start()
        ms_dev, tails = 0.0, []
        t0 = time.time()
        for _ in range(steps_):
            sampler.advance(iters_, sync=True)
            ms_dev += sampler.last_ms()
            mean_s, max_s = sampler.launch_tail()
            tails.append(1.0 - mean_s / max_s if max_s > 0 else 0.0)
        barrier()
        wall = time.time() - t0
        clk = clocks_.stop() if clocks_ is not None else None
        c1 = sampler.counters()
        prof = sampler.profile()
        sampler.close(); batch.close()
        return dict(ms=ms_dev, wall=wall, grads=c1["n_grad"] - c0["n_grad"], leaps=c1["n_leapfrog"] - c0["n_leapfrog"],
                    launches=c1["n_launches"] - c0["n_launches"], clocks=clk, tail=float(np.mean(tails)), profile=prof,
                    total_iters=total_iters)

    r = resident(model_res, genes, gene_ids, iters, args.steps, args.warmup, ClockSampler(local) if rank == 0 else None)
    if model_res is not model:
        model_res.close()
    ms_dev, grads, leaps, launches, clk, t_wall = r["ms"], r["grads"], r["leaps"], r["launches"], r["clocks"], r["wall"]
    note("resident arm: %d timed steps done (launch tail %.1f %%)" % (args.steps, 100 * r["tail"]))

    # ---- e2e arm: host buffers through the C ABI, copies inside the timed region --------------------
    # One step = one complete small job for a gene block through the public API, the way bin/splotch uses the
    # path: counts of Ge genes (page-locked host memory) -> Batch (H2D) -> Sampler -> warm-up + sampling
    # transitions from initialisation -> draws of the non-spot block + posterior summaries of every spot (D2H
    # into page-locked host arrays).  The page-locked buffers are allocated once, outside the timed region.
    Ge = min(G, args.e2e_genes or (4736 if N < 10000 else 74))
    e2e_iters = args.e2e_iters or (50 if N < 10000 else 20)
    e_warm, e_samp = e2e_iters // 2, e2e_iters - e2e_iters // 2
    counts_host = api.pinned_empty((Ge, N), np.int32)
    counts_host[...] = genes["counts"][:Ge]
    pm = genes["beta_prior_mean"][:Ge] if "beta_prior_mean" in genes else None
    ps = genes["beta_prior_std"][:Ge] if "beta_prior_std" in genes else None
    out_summ = api.pinned_empty((6, Ge, N))
    out_draws = (api.pinned_empty((Ge, args.chains, e_samp, 7)), api.pinned_empty((Ge, args.chains, e_samp, model.n_small)),
                 api.pinned_empty((Ge, args.chains), np.int32))

    def e2e_step():
        b = api.Batch(model, counts_host, pm, ps, gene_ids=gene_ids[:Ge])       # H2D of this step's inputs
        s = api.Sampler(b, api.nuts_cfg(e_warm, e_samp, args.chains, seed=7))
        s.run()
        diag, small, nd = s.draws(out=out_draws)                          # D2H of the step's results
        summ = s.spot_summaries(out=out_summ)
        cnt = s.counters()
        assert int(nd.min()) == e_samp and np.isfinite(summ["lambda/mean"]).all()
        d2h = diag.nbytes + small.nbytes + nd.nbytes + sum(v.nbytes for v in summ.values())
        s.close(); b.close()
        return cnt["n_grad"], counts_host.nbytes + (pm.nbytes + ps.nbytes if pm is not None else 0), d2h, cnt["n_launches"]

    e2e_step()
    note("e2e arm: warm-up step done (%d genes x %d transitions)" % (Ge, e2e_iters))
    barrier()
    t0 = time.time()
    e_grads, h2d, d2h, e_launches = 0, 0, 0, 0
    e_steps = 2        # each step is a complete job from initialisation
    for _ in range(e_steps):
        g, a, b_, nl = e2e_step()
        e_grads += g; h2d = a; d2h = b_; e_launches += nl
    barrier()
    e_dt = time.time() - t0
    note("e2e arm: %d timed steps done" % e_steps)
    del counts_host, out_summ, out_draws

    # ---- genes / hour to bulk-ESS >= 400 (second half of the BASELINE metric) ------------------------------
    ess = None
    ess_genes = args.ess_genes if args.ess_genes >= 0 else (16 if N >= 10000 else 592)
    if ess_genes > 0:
        ess = ess_leg(api, model, genes, min(G, ess_genes), args.ess_n, args.chains, gene_ids=gene_ids)
        note("ESS leg done (%d genes, -n %d, %.1f s)" % (ess["genes"], args.ess_n, ess["device_seconds"]))
        if world > 1:                     # whole-job rate: ranks work on disjoint genes at the same time
            t = torch.tensor([ess["genes_per_hour"], ess["genes_per_hour_to_ess400"],
                              ess["beta_only"]["genes_per_hour_to_ess400"], ess["grad_evals_per_sec"]], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ess["genes_per_hour"], ess["genes_per_hour_to_ess400"] = float(t[0]), float(t[1])
            ess["beta_only"]["genes_per_hour_to_ess400"] = float(t[2])
            ess["grad_evals_per_sec"] = float(t[3])
            ess["note"] += "; rates summed over %d GPUs (each its own %d genes), ESS statistics are rank 0's" % (world, ess["genes"])
        ess["transcriptome_20k_genes_minutes"] = 20000.0 / ess["genes_per_hour"] * 60.0 if ess["genes_per_hour"] > 0 else None

    # ---- secondary record: the c2 shape (BASELINE configs[1]) on this GPU, resident arm only --------------
    secondary = None
    if world == 1 and wl_name != "c2" and not args.no_secondary:
        fam2, shared2, genes2 = make_inputs("c2", 2960, 0)
        model2 = api.Model(shared2, fam2)
        r2 = resident(model2, genes2, np.arange(2960, dtype=np.uint32), 2, 4, 5)
        bl2, bg2 = 56 * model2.dim + 4 * model2.N, 16 * model2.dim + 4 * model2.N
        ach2 = (r2["leaps"] * bl2 + (r2["grads"] - r2["leaps"]) * bg2) / (r2["ms"] * 1e-3) / 1e9
        secondary = {"c2": {"workload": WORKLOADS["c2"][3].replace("all 11313 genes", "2960 genes"), "value": r2["grads"] / (r2["ms"] * 1e-3),
                            "unit": UNIT, "ms_per_step": r2["ms"] / 4, "steps": 4, "warmup": 5,
                            "nuts_transitions_per_chain_per_step": 2, "roofline_achieved_GBs": ach2,
                            "roofline_frac": ach2 / peaks()[0], "launch_tail_frac": r2["tail"],
                            "phase": "Stan warm-up transitions 10..18"}}
        model2.close()
        note("secondary c2 record done")

    # ---- reduce over ranks ---------------------------------------------------------------------------
    tail = r["tail"]
    if world > 1:
        t = torch.tensor([ms_dev, e_dt, tail], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, e_dt, tail = float(t[0]), float(t[1]), float(t[2])
        t = torch.tensor([grads, leaps, e_grads, launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        grads, leaps, e_grads, launches = (float(x) for x in t)
    value = grads / (ms_dev * 1e-3)
    e2e_value = e_grads / e_dt

    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = (leaps * B_leap + (grads - leaps) * B_grad) / max(1, world)   # per GPU
        achieved = alg_bytes / (ms_dev * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "traffic_%s.json" % wl[0])
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            traffic = tj.get("dram_bytes_per_grad", 0) * (grads / max(1, world)) / max(1, args.steps) or None
            traffic_src = ("EXTRAPOLATED, not measured in this run: DRAM bytes per gradient of the stored `ncu --set full` "
                           "capture profiles/traffic_%s.json x gradients per launch" % wl[0])
        # fp64 work per leapfrog: per spot 2K (forward contraction) + 2K (its transpose) + ~150 for the two
        # exp, the reciprocal, drift, CAR and the kicks
        K = model.K
        flops = leaps * N * (150.0 + 4.0 * K) / max(1, world)
        fpk = fp64_peak()
        prof = r["profile"]
        ptot = max(1, sum(prof[k] for k in ("cyc_eval", "cyc_merge", "cyc_copy", "cyc_other")))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl[3], "genes_in_list": G_list if strong else G_list * world,
                       "genes_on_rank0": G, "chains": args.chains,
                       "nuts_transitions_per_chain_per_step": iters,
                       "resident_scope": "one gene-chain per CTA (CSPLOTCH_CLUSTER=1 for the resident arm; e2e and ess legs: the library's own choice)",
                       "sharding": ("one fixed gene list for every N, dealt longest-first over the ranks (shard.deal), "
                                    "no data-path collective") if strong else "a slice of the gene list per GPU",
                       "l2_policy": "inputs larger than L2: %.1f GB of chain state per GPU streamed every step"
                                    % ((4 * model.dim * 8 * G * args.chains) / 1e9),
                       "phase": "Stan warm-up transitions %d..%d of every chain (adaptive NUTS, max_depth 10); the ess "
                                "record covers a whole job incl. the sampling phase" % (args.warmup * iters, r["total_iters"])},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "what": "%d genes x %d chains per step and GPU: Batch(host counts) -> Sampler, %d warm-up + %d sampling "
                            "transitions from initialisation (the deepest trees of a run) -> draws + spot summaries to host, "
                            "through the C ABI; wall clock, %d steps" % (Ge, args.chains, e_warm, e_samp, e_steps)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "k_nuts_advance",
                         "algorithmic_bytes_per_leapfrog": B_leap, "algorithmic_bytes_per_gradient": B_grad,
                         "launches": int(launches / max(1, world)), "avg_launch_ms": ms_dev / max(1, args.steps),
                         "fp64_flops_per_leapfrog": N * (150.0 + 4.0 * K), "fp64_peak_tflops_measured": fpk,
                         "fp64_frac": (flops / (ms_dev * 1e-3) / 1e12 / fpk) if fpk else None,
                         "launch_tail_frac": tail,
                         # the same fraction over the time the resident slots were busy (a production launch advances
                         # chains by tens of transitions and has no such tail): achieved / (1 - tail) / peak
                         "frac_excluding_launch_tail": (achieved / peak / (1.0 - tail)) if tail < 1.0 else None,
                         "time_shares": {k[4:]: prof[k] / ptot for k in ("cyc_eval", "cyc_merge", "cyc_copy", "cyc_other")}},
            "clocks": clk,
            "leapfrogs": leaps, "grad_evals": grads, "wall_s": t_wall,
        }
        if ess is not None:
            line["ess"] = ess
            line["genes_per_hour_to_ess400"] = ess["genes_per_hour_to_ess400"]
        if secondary is not None:
            line["secondary"] = secondary
        if world == 1 and not args.no_cpu:
            nthr = os.cpu_count() or 1
            cg, cdt, sample, cwall = cpu_reference_run(family, shared, genes, float(args.cpu_iters or 20), nthr)
            line["cpu_baseline"] = {"value": cg / cdt, "unit": UNIT, "cores": nthr, "kind": "port", "sample": sample,
                                    "seconds": cwall, "wall_clock_value": cg / cwall,
                                    "note": "CPU restatement (oracle/), not CmdStan: CmdStan cannot be built in this image"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
