"""`splotch_vinit` against plain `splotch` on c2 genes: gradient evaluations, device time, bulk-ESS and
R-hat of the betas for (a) N warm-up transitions from uniform(-2, 2) starts and (b) multi-path Pathfinder starts +
150 warm-up transitions (bin/splotch_vinit:137-194).  usage: vinit_compare.py [G] [N] [max_lbfgs_iters]"""
import json
import sys
import time
import numpy as np
sys.path.insert(0, ".")
from csplotch_b200 import api, synth, pathfinder, diagnostics as dg

G = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cfg = synth.CONFIGS["c2"]
family, shared = cfg["family"], cfg["shared"]()
genes = synth.simulate_counts(shared, G, cfg["seed"])
model = api.Model(shared, family)
batch = api.Batch(model, genes["counts"])


def run(num_warmup, theta_init):
    s = api.Sampler(batch, api.nuts_cfg(num_warmup, n, 4, seed=11), theta_init=theta_init)
    ms = 0.0
    while s.status()[1].min() < num_warmup + n:
        s.advance(50)
        ms += s.last_ms()
    ex = s.extract_small()
    bkeys = [k for k in ex if k.startswith("beta_level")]
    ess, rhat = [], []
    for g in range(G):
        cb = np.concatenate([ex[k][g].reshape(4 * n, -1) for k in bkeys], axis=1).reshape(4, n, -1)
        e, r = dg.summarize(cb)
        ess.append(e); rhat.append(r)
    d = ex["diag"]
    return dict(num_warmup=num_warmup, device_seconds=ms * 1e-3, grad_evals=s.counters()["n_grad"],
                median_min_bulk_ess_beta=float(np.median(ess)), median_max_rhat_beta=float(np.median(rhat)),
                mean_stepsize=float(d[..., 2].mean()), mean_n_leapfrog=float(d[..., 4].mean()),
                divergent_frac=float(d[..., 5].mean()), failed_chains=int((s.status()[0] != 0).sum()))


res = dict(shape="c2", genes=G, chains=4, num_samples=n)
res["plain"] = run(n, None)
t0 = time.time()
pf = pathfinder.pathfinder(pathfinder.BatchTarget(batch), num_paths=4, num_draws=200, num_out=4, seed=11,
                           max_lbfgs_iters=iters)
res["pathfinder"] = dict(wall_seconds=time.time() - t0, max_lbfgs_iters=iters, lp_grad_evals=pf.n_lp_grad, lp_evals=pf.n_lp,
                         median_best_iter=float(np.median(pf.best_iter)), median_pareto_k=float(np.median(pf.pareto_k)),
                         median_lp_of_starts=float(np.median(pf.lp)), failed_genes=int((~np.isfinite(pf.lp).all(axis=1)).sum()))
res["vinit"] = run(150, pf.draws)
print(json.dumps(res))
