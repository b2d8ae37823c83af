"""Second half of the BASELINE metric: genes / hour to bulk-ESS >= 400 (min over beta_level_* entries and the
scalar parameters, 4 chains pooled, rank-normalised; csplotch_b200/diagnostics.py) for a full
`splotch -n N -c 4` run (N warm-up + N sampling transitions) of a gene batch."""
import json
import sys
import time
import numpy as np
sys.path.insert(0, ".")
from csplotch_b200 import api, synth, diagnostics as dg

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
G = int(sys.argv[2]) if len(sys.argv) > 2 else 592
n = int(sys.argv[3]) if len(sys.argv) > 3 else 200
n_eval = int(sys.argv[4]) if len(sys.argv) > 4 else 64
if which == "c1":
    import os
    from csplotch_b200 import rdump
    datas = [rdump.read_rdump(os.path.join("tests", "golden", "c1", "data_%d.R" % g)) for g in range(1, 8)]
    family = "comp_splotch_stan_model"
    shared = {k: v for k, v in datas[0].items() if k not in rdump.GENE_KEYS}
    genes = {"counts": np.stack([d["counts"] for d in datas]), "beta_prior_mean": np.stack([d["beta_prior_mean"] for d in datas]),
             "beta_prior_std": np.stack([d["beta_prior_std"] for d in datas])}
    G = 7
else:
    cfg = synth.CONFIGS[which]
    family, shared = cfg["family"], cfg["shared"]()
    genes = synth.simulate_counts(shared, G, cfg["seed"])
model = api.Model(shared, family)
batch = api.Batch(model, genes["counts"], genes.get("beta_prior_mean"), genes.get("beta_prior_std"))
t0 = time.time()
s = api.Sampler(batch, api.nuts_cfg(n, n, 4, seed=11))
ms = 0.0
while True:
    s.advance(50)
    ms += s.last_ms()
    if s.status()[1].min() >= 2 * n:
        break
wall = time.time() - t0
cnt = s.counters()
ex = s.extract_small()
bkeys = [k for k in ex if k.startswith("beta_level")]
skeys = [k for k in ("tau", "a", "sigma", "sigma_level_2", "sigma_level_3", "theta") if k in ex]
ess_min, rhat_max, ess_beta, rhat_beta = [], [], [], []
for g in range(min(G, n_eval)):
    cb = np.concatenate([ex[k][g].reshape(4 * n, -1) for k in bkeys], axis=1).reshape(4, n, -1)
    cs = np.concatenate([ex[k][g].reshape(4 * n, -1) for k in skeys], axis=1).reshape(4, n, -1)
    eb, rb = dg.summarize(cb)
    es, rs = dg.summarize(cs)
    ess_beta.append(eb); rhat_beta.append(rb)
    ess_min.append(min(eb, es)); rhat_max.append(max(rb, rs))
ess_min, rhat_max, ess_beta, rhat_beta = map(np.array, (ess_min, rhat_max, ess_beta, rhat_beta))
diag = ex["diag"]
res = dict(shape=which, family=family, genes=G, chains=4, num_warmup=n, num_samples=n, device_seconds=ms * 1e-3, wall_seconds=wall,
           grad_evals=cnt["n_grad"], grad_evals_per_s=cnt["n_grad"] / (ms * 1e-3), genes_per_hour=G / (ms * 1e-3) * 3600,
           genes_evaluated_for_ess=int(len(ess_min)), frac_genes_ess_ge_400=float((ess_min >= 400).mean()),
           median_min_bulk_ess=float(np.median(ess_min)), genes_per_hour_to_ess400=float(G / (ms * 1e-3) * 3600 * (ess_min >= 400).mean()),
           median_max_rhat=float(np.median(rhat_max)), frac_rhat_lt_1_01=float((rhat_max < 1.01).mean()),
           beta_only=dict(median_min_bulk_ess=float(np.median(ess_beta)), frac_genes_ess_ge_400=float((ess_beta >= 400).mean()),
                          genes_per_hour_to_ess400=float(G / (ms * 1e-3) * 3600 * (ess_beta >= 400).mean()),
                          median_max_rhat=float(np.median(rhat_beta)), frac_rhat_lt_1_01=float((rhat_beta < 1.01).mean())),
           mean_treedepth=float(diag[..., 3].mean()), mean_n_leapfrog=float(diag[..., 4].mean()),
           divergent_frac=float(diag[..., 5].mean()), mean_accept_stat=float(diag[..., 1].mean()),
           failed_chains=int((s.status()[0] != 0).sum()))
print(json.dumps(res))
