"""Small driver for ncu: stand-alone log_prob_grad / fixed-step leapfrog launches (the same fused code
as the sampler's leapfrog)."""
import sys
import time
import numpy as np
sys.path.insert(0, ".")
from csplotch_b200 import api, synth

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 592
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mode = sys.argv[4] if len(sys.argv) > 4 else "grad"
cfg = synth.CONFIGS[which]
shared = cfg["shared"]()
genes = synth.simulate_counts(shared, 64, 5)
model = api.Model(shared, cfg["family"])
batch = api.Batch(model, genes["counts"], genes.get("beta_prior_mean"), genes.get("beta_prior_std"))
rng = np.random.default_rng(0)
theta = rng.normal(0, 0.5, (B, model.dim))
go = np.arange(B) % 64
for r in range(reps):
    t = time.time()
    if mode == "grad":
        lp, g = batch.log_prob_grad(theta, gene_of=go)
        print(which, "grad B", B, "wall", round(time.time() - t, 3), "lp0", lp[0])
    else:
        p0 = rng.normal(0, 1, (B, model.dim))
        th, p, H = batch.leapfrog_fixed(theta, p0, np.ones((B, model.dim)), 1e-4, 4, gene_of=go)
        print(which, "leap B", B, "wall", round(time.time() - t, 3), "H0", H[0])
