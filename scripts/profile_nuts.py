"""Driver for ncu on the sampler kernel itself: a short resident run (prints the gradient counters of the
profiled launch so that DRAM bytes can be divided by gradient evaluations)."""
import json
import sys
import numpy as np
sys.path.insert(0, ".")
from csplotch_b200 import api, synth

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
G = int(sys.argv[2]) if len(sys.argv) > 2 else 148
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 1
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 0      # untimed transitions first, so that the profiled launch sees trees of normal depth
cfg = synth.CONFIGS[which]
shared = cfg["shared"]()
genes = synth.simulate_counts(shared, G, 5)
model = api.Model(shared, cfg["family"])
batch = api.Batch(model, genes["counts"], genes.get("beta_prior_mean"), genes.get("beta_prior_std"))
s = api.Sampler(batch, api.nuts_cfg(50, 0, 4, seed=1))
out = []
if warm:
    s.advance(warm)                        # launch 0
    out.append(dict(step=-1, ms=s.last_ms(), grads=s.counters()["n_grad"], leapfrogs=s.counters()["n_leapfrog"]))
for step in range(2 if warm else 3):       # launch 0 includes initialisation; ncu profiles launch index 2
    c0 = s.counters()
    s.advance(iters)
    c1 = s.counters()
    out.append(dict(step=step, ms=s.last_ms(), grads=c1["n_grad"] - c0["n_grad"], leapfrogs=c1["n_leapfrog"] - c0["n_leapfrog"]))
print(json.dumps(dict(shape=which, G=G, N=model.N, D=model.dim, launches=out)))
