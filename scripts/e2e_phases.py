"""Wall-clock breakdown of one end-to-end step (the bench's e2e arm) into its host-API phases (GPU box)."""
import json
import sys
import time
import numpy as np
sys.path.insert(0, ".")
from csplotch_b200 import api, synth

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
G = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
cfg = synth.CONFIGS[which]
shared = cfg["shared"]()
genes = synth.simulate_counts(shared, G, 5)
model = api.Model(shared, cfg["family"])
counts = np.ascontiguousarray(genes["counts"])
pm, ps = genes.get("beta_prior_mean"), genes.get("beta_prior_std")
for rep in range(2):
    t = [time.time()]
    b = api.Batch(model, counts, pm, ps); t.append(time.time())
    s = api.Sampler(b, api.nuts_cfg(iters // 2, iters - iters // 2, 4, seed=7)); t.append(time.time())
    s.run(); t.append(time.time())
    ms = s.last_ms()
    d = s.draws(); t.append(time.time())
    sm = s.spot_summaries(); t.append(time.time())
    cnt = s.counters()
    s.close(); b.close(); t.append(time.time())
    ph = np.diff(t)
    print(json.dumps(dict(shape=which, G=G, iters=iters, rep=rep, batch_s=ph[0], sampler_create_s=ph[1], run_s=ph[2],
                          kernel_ms=ms, draws_s=ph[3], summaries_s=ph[4], close_s=ph[5], total_s=t[-1] - t[0],
                          grads=cnt["n_grad"], grads_per_s_e2e=cnt["n_grad"] / (t[-1] - t[0]),
                          grads_per_s_kernel=cnt["n_grad"] / (ms * 1e-3))), flush=True)
