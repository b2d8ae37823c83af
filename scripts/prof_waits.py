"""Development instrumentation driver (build with CSPLOTCH_DEV_DEFS=-DCSPLOTCH_PROF_WAITS, --dev): cycles thread 0 of
every CTA spends in each wait of the tile loop, per gradient evaluation."""
import json, os, sys
import numpy as np
sys.path.insert(0, ".")
from csplotch_b200 import api, synth

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
G = int(sys.argv[2]) if len(sys.argv) > 2 else 111
cfg = synth.CONFIGS[which]
shared = cfg["shared"]()
genes = synth.simulate_counts(shared, G, 5)
model = api.Model(shared, cfg["family"])
batch = api.Batch(model, genes["counts"], genes.get("beta_prior_mean"), genes.get("beta_prior_std"))
s = api.Sampler(batch, api.nuts_cfg(100, 0, 4, seed=1))
s.advance(6)
c0 = s.counters(); p0 = s.profile()
s.advance(4)
c1 = s.counters(); p1 = s.profile()
ng = c1["n_grad"] - c0["n_grad"]
NT = (model.N + 255) // 256
out = dict(shape=which, G=G, grads=ng, NT=NT, ms=s.last_ms(),
           cyc_eval=p1["cyc_eval"] / ng, w_issue_warp0=p1["cyc_merge"] / ng, w_rest=p1["cyc_copy"] / ng,
           w_ring=p1["cyc_other"] / ng, w_release_warp0=p1["n_merge"] / ng, w_bins=p1["n_copy"] / ng,
           w_landed_to_ringwait=p1["cyc_pro"] / ng, w_ring_to_release=p1["cyc_fin"] / ng)
out = {k: (round(v) if isinstance(v, float) and v > 100 else v) for k, v in out.items()}
print(json.dumps(out))
