import sys
import numpy as np
sys.path.insert(0, ".")
import oracle
from csplotch_b200 import api, synth

family, n_levels, zi, car = sys.argv[1] if len(sys.argv) > 1 else "comp_splotch_stan_model", 3, 1, 1
shared = synth.shared_tiny(seed=5, family=family, n_levels=n_levels, zi=zi, car=car)
genes = synth.simulate_counts(shared, 2, 6)
model = api.Model(shared, family)
b = api.Batch(model, genes["counts"], genes.get("beta_prior_mean"), genes.get("beta_prior_std"), gene_id0=9)
nw, ns = 40, 15
cfg = api.nuts_cfg(nw, ns, 1, seed=4, keep_draws=True)
s = api.Sampler(b, cfg)
ocfg = oracle.nuts_cfg(nw, ns, seed=4)
o = oracle.Oracle(synth.gene_dict(shared, genes, 0), family)
want = o.nuts(ocfg, gene_id=9, chain_id=1, save_warmup=True)
w = want["draws"]
for it in range(nw + ns):
    s.advance(1)
    th, eps, mi = s.state()
    dq = np.max(np.abs(th[0, 0] - w[it, 7:]))
    cnt = s.counters()
    print(it, "dq=%.3e" % dq, "eps_gpu=%.6g" % eps[0, 0], "oracle: eps=%.6g depth=%d nleap=%d div=%d acc=%.4f" % (
        w[it, 2], w[it, 3], w[it, 4], w[it, 5], w[it, 1]), "gpu nleap_total=%d" % cnt["n_leapfrog"])
    if dq > 1e-3:
        break
