"""Scratch probe (GPU box): throughput of the sampler on named shapes; prints one line per shape."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, ".")
from csplotch_b200 import api, synth

def probe(name, family, shared, G, nch, iters, chunk):
    genes = synth.simulate_counts(shared, G, 99)
    model = api.Model(shared, family)
    batch = api.Batch(model, genes["counts"], genes.get("beta_prior_mean"), genes.get("beta_prior_std"))
    cfg = api.nuts_cfg(iters, 0, nch, seed=1)
    s = api.Sampler(batch, cfg)
    done = 0; out = []
    while done < iters:
        c0 = s.counters()
        t = time.time(); s.advance(chunk); ms = s.last_ms()
        c1 = s.counters()
        done += chunk
        nl = c1["n_grad"] - c0["n_grad"]
        out.append((done, ms, nl, nl / (ms * 1e-3)))
    try:
        print(json.dumps(dict(shape=name, layout=s.layout(), tail=s.launch_tail())), flush=True)
    except AttributeError:
        pass   # an older library variant (A-B runs)
    pr = s.profile(); tot = sum(pr[k] for k in ("cyc_eval", "cyc_merge", "cyc_copy", "cyc_other"))
    cnt = s.counters()
    print(json.dumps(dict(shape=name, profile={k: round(pr[k] / tot, 3) for k in ("cyc_eval", "cyc_merge", "cyc_copy", "cyc_other")},
                          cyc_per_eval=round(pr["cyc_eval"] / cnt["n_grad"]), pro=round(pr["cyc_pro"] / cnt["n_grad"]), fin=round(pr["cyc_fin"] / cnt["n_grad"]), cyc_per_merge=round(pr["cyc_merge"] / max(1, pr["n_merge"])),
                          merges_per_leap=round(pr["n_merge"] / cnt["n_leapfrog"], 2), copies_per_leap=round(pr["n_copy"] / cnt["n_leapfrog"], 2))), flush=True)
    N, D = model.N, model.dim
    bl = 56 * D + 4 * N
    for (d, ms, nl, rate) in out:
        print(json.dumps(dict(shape=name, G=G, chains=nch, iters_done=d, ms=round(ms, 2), grads=nl,
                              grad_per_s=round(rate, 1), algGBs=round(rate * bl / 1e9, 1))), flush=True)

if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c3"]
    if "c2" in which:
        probe("c2", "splotch_stan_model", synth.shared_c2(), int(os.environ.get("PROBE_G", "148")), 4, 60, 20)
    if "c3" in which:
        probe("c3", "comp_splotch_stan_model", synth.shared_c3(), 74, 4, 12, 4)
    if "c4" in which:
        probe("c4", "splotch_stan_model_csr", synth.shared_c4(), 74, 4, 6, 2)
