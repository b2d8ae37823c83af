"""GPU box: device-timed microbenchmark of the fused leapfrog pass (csplotch_leapfrog_bench) on named shapes.

    python scripts/leap_bench.py [c2 c3 c4] [--waves W] [--steps S]

Prints one JSON line per shape: ns per evaluation (launch time / evaluations), the algorithmic GB/s
(B_leap = 56 D + 4 N per leapfrog) and its fraction of the measured HBM peak.  The library is
whatever CSPLOTCH_LIB names (variants of scripts/build_variants.sh), else the product library."""
import json
import os
import sys

sys.path.insert(0, ".")
import numpy as np

from csplotch_b200 import api, synth


def main():
    args = [a for a in sys.argv[1:] if a in synth.CONFIGS]
    shapes = args or ["c2", "c3"]
    waves = int(sys.argv[sys.argv.index("--waves") + 1]) if "--waves" in sys.argv else 2
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 8
    peak = 6533.2
    pk = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    for name in shapes:
        cfg = synth.CONFIGS[name]
        shared = cfg["shared"]()
        genes = synth.simulate_counts(shared, 16, 5)
        model = api.Model(shared, cfg["family"])
        batch = api.Batch(model, genes["counts"], genes.get("beta_prior_mean"), genes.get("beta_prior_std"))
        per_sm = 2                      # resident CTAs per SM of every kernel (round 1: 3 for K <= 2)
        per_sm = int(os.environ.get("CSP_CTAS_PER_SM", per_sm))
        for w in sorted({1, waves}):
            B = 148 * per_sm * w
            ms = batch.leapfrog_bench(B, 1e-4, steps, 3)
            evals = B * (steps + 1)
            bl = 56 * model.dim + 4 * model.N
            gbs = B * steps * bl / (ms * 1e-3) / 1e9
            print(json.dumps(dict(shape=name, lib=os.path.basename(os.environ.get("CSPLOTCH_LIB", "product")), B=B,
                                  steps=steps, ms=round(ms, 3), ns_per_eval=round(ms * 1e6 / evals, 1),
                                  evals_per_s=round(evals / (ms * 1e-3)), alg_GBs=round(gbs, 1),
                                  frac_hbm=round(gbs / peak, 4))), flush=True)
        batch.close(); model.close()


if __name__ == "__main__":
    main()
