#!/usr/bin/env python
"""OFF-BOX recipe: pin the oracle (and through it the CUDA path) against Stan itself.

Neither stanc, CmdStan, BridgeStan nor a network exists in the build image, so `log_prob` / `grad_log_prob` of the
reference's models cannot be evaluated there and parity stays "unpinned" at the Stan boundary.  Run THIS script on any
machine with BridgeStan (`pip install bridgestan`, it fetches and builds Stan on first use) and a checkout of the
reference, then commit the file it writes; tests/test_stan_vectors.py (CPU: oracle, GPU: CUDA path) consumes it when
present and skips otherwise.

    python tests/golden/make_stan_vectors.py --reference /path/to/cSplotch [--out tests/golden/stan_vectors_c1.npz]

What it does, per gene g = 1..7 of tests/golden/c1/data_<g>.R (the reference's own example through its own input
pipeline) with stan/comp_splotch_stan_model.stan:
  * converts the R dump to Stan JSON (same values, same shapes);
  * compiles the model with BridgeStan and evaluates log_density_gradient(theta, propto=True, jacobian=J) for J in
    {True, False} at seeded unconstrained points theta ~ N(0, 1) and N(0, 2^2) (the same generator the tests use);
  * stores theta, lp, grad and BridgeStan's param_unc_names(): the names fix the stanc-version-dependent flattening of
    `row_vector[K] beta_raw[L, C]`, so the consumer permutes by NAME, not by assumed order.
An equivalent CmdStan route (>= 2.31): `./comp_splotch_stan_model log_prob jacobian=1 unconstrained_params=theta.csv
data file=data_<g>.json` — same numbers, CSV instead of an API.
Also writes the same for stan/splotch_stan_model.stan and stan/splotch_stan_model_csr.stan on a c2-geometry gene when
--all-families is given (synthetic counts from csplotch_b200.synth, seed fixed).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def to_stan_json(d: dict) -> dict:
    out = {}
    for k, v in d.items():
        a = np.asarray(v)
        if a.ndim == 0:
            out[k] = int(a) if np.issubdtype(a.dtype, np.integer) else float(a)
        elif np.issubdtype(a.dtype, np.integer):
            out[k] = a.astype(int).tolist()
        else:
            out[k] = a.astype(float).tolist()
    # W_n is declared `int W_n` in the model but stored as a length-1 array in the R dump
    if "W_n" in out and isinstance(out["W_n"], list):
        out["W_n"] = int(out["W_n"][0]) if out["W_n"] else 0
    return out


def points(dim: int, seed: int, n: int = 3):
    rng = np.random.default_rng(seed)
    th = rng.normal(0.0, 1.0, (n, dim))
    th[n // 2:] *= 2.0
    return th


def evaluate(stan_file: str, data: dict, seed: int, tmpdir: str):
    import bridgestan as bs
    path = os.path.join(tmpdir, "data_%d.json" % seed)
    with open(path, "w") as f:
        json.dump(to_stan_json(data), f)
    m = bs.StanModel(stan_file, data=path)
    dim = m.param_unc_num()
    th = points(dim, seed)
    lp = np.zeros((2, len(th)))
    grad = np.zeros((2, len(th), dim))
    for ji, jac in enumerate((True, False)):
        for i, t in enumerate(th):
            lp[ji, i], grad[ji, i] = m.log_density_gradient(t, propto=True, jacobian=jac)
    return dict(theta=th, lp_jac=lp[0], grad_jac=grad[0], lp_nojac=lp[1], grad_nojac=grad[1],
                unc_names=np.array(m.param_unc_names()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="checkout of adaly/cSplotch (for stan/*.stan)")
    ap.add_argument("--out", default=os.path.join(HERE, "stan_vectors_c1.npz"))
    ap.add_argument("--all-families", action="store_true")
    args = ap.parse_args()
    import tempfile
    from csplotch_b200 import rdump, synth
    res = {}
    with tempfile.TemporaryDirectory() as tmp:
        stan = os.path.join(args.reference, "stan", "comp_splotch_stan_model.stan")
        for g in range(1, 8):
            d = rdump.read_rdump(os.path.join(HERE, "c1", "data_%d.R" % g))
            for k, v in evaluate(stan, d, 100 + g, tmp).items():
                res["c1_g%d_%s" % (g, k)] = v
        if args.all_families:
            shared = synth.shared_c2()
            genes = synth.simulate_counts(shared, 1, 4242)
            d = synth.gene_dict(shared, genes, 0)
            for fam in ("splotch_stan_model", "splotch_stan_model_csr"):
                for k, v in evaluate(os.path.join(args.reference, "stan", fam + ".stan"), d, 4242, tmp).items():
                    res["c2_%s_%s" % (fam, k)] = v
    import bridgestan
    res["bridgestan_version"] = np.array(getattr(bridgestan, "__version__", "unknown"))
    np.savez_compressed(args.out, **res)
    print("wrote", args.out, "with", len(res), "arrays")


if __name__ == "__main__":
    main()
