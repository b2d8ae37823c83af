#!/usr/bin/env python
"""Known-answer vectors for the density on the reference's own c1 files (tests/golden/c1/data_<g>.R): seeded
unconstrained points theta, lp (propto, Jacobian on and off) and the gradient, from the C oracle, accepted only where
the independent torch-autograd transcription of the .stan file (tests/stan_transcript.py) agrees to 1e-11.
No Stan toolchain exists in this image, so these pin the restatement, not CmdStan itself.
Writes tests/golden/c1_lpgrad.npz.  Run from the repo root: python tests/golden/make_lpgrad_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from csplotch_b200 import rdump  # noqa: E402
from tests.stan_transcript import stan_log_prob_grad  # noqa: E402

FAMILY = "comp_splotch_stan_model"
rng = np.random.default_rng(20240)
theta, lp_j, lp_n, grad_j = [], [], [], []
for g in range(1, 8):
    d = rdump.read_rdump(os.path.join(HERE, "c1", "data_%d.R" % g))
    o = oracle.Oracle(d, FAMILY)
    for scale in (0.5, 1.0, 2.0):
        th = rng.normal(0, scale, o.dim)
        lj, gj = o.log_prob_grad(th, jacobian=True)
        ln = o.log_prob_grad(th, jacobian=False, want_grad=False)
        lj2, gj2 = stan_log_prob_grad(d, FAMILY, th, jacobian=True)
        ln2, _ = stan_log_prob_grad(d, FAMILY, th, jacobian=False)
        assert abs(lj - lj2) <= 1e-11 * max(1, abs(lj2)) and abs(ln - ln2) <= 1e-11 * max(1, abs(ln2))
        assert np.max(np.abs(gj - gj2) / np.maximum(1, np.abs(gj2))) < 1e-11
        theta.append(th); lp_j.append(lj); lp_n.append(ln); grad_j.append(gj)
np.savez_compressed(os.path.join(HERE, "c1_lpgrad.npz"), theta=np.array(theta).reshape(7, 3, -1),
                    lp_jacobian=np.array(lp_j).reshape(7, 3), lp_no_jacobian=np.array(lp_n).reshape(7, 3),
                    grad_jacobian=np.array(grad_j).reshape(7, 3, -1))
print("wrote c1_lpgrad.npz", np.array(lp_j).reshape(7, 3)[0])
