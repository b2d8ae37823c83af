import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from csplotch_b200 import api, synth
import test_gpu_parity as T
family = "comp_splotch_stan_model"
shared = T._shuffled_grid_design(4, family, side=24)
genes = synth.simulate_counts(shared, 2, 6)
b = T._batch(shared, genes, family, gene_id0=3)
s = api.Sampler(b, api.nuts_cfg(20, 5, 2, seed=9))
s.advance(2)
print("ok", s.counters())
