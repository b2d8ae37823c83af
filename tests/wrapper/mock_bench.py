"""TEST INFRASTRUCTURE.  Dry run of bench.py's GPU arm on a box without a GPU: torch.cuda and the device classes are mocked (oracle-backed
stand-ins), to catch Python-level mistakes in the arm's control flow / JSON assembly."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from csplotch_b200 import api, rdump
from tests import test_vinit_host as fakes

torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda x: None
torch.cuda.synchronize = lambda *a: None
api.check = lambda rc: None
api.pinned_empty = lambda shape, dtype=np.float64: np.empty(shape, dtype=dtype)

class Batch(fakes._Batch):
    def close(self): pass

class Sampler(fakes._Sampler):
    def __init__(self, *a, **k):
        super().__init__(*a, **k); self._adv = 0; self._calls = 0; self._ran = False
    def advance(self, n, sync=True):
        self._adv += n; self._calls += 1
    def last_ms(self): return 5.0
    def launch_tail(self): return 0.9, 1.0
    def profile(self): return dict(cyc_eval=80, cyc_merge=15, cyc_copy=1, cyc_other=4, n_merge=1, n_copy=1, cyc_pro=1, cyc_fin=1)
    def run(self, chunk=None):
        super().run(chunk); self._ran = True; self.total_ms = 12.0; self._calls += 1; return self
    def counters(self):
        if self._ran: return super().counters()
        return dict(n_grad=100 * self._adv, n_leapfrog=90 * self._adv, n_divergent=0, n_launches=self._calls)
    def draws(self, out=None):
        d, s, n = super().draws()
        if out is not None:
            out[0][...] = d; out[1][...] = s; out[2][...] = n
            return out
        return d, s, n
    def spot_summaries(self, out=None):
        N = self.model.N
        return {k: np.ones((self.G, N)) for k in ("log_lambda/mean", "log_lambda/std", "lambda/mean", "lambda/std", "psi/mean", "psi/std")}
    def close(self): pass

class Model(fakes._Model):
    def close(self): pass

api.Model, api.Batch, api.Sampler = Model, Batch, Sampler
gdir = os.path.join(ROOT, "tests", "golden", "c1")
datas = [rdump.read_rdump(os.path.join(gdir, "data_%d.R" % g)) for g in (1, 2, 3)]
shared = {k: v for k, v in datas[0].items() if k not in rdump.GENE_KEYS}
genes = {k: np.stack([d[k] for d in datas]) for k in rdump.GENE_KEYS}
bench.make_inputs = lambda workload, G, rank, head_only=False: ("comp_splotch_stan_model", shared, genes)
sys.argv = ["bench.py", "--steps", "2", "--warmup", "1", "--no-cpu", "--genes", "3", "--e2e-genes", "2", "--e2e-iters", "8",
            "--ess-genes", "2", "--ess-n", "10", "--no-secondary"]
rc = bench.main()
print("rc", rc, file=sys.stderr)
