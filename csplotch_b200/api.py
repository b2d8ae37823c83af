"""Host-side mirror of the reference's interface for the sampling path.

Reference call sites this stands in for (relative to /root/reference):
  * bin/splotch:129-133      `BINARY sample num_samples=N num_warmup=N random id=CHAIN data file=… output file=…`
  * simulation/fitting.py:35-51   `pystan.StanModel(file).sampling(data=dict, iter=…, chains=…)`,
                                  `fit.extract(permuted=True)['beta_level_1']`
Everything numerical happens in libcsplotch_b200.so (CUDA, sm_100a) through the C ABI of
include/csplotch.h; this module only marshals numpy arrays.  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import CsplotchError, Dims, NutsCfg, SharedDesc, check

FAMILIES = {
    "splotch_stan_model": 0,
    "comp_splotch_stan_model": 1,
    "splotch_stan_model_csr": 2,
}


def family_from_binary(path_or_name) -> int:
    """`-b BINARY` of bin/splotch: the basename selects the model family."""
    if isinstance(path_or_name, int):
        return path_or_name
    name = os.path.basename(str(path_or_name))
    for suffix in (".stan", ".exe"):
        if name.endswith(suffix):
            name = name[: -len(suffix)]
    if name not in FAMILIES:
        raise ValueError("unknown Splotch binary %r; expected one of %s" % (name, sorted(FAMILIES)))
    return FAMILIES[name]


def _i32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32).reshape(-1))


def _f64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class _PinnedOwner:
    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __del__(self):
        try:
            if self.ptr:
                self.lib.csplotch_host_free(C.c_void_p(self.ptr))
                self.ptr = None
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64):
    """numpy array in page-locked host memory (csplotch_host_alloc); freed with the array.  Use it for counts
    handed to Batch and as `out=` of Sampler.draws / spot_summaries: transfers then run at PCIe speed."""
    lib = _lib.lib()
    _lib.require_gpu()
    shape = (shape,) if np.isscalar(shape) else tuple(int(x) for x in shape)
    dt = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
    if nbytes == 0:
        return np.empty(shape, dtype=dt)
    p = C.c_void_p()
    check(lib.csplotch_host_alloc(C.c_uint64(nbytes), C.byref(p)))
    buf = (C.c_ubyte * nbytes).from_address(p.value)
    buf._owner = _PinnedOwner(lib, p.value)       # arr.base -> buf -> owner: freed when the last view dies
    return np.frombuffer(buf, dtype=dt).reshape(shape)


class Model:
    """Gene-shared covariates (the data{} block minus `counts`, `beta_prior_*`) resident on the GPU."""

    def __init__(self, data: dict, family):
        self.lib = _lib.lib()
        _lib.require_gpu()
        fam = family_from_binary(family)
        self.family = fam
        d = data
        k = {}
        k["N_spots"] = _i32(np.atleast_1d(d["N_spots"]))
        N = int(k["N_spots"].sum())
        K = int(d.get("N_celltypes", 1)) if fam == 1 else 1
        car = int(d.get("car", 0))
        k["l2"] = _i32(d.get("level_2_mapping", []))
        k["l3"] = _i32(d.get("level_3_mapping", []))
        k["tm"] = _i32(np.atleast_1d(d["tissue_mapping"]))
        k["sf"] = _f64(d["size_factors"])
        k["D"] = _i32(d["D"])
        if k["sf"].size != N or k["D"].size != N:
            raise CsplotchError("size_factors / D must have sum(N_spots) entries")
        k["E"] = None
        if fam == 1:
            E = np.asarray(d["E"], dtype=np.float64)
            if E.size != N * K:
                raise CsplotchError("E must be (sum(N_spots), N_celltypes)")
            k["E"] = np.ascontiguousarray(E.reshape(N, K)).reshape(-1)
        W_n = int(np.atleast_1d(d.get("W_n", [0]))[0]) if car and np.size(d.get("W_n", [])) else 0
        k["W"] = k["U"] = k["V"] = k["Ds"] = k["eig"] = None
        if car:
            if "W_sparse" in d and np.size(d["W_sparse"]):
                k["W"] = _i32(np.asarray(d["W_sparse"]).reshape(W_n, 2))
            if "U" in d and np.size(d["U"]):
                k["U"] = _i32(d["U"])
                k["V"] = _i32(d["V"])
                if k["U"].size != N + 1:
                    raise CsplotchError("U must have N+1 entries")
                if not W_n:
                    W_n = k["V"].size // 2
            if not np.size(d.get("eig_values", [])) or not np.size(d.get("D_sparse", [])):
                # a dictionary without the CAR pre-compute (generate_dictionary's eigvalsh block, utils.py:749-759):
                # done here on the GPU; tissues beyond 9 999 spots need `coords` for the chunked solution
                from . import car as _car
                pre = _car.precompute(k["N_spots"], np.asarray(d["W_sparse"]).reshape(W_n, 2), coords=d.get("coords"))
                d = dict(d, eig_values=pre["eig_values"], D_sparse=pre["D_sparse"])
            k["Ds"] = _f64(d["D_sparse"])
            k["eig"] = _f64(d["eig_values"])
            if k["Ds"].size != N or k["eig"].size != N:
                raise CsplotchError("D_sparse / eig_values must have sum(N_spots) entries")
        self._keep = k
        desc = SharedDesc(
            fam, int(d["N_tissues"]), _p(k["N_spots"]), int(d["N_covariates"]), K,
            int(d["N_level_1"]), int(d.get("N_level_2", 0)), int(d.get("N_level_3", 0)),
            int(d.get("zi", 0)), car, int(d.get("nb", 0)) if fam == 1 else 0,
            _p(k["l2"]), _p(k["l3"]), _p(k["tm"]), _p(k["sf"]), _p(k["D"]), _p(k["E"]),
            W_n, _p(k["W"]), _p(k["U"]), _p(k["V"]), _p(k["Ds"]), _p(k["eig"]))
        h = C.c_void_p()
        check(self.lib.csplotch_model_create(C.byref(desc), C.byref(h)))
        self.h = h
        dims = Dims()
        check(self.lib.csplotch_model_dims(self.h, C.byref(dims)))
        self.N, self.dim, self.n_beta, self.n_scalar = dims.N, dims.dim, dims.n_beta, dims.n_scalar
        self.n_small, self.n_outputs = dims.n_small, dims.n_outputs
        self.K = K
        self.C = int(d["N_covariates"])
        self.levels = (int(d["N_level_1"]), int(d.get("N_level_2", 0)), int(d.get("N_level_3", 0)))
        self.car, self.zi = car, int(d.get("zi", 0))
        self.nb = int(d.get("nb", 0)) if fam == 1 else 0
        self.level_2_mapping = np.asarray(d.get("level_2_mapping", []), dtype=np.int64).reshape(-1)
        self.level_3_mapping = np.asarray(d.get("level_3_mapping", []), dtype=np.int64).reshape(-1)

    def close(self):
        if getattr(self, "h", None):
            self.lib.csplotch_model_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- names / shapes of the Stan CSV columns ----------------------------
    def column_names(self):
        L1, L2, L3 = self.levels
        L = L1 + L2 + L3
        N, Cc, K = self.N, self.C, self.K
        comp = self.family == 1
        cols = []
        if self.car:
            cols += ["psi.%d" % (i + 1) for i in range(N)]
        if comp:
            cols += ["beta_raw.%d.%d.%d" % (i + 1, j + 1, k + 1) for k in range(K) for j in range(Cc) for i in range(L)]
        else:
            cols += ["beta_raw.%d.%d" % (i + 1, j + 1) for j in range(Cc) for i in range(L)]
        if self.car:
            cols += ["tau.1", "a.1"]
        cols += ["sigma"]
        if L2:
            cols += ["sigma_level_2.1"]
        if L3:
            cols += ["sigma_level_3.1"]
        if self.zi:
            cols += ["theta.1"]
        if self.nb:
            cols += ["phi.1"]
        cols += ["noise_raw.%d" % (i + 1) for i in range(N)]
        cols += ["log_lambda.%d" % (i + 1) for i in range(N)]
        if self.nb:
            cols += ["phi_arr.%d" % (i + 1) for i in range(N)]
        for lev, nr in ((1, L1), (2, L2), (3, L3)):
            if comp:
                cols += ["beta_level_%d.%d.%d.%d" % (lev, i + 1, j + 1, k + 1)
                         for k in range(K) for j in range(Cc) for i in range(nr)]
            else:
                cols += ["beta_level_%d.%d.%d" % (lev, i + 1, j + 1) for j in range(Cc) for i in range(nr)]
        assert len(cols) == self.n_outputs
        return cols


def param_layout(model: "Model"):
    """[(name, theta_index, transform)] of every parameter in the order of the unconstrained vector (include/csplotch.h:
    psi | beta_raw | tau | a | sigma | sigma_level_2 | sigma_level_3 | theta | phi | noise_raw; beta_raw[i, j] at i + L*j,
    for comp_splotch_stan_model beta_raw[i, j, k] at (i*C + j)*K + k).  Names are the dotted CSV column names."""
    L = sum(model.levels)
    N, Cc, K = model.N, model.C, model.K
    out, o = [], 0
    if model.car:
        out += [("psi.%d" % (i + 1), o + i, "id") for i in range(N)]
        o += N
    if model.family == 1:
        out += [("beta_raw.%d.%d.%d" % (i + 1, j + 1, k + 1), o + (i * Cc + j) * K + k, "id")
                for i in range(L) for j in range(Cc) for k in range(K)]
    else:
        out += [("beta_raw.%d.%d" % (i + 1, j + 1), o + i + L * j, "id") for j in range(Cc) for i in range(L)]
    o += L * Cc * K
    scal = ([("tau.1", "log"), ("a.1", "logit")] if model.car else []) + [("sigma", "log")]
    scal += [("sigma_level_2.1", "log")] if model.levels[1] else []
    scal += [("sigma_level_3.1", "log")] if model.levels[2] else []
    scal += [("theta.1", "logit")] if model.zi else []
    scal += [("phi.1", "logphi")] if model.nb else []
    for name, tr in scal:
        out.append((name, o, tr))
        o += 1
    out += [("noise_raw.%d" % (i + 1), o + i, "id") for i in range(N)]
    o += N
    assert o == model.dim
    return out


def unconstrain(model: "Model", init: dict, radius: float = 2.0, seed: int = 0) -> np.ndarray:
    """CmdStan `init=FILE.json` -> unconstrained theta.  `init` maps parameter names to CONSTRAINED values, either by
    dotted CSV column name (the files bin/splotch_vinit:177-182 writes from a Pathfinder draw) or by variable name with
    (nested) arrays as in CmdStan's JSON format.  Non-parameter keys (log_lambda.*, beta_level_*, lp__) are ignored;
    parameters that are absent start uniform(-radius, radius) on the unconstrained scale, as in Stan."""
    flat = {}

    def walk(prefix, v):
        if isinstance(v, (list, tuple, np.ndarray)):
            for i, x in enumerate(v):
                walk("%s.%d" % (prefix, i + 1), x)
        else:
            flat[prefix] = float(v)

    for k, v in init.items():
        walk(str(k).strip(), v)
    rng = np.random.default_rng(seed)
    theta = rng.uniform(-radius, radius, model.dim)
    for name, idx, tr in param_layout(model):
        v = flat.get(name)
        if v is None and name.endswith(".1"):
            v = flat.get(name[:-2])        # scalars declared as real x[1]: "tau" or "tau.1"
        if v is None:
            continue
        if tr == "log":
            if not v > 0:
                raise CsplotchError("init: %s must be positive, got %r" % (name, v))
            v = np.log(v)
        elif tr == "logit":
            if not 0 < v < 1:
                raise CsplotchError("init: %s must lie in (0, 1), got %r" % (name, v))
            v = np.log(v) - np.log1p(-v)
        elif tr == "logphi":
            if not v > 1e-6:
                raise CsplotchError("init: %s must exceed 1e-6, got %r" % (name, v))
            v = np.log(v - 1e-6)
        theta[idx] = v
    return theta


class Batch:
    """counts (+ beta priors) of G genes on the GPU.  counts: (G, N) int; priors: (G, K) for comp."""

    def __init__(self, model: Model, counts, beta_prior_mean=None, beta_prior_std=None, gene_id0: int = 0,
                 gene_ids=None):
        """`gene_ids` (length G) gives every gene of the batch its own id instead of gene_id0 + g: the id keys the
        RNG stream, so a gene list can be handed over in any order (cost-sorted, dealt over ranks: shard.deal)."""
        self.model = model
        self.lib = model.lib
        c = np.ascontiguousarray(np.asarray(counts, dtype=np.int32).reshape(-1, model.N))
        self.G = c.shape[0]
        pm = ps = None
        if model.family == 1:
            if beta_prior_mean is None or beta_prior_std is None:
                raise CsplotchError("comp_splotch_stan_model needs beta_prior_mean / beta_prior_std")
            pm = np.ascontiguousarray(np.asarray(beta_prior_mean, dtype=np.float64).reshape(self.G, model.K))
            ps = np.ascontiguousarray(np.asarray(beta_prior_std, dtype=np.float64).reshape(self.G, model.K))
        self.beta_prior_mean, self.beta_prior_std = pm, ps
        h = C.c_void_p()
        check(self.lib.csplotch_batch_create(model.h, _p(c), _p(pm) if pm is not None else None,
                                             _p(ps) if ps is not None else None, C.c_int32(self.G),
                                             C.c_uint32(gene_id0), C.byref(h)))
        self.h = h
        self.gene_id0 = gene_id0
        self.gene_ids = np.arange(self.G, dtype=np.uint32) + np.uint32(gene_id0)
        if gene_ids is not None:
            ids = np.ascontiguousarray(np.asarray(gene_ids, dtype=np.uint32).reshape(-1))
            if ids.size != self.G:
                raise CsplotchError("gene_ids must have one entry per gene")
            check(self.lib.csplotch_batch_set_gene_ids(self.h, _p(ids)))
            self.gene_ids = ids

    def close(self):
        if getattr(self, "h", None):
            self.lib.csplotch_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # model.log_prob / grad_log_prob of the stanc-generated class (propto=true)
    def log_prob_grad(self, theta, gene_of=None, jacobian=True, want_grad=True):
        m = self.model
        th = np.ascontiguousarray(np.asarray(theta, dtype=np.float64).reshape(-1, m.dim))
        B = th.shape[0]
        go = _i32(gene_of) if gene_of is not None else None
        lp = np.empty(B)
        grad = np.empty((B, m.dim)) if want_grad else None
        check(self.lib.csplotch_log_prob_grad(self.h, _p(th), _p(go) if go is not None else None, C.c_int32(B),
                                              _p(lp), _p(grad) if want_grad else None,
                                              C.c_int32(1 if jacobian else 0)))
        return (lp, grad) if want_grad else lp

    def write_array(self, theta, gene_of=None):
        m = self.model
        th = np.ascontiguousarray(np.asarray(theta, dtype=np.float64).reshape(-1, m.dim))
        B = th.shape[0]
        go = _i32(gene_of) if gene_of is not None else None
        out = np.empty((B, m.n_outputs))
        check(self.lib.csplotch_write_array(self.h, _p(th), _p(go) if go is not None else None, C.c_int32(B), _p(out)))
        return out

    def leapfrog_fixed(self, theta0, p0, inv_metric, eps, n_steps, gene_of=None):
        m = self.model
        th = np.ascontiguousarray(np.asarray(theta0, dtype=np.float64).reshape(-1, m.dim))
        B = th.shape[0]
        p = np.ascontiguousarray(np.asarray(p0, dtype=np.float64).reshape(B, m.dim))
        mi = np.ascontiguousarray(np.asarray(inv_metric, dtype=np.float64).reshape(B, m.dim))
        go = _i32(gene_of) if gene_of is not None else None
        tho, po, H = np.empty((B, m.dim)), np.empty((B, m.dim)), np.empty(B)
        check(self.lib.csplotch_leapfrog_fixed(self.h, _p(th), _p(p), _p(mi), _p(go) if go is not None else None,
                                               C.c_int32(B), C.c_double(eps), C.c_int32(n_steps), _p(tho), _p(po),
                                               _p(H)))
        return tho, po, H

    def leapfrog_bench(self, B, eps=1e-4, n_steps=8, reps=3):
        """Diagnostic: average device milliseconds of one launch of B fixed-step trajectories (csplotch_leapfrog_bench)."""
        ms = C.c_float()
        check(self.lib.csplotch_leapfrog_bench(self.h, C.c_int32(B), C.c_double(eps), C.c_int32(n_steps), C.c_int32(reps),
                                               C.byref(ms)))
        return float(ms.value)


def nuts_cfg(num_warmup, num_samples, num_chains=4, seed=0, keep_draws=False, **kw) -> NutsCfg:
    cfg = NutsCfg()
    _lib.lib().csplotch_nuts_cfg_default(C.byref(cfg))
    cfg.num_warmup, cfg.num_samples, cfg.num_chains = int(num_warmup), int(num_samples), int(num_chains)
    cfg.seed = int(seed) & 0xFFFFFFFF
    cfg.keep_draws = 1 if keep_draws else 0
    for key, val in kw.items():
        if not hasattr(cfg, key):
            raise TypeError("unknown sampler option %r" % key)
        setattr(cfg, key, val)
    return cfg


class Sampler:
    """hmc_nuts_diag_e_adapt for every gene x chain of a Batch (one CTA per gene-chain)."""

    def __init__(self, batch: Batch, cfg: NutsCfg, theta_init=None):
        self.batch, self.cfg, self.lib = batch, cfg, batch.lib
        self.model = batch.model
        h = C.c_void_p()
        check(self.lib.csplotch_sampler_create(batch.h, C.byref(cfg), C.byref(h)))
        self.h = h
        self.G, self.nch = batch.G, cfg.num_chains
        if theta_init is not None:
            ti = np.ascontiguousarray(np.asarray(theta_init, dtype=np.float64).reshape(self.G, self.nch, self.model.dim))
            check(self.lib.csplotch_sampler_init(self.h, _p(ti)))

    def close(self):
        if getattr(self, "h", None):
            self.lib.csplotch_sampler_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def advance(self, n_iter: int, sync: bool = True):
        check(self.lib.csplotch_sampler_advance(self.h, C.c_int32(n_iter), C.c_int32(1 if sync else 0)))

    def sync(self):
        check(self.lib.csplotch_sampler_sync(self.h))

    def run(self, chunk: int | None = None):
        total = self.cfg.num_warmup + self.cfg.num_samples
        chunk = total if not chunk else chunk
        done = 0
        self.total_ms = 0.0                  # device time of the launches (CUDA events), for the batch report
        while done < total:
            n = min(chunk, total - done)
            self.advance(n, sync=True)
            self.total_ms += self.last_ms()
            done += n
        return self

    def last_ms(self) -> float:
        ms = C.c_float(0)
        check(self.lib.csplotch_sampler_last_ms(self.h, C.byref(ms)))
        return ms.value

    def counters(self):
        a, b, c, d = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
        check(self.lib.csplotch_sampler_counters(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(n_leapfrog=a.value, n_grad=b.value, n_divergent=c.value, n_launches=d.value)

    def profile(self):
        out = (C.c_int64 * 8)()
        check(self.lib.csplotch_sampler_profile(self.h, out))
        keys = ["cyc_eval", "cyc_merge", "cyc_copy", "cyc_other", "n_merge", "n_copy", "cyc_pro", "cyc_fin"]
        return dict(zip(keys, [int(x) for x in out]))

    def layout(self):
        """(gene-chains resident at once, CTAs of the thread-block cluster working on each)"""
        a, b = C.c_int32(), C.c_int32()
        check(self.lib.csplotch_sampler_layout(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def launch_tail(self):
        """Diagnostic: (mean, max) busy seconds over the resident CTAs of the last launch; 1 - mean / max is the
        fraction of the launch the average SM slot spent idle waiting for the longest chain."""
        n = C.c_int32()
        check(self.lib.csplotch_sampler_slot_ns(self.h, None, C.c_int32(0), C.byref(n)))
        ns = np.zeros(n.value, dtype=np.int64)
        check(self.lib.csplotch_sampler_slot_ns(self.h, _p(ns), C.c_int32(n.value), C.byref(n)))
        return float(ns.mean()) * 1e-9, float(ns.max()) * 1e-9

    def status(self):
        st = np.zeros((self.G, self.nch), dtype=np.int32)
        it = np.zeros((self.G, self.nch), dtype=np.int32)
        check(self.lib.csplotch_sampler_status(self.h, _p(st), _p(it)))
        return st, it

    def draws(self, out=None):
        """(diag [G,ch,S,7], small [G,ch,S,n_small] unconstrained beta_raw|scalars, n_draws [G,ch]);
        `out`: that triple preallocated (e.g. `pinned_empty`), filled in place"""
        ns = self.cfg.num_samples
        if out is not None:
            diag, small, nd = out
            assert diag.shape == (self.G, self.nch, ns, 7) and small.shape == (self.G, self.nch, ns, self.model.n_small)
            assert nd.shape == (self.G, self.nch) and nd.dtype == np.int32 and diag.dtype == small.dtype == np.float64
            assert diag.flags.c_contiguous and small.flags.c_contiguous and nd.flags.c_contiguous
        else:
            diag = np.empty((self.G, self.nch, ns, 7))
            small = np.empty((self.G, self.nch, ns, self.model.n_small))
            nd = np.zeros((self.G, self.nch), dtype=np.int32)
        check(self.lib.csplotch_sampler_get_draws(self.h, _p(diag), _p(small), _p(nd)))
        return diag, small, nd

    def theta_draws(self, g0: int = 0, n_genes: int | None = None):
        """every sampling draw of the unconstrained vector, (n_genes, chains, S, dim), for genes g0 .. g0 + n_genes - 1 of
        the batch (default: all — at 10^5 spots fetch a few genes at a time)"""
        n = self.G - g0 if n_genes is None else int(n_genes)
        out = np.empty((n, self.nch, self.cfg.num_samples, self.model.dim))
        check(self.lib.csplotch_sampler_get_theta_draws_range(self.h, C.c_int32(g0), C.c_int32(n), _p(out)))
        return out

    def state(self):
        th = np.empty((self.G, self.nch, self.model.dim))
        eps = np.empty((self.G, self.nch))
        mi = np.empty((self.G, self.nch, self.model.dim))
        check(self.lib.csplotch_sampler_get_state(self.h, _p(th), _p(eps), _p(mi)))
        return th, eps, mi

    def spot_summaries(self, out=None):
        """dict of (G, N) arrays: mean/std (ddof=0, pooled over chains) of log_lambda, lambda, psi.
        `out`: a preallocated float64 array (6, G, N) (e.g. `pinned_empty`) the six results are views of"""
        N = self.model.N
        if out is not None:
            assert out.shape == (6, self.G, N) and out.dtype == np.float64 and out.flags.c_contiguous
            arrs = [out[i] for i in range(6)]
        else:
            arrs = [np.empty((self.G, N)) for _ in range(6)]
        check(self.lib.csplotch_sampler_get_spot_summaries(self.h, *[_p(a) for a in arrs]))
        keys = ["log_lambda/mean", "log_lambda/std", "lambda/mean", "lambda/std", "psi/mean", "psi/std"]
        out = dict(zip(keys, arrs))
        if not self.model.car:
            out.pop("psi/mean"); out.pop("psi/std")
        return out

    # ---- constrained small block: what read_stan_csv would give for the non-spot variables ------
    def extract_small(self, genes=None, raw=None):
        """Per gene g: dict var -> array with the sample axis first, chains concatenated chain-major
        (the order of `combined_<g>.csv`, bin/splotch:138-140).  beta_level_k: (S, L_k, C[, K]).
        `genes`: indices of the genes wanted (default all; the first axis of every result then runs over them) and
        `raw`: the triple of a previous `draws()` call — at 10^5 spots x K = 10 the non-spot block of a 512-gene batch
        is tens of GB, so callers that walk over genes constrain a few at a time instead of the whole batch at once."""
        m = self.model
        diag, small, nd = self.draws() if raw is None else raw
        pm, ps = self.batch.beta_prior_mean, self.batch.beta_prior_std
        G = self.G
        if genes is not None:
            genes = np.atleast_1d(np.asarray(genes, dtype=np.int64))
            diag, small, nd = diag[genes], small[genes], nd[genes]
            if pm is not None:
                pm, ps = pm[genes], ps[genes]
            G = len(genes)
        L1, L2, L3 = m.levels
        L, Cc, K = L1 + L2 + L3, m.C, m.K
        nch, ns = self.nch, self.cfg.num_samples
        raw = small[..., : m.n_beta].reshape(G, nch * ns, L, Cc, K)
        sc = small[..., m.n_beta:].reshape(G, nch * ns, m.n_scalar)
        names = []
        if m.car:
            names += ["tau", "a"]
        names += ["sigma"]
        if L2:
            names += ["sigma_level_2"]
        if L3:
            names += ["sigma_level_3"]
        if m.zi:
            names += ["theta"]
        if m.nb:
            names += ["phi"]
        out = {}
        for i, nme in enumerate(names):
            u = sc[..., i]
            out[nme] = (1.0 / (1.0 + np.exp(-u))) if nme in ("a", "theta") else np.exp(u) + (1e-6 if nme == "phi" else 0.0)
        if m.family == 1:
            std = ps[:, None, None, None, :]
            mean = pm[:, None, None, None, :]
            b1 = raw[:, :, :L1] * std + mean
        else:
            b1 = 2.0 * raw[:, :, :L1]
        out["beta_level_1"] = b1
        if L2:
            b2 = b1[:, :, m.level_2_mapping - 1] + out["sigma_level_2"][..., None, None, None] * raw[:, :, L1:L1 + L2]
            out["beta_level_2"] = b2
        if L3:
            b3 = b2[:, :, m.level_3_mapping - 1] + out["sigma_level_3"][..., None, None, None] * raw[:, :, L1 + L2:]
            out["beta_level_3"] = b3
        if m.family != 1:
            for key in list(out):
                if key.startswith("beta_level"):
                    out[key] = out[key][..., 0]
        out["diag"] = diag.reshape(G, nch * ns, 7)
        out["n_draws"] = nd
        return out


class Fit:
    """Result of SplotchModel.sampling(): PyStan-like `extract()` for one gene."""

    def __init__(self, sampler: Sampler, g: int = 0):
        self.sampler, self.g = sampler, g
        self._small = sampler.extract_small()
        self._spots = None

    def extract(self, pars=None, permuted=True):
        out = {k: v[self.g] for k, v in self._small.items() if k not in ("diag", "n_draws")}
        out["lp__"] = self._small["diag"][self.g][:, 0]
        if self.sampler.cfg.keep_draws:
            m = self.sampler.model
            th = self.sampler.theta_draws()[self.g].reshape(-1, m.dim)
            wa = self.sampler.batch.write_array(th, gene_of=np.full(len(th), self.g, dtype=np.int32))
            cols = m.column_names()
            i0 = cols.index("log_lambda.1")
            out["log_lambda"] = wa[:, i0:i0 + m.N]
            if m.car:
                out["psi"] = wa[:, :m.N]
        if pars is not None:
            pars = [pars] if isinstance(pars, str) else list(pars)
            out = {k: out[k] for k in pars}
        return out

    def sampler_diagnostics(self):
        d = self._small["diag"][self.g]
        names = ["lp__", "accept_stat__", "stepsize__", "treedepth__", "n_leapfrog__", "divergent__", "energy__"]
        return {n: d[:, i] for i, n in enumerate(names)}


class SplotchModel:
    """Stand-in for `pystan.StanModel(file=…)` / the compiled CmdStan binary of one .stan file."""

    def __init__(self, file=None, model_name=None):
        self.family = family_from_binary(file if file is not None else model_name)

    def sampling(self, data: dict, iter: int = 2000, chains: int = 4, warmup: int | None = None, seed: int = 0,
                 keep_draws: bool = False, **control) -> Fit:
        """PyStan semantics: `iter` counts warm-up + sampling, warm-up defaults to iter // 2."""
        warmup = iter // 2 if warmup is None else warmup
        shared = {k: v for k, v in data.items() if k not in ("counts", "beta_prior_mean", "beta_prior_std")}
        model = Model(shared, self.family)
        batch = Batch(model, np.asarray(data["counts"])[None, :],
                      None if self.family != 1 else np.asarray(data["beta_prior_mean"])[None, :],
                      None if self.family != 1 else np.asarray(data["beta_prior_std"])[None, :])
        cfg = nuts_cfg(warmup, iter - warmup, chains, seed=seed, keep_draws=keep_draws, **control)
        s = Sampler(batch, cfg).run()
        return Fit(s, 0)
