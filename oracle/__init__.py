"""TEST INFRASTRUCTURE ONLY — ctypes binding of the CPU oracle (oracle/csplotch_oracle.c).

PARITY UNPINNED at the Stan boundary (no CmdStan in this image; the reference
ships no golden vectors for this path — SURVEY.md §8c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package.  The product
(csplotch_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libcsplotch_oracle.so")

FAMILIES = {
    "splotch_stan_model": 0,
    "comp_splotch_stan_model": 1,
    "splotch_stan_model_csr": 2,
}


def build(force: bool = False) -> str:
    """Compile the C restatement with the committed Makefile."""
    src = os.path.join(_HERE, "csplotch_oracle.c")
    hdr = os.path.join(_HERE, "csplotch_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr)
    )
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


def build_fast() -> str:
    """The timing build for bench.py's CPU baseline (`make fast`: -O3 -march=native -DORC_FAST).  Always rebuilt on the
    machine that runs it (-march=native code of another host may not execute here); select it with
    CSPLOTCH_ORACLE_LIB=<returned path> before the first oracle call.  Never used for parity."""
    subprocess.check_call(["make", "-s", "-B", "-C", _HERE, "fast"])
    return os.path.join(_HERE, "_build", "libcsplotch_oracle_fast.so")


class _OrcData(C.Structure):
    _fields_ = [
        ("family", C.c_int32),
        ("N_tissues", C.c_int32),
        ("N_spots", C.c_void_p),
        ("N_covariates", C.c_int32),
        ("N_celltypes", C.c_int32),
        ("N_level_1", C.c_int32),
        ("N_level_2", C.c_int32),
        ("N_level_3", C.c_int32),
        ("zi", C.c_int32),
        ("car", C.c_int32),
        ("nb", C.c_int32),
        ("level_2_mapping", C.c_void_p),
        ("level_3_mapping", C.c_void_p),
        ("tissue_mapping", C.c_void_p),
        ("size_factors", C.c_void_p),
        ("D", C.c_void_p),
        ("E", C.c_void_p),
        ("W_n", C.c_int32),
        ("W_sparse", C.c_void_p),
        ("U", C.c_void_p),
        ("V", C.c_void_p),
        ("D_sparse", C.c_void_p),
        ("eig_values", C.c_void_p),
        ("counts", C.c_void_p),
        ("beta_prior_mean", C.c_void_p),
        ("beta_prior_std", C.c_void_p),
    ]


class _NutsCfg(C.Structure):
    _fields_ = [
        ("num_warmup", C.c_int32),
        ("num_samples", C.c_int32),
        ("max_depth", C.c_int32),
        ("delta", C.c_double),
        ("gamma", C.c_double),
        ("kappa", C.c_double),
        ("t0", C.c_double),
        ("init_buffer", C.c_int32),
        ("term_buffer", C.c_int32),
        ("window", C.c_int32),
        ("init_radius", C.c_double),
        ("stepsize", C.c_double),
        ("seed", C.c_uint32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        alt = os.environ.get("CSPLOTCH_ORACLE_LIB")       # e.g. the sanitizer build (`make -C oracle sanitize`)
        if not alt:
            build()
        _lib = C.CDLL(alt or _LIB_PATH)
        _lib.orc_log_prob_grad.restype = C.c_double
        _lib.orc_nuts_batch.restype = C.c_int64
        _lib.orc_last_batch_busy_seconds.restype = C.c_double
        _lib.orc_dim.restype = C.c_int32
        _lib.orc_num_spots.restype = C.c_int32
        _lib.orc_num_outputs.restype = C.c_int32
        _lib.orc_nuts_run.restype = C.c_int32
        _lib.orc_nuts_generic.restype = C.c_int32
    return _lib


def nuts_cfg(num_warmup, num_samples, seed=1, max_depth=10, delta=0.8, gamma=0.05, kappa=0.75, t0=10.0,
             init_buffer=75, term_buffer=50, window=25, init_radius=2.0, stepsize=1.0):
    return _NutsCfg(num_warmup, num_samples, max_depth, delta, gamma, kappa, t0, init_buffer, term_buffer,
                    window, init_radius, stepsize, seed)


def _i32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32).reshape(-1))


def _f64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class Oracle:
    """One gene's Stan data block (R-dump keys, 1-based indices) bound to the C oracle."""

    def __init__(self, data: dict, family: str | int):
        fam = FAMILIES[family] if isinstance(family, str) else int(family)
        self.family = fam
        d = data
        keep = {}
        keep["N_spots"] = _i32(np.atleast_1d(d["N_spots"]))
        self.N = int(keep["N_spots"].sum())
        self.K = int(d.get("N_celltypes", 1)) if fam == 1 else 1
        car = int(d.get("car", 0))
        keep["l2"] = _i32(d.get("level_2_mapping", []))
        keep["l3"] = _i32(d.get("level_3_mapping", []))
        keep["tm"] = _i32(np.atleast_1d(d["tissue_mapping"]))
        keep["sf"] = _f64(d["size_factors"])
        keep["D"] = _i32(d["D"])
        keep["E"] = _f64(np.asarray(d["E"], dtype=np.float64).reshape(self.N, self.K)) if fam == 1 else None
        W_n = int(np.atleast_1d(d.get("W_n", [0]))[0]) if car else 0
        keep["W"] = None
        keep["U"] = keep["V"] = None
        if car:
            if "W_sparse" in d and np.size(d["W_sparse"]):
                keep["W"] = _i32(np.asarray(d["W_sparse"]).reshape(W_n, 2))
            if "U" in d and np.size(d["U"]):
                keep["U"] = _i32(d["U"])
                keep["V"] = _i32(d["V"])
            keep["Ds"] = _f64(d["D_sparse"])
            keep["eig"] = _f64(d["eig_values"])
        else:
            keep["Ds"] = keep["eig"] = None
        keep["counts"] = _i32(d["counts"])
        keep["pm"] = _f64(d["beta_prior_mean"]) if fam == 1 else None
        keep["ps"] = _f64(d["beta_prior_std"]) if fam == 1 else None
        if fam in (0, 1) and car:
            assert keep["W"] is not None, "families 0/1 need W_sparse"
        if fam == 2 and car:
            assert keep["U"] is not None, "family 2 needs U, V"
        self._keep = keep
        self.c = _OrcData(
            fam, int(d["N_tissues"]), _ptr(keep["N_spots"]), int(d["N_covariates"]), self.K,
            int(d["N_level_1"]), int(d.get("N_level_2", 0)), int(d.get("N_level_3", 0)),
            int(d.get("zi", 0)), car, int(d.get("nb", 0)) if fam == 1 else 0,
            _ptr(keep["l2"]), _ptr(keep["l3"]), _ptr(keep["tm"]), _ptr(keep["sf"]), _ptr(keep["D"]),
            _ptr(keep["E"]), W_n, _ptr(keep["W"]), _ptr(keep["U"]), _ptr(keep["V"]), _ptr(keep["Ds"]),
            _ptr(keep["eig"]), _ptr(keep["counts"]), _ptr(keep["pm"]), _ptr(keep["ps"]),
        )
        self.dim = lib().orc_dim(C.byref(self.c))
        self.n_out = lib().orc_num_outputs(C.byref(self.c))

    # -- density ---------------------------------------------------------
    def log_prob_grad(self, theta, jacobian=True, want_grad=True):
        th = _f64(theta)
        assert th.size == self.dim
        g = np.empty(self.dim) if want_grad else None
        lp = lib().orc_log_prob_grad(C.byref(self.c), _ptr(th), _ptr(g) if want_grad else None,
                                     C.c_int32(1 if jacobian else 0))
        return (lp, g) if want_grad else lp

    def write_array(self, theta):
        th = _f64(theta)
        out = np.empty(self.n_out)
        lib().orc_write_array(C.byref(self.c), _ptr(th), _ptr(out))
        return out

    def unconstrain(self, params):
        p = _f64(params)
        th = np.empty(self.dim)
        lib().orc_unconstrain(C.byref(self.c), _ptr(p), _ptr(th))
        return th

    def leapfrog_fixed(self, theta0, p0, inv_metric, eps, n_steps, trajectory=False):
        th, p, m = _f64(theta0), _f64(p0), _f64(inv_metric)
        tho, po = np.empty(self.dim), np.empty(self.dim)
        H = C.c_double(0)
        traj = np.empty((n_steps + 1, 2 * self.dim + 1)) if trajectory else None
        lib().orc_leapfrog_fixed(C.byref(self.c), _ptr(th), _ptr(p), _ptr(m), C.c_double(eps),
                                 C.c_int32(n_steps), _ptr(tho), _ptr(po), C.byref(H),
                                 _ptr(traj) if trajectory else None)
        return (tho, po, H.value, traj) if trajectory else (tho, po, H.value)

    # -- sampler -----------------------------------------------------------
    def nuts(self, cfg, gene_id=0, chain_id=1, q_init=None, save_warmup=False):
        rows = cfg.num_samples + (cfg.num_warmup if save_warmup else 0)
        draws = np.zeros((rows, 7 + self.dim))
        inv_m = np.empty(self.dim)
        eps = C.c_double(0)
        ng = C.c_int64(0)
        qi = _f64(q_init) if q_init is not None else None
        st = lib().orc_nuts_run(C.byref(self.c), C.byref(cfg), C.c_uint32(gene_id), C.c_uint32(chain_id),
                                _ptr(qi) if qi is not None else None, C.c_int32(1 if save_warmup else 0),
                                _ptr(draws), _ptr(inv_m), C.byref(eps), C.byref(ng))
        return dict(status=st, draws=draws, inv_metric=inv_m, stepsize=eps.value, n_grad=ng.value)

    # -- gene-parallel batches (CPU baseline) -------------------------------
    def nuts_batch(self, counts, cfg, n_chains=4, prior_mean=None, prior_std=None, gene_id0=0,
                   n_threads=0, want_draws=True, time_budget=0.0):
        """time_budget > 0 (bench.py's bounded CPU samples): every chain stops at its next leapfrog once that many
        seconds have passed; n_grad counts what was done."""
        counts = np.ascontiguousarray(counts, dtype=np.int32).reshape(-1, self.N)
        G = counts.shape[0]
        pm = _f64(prior_mean) if prior_mean is not None else None
        ps = _f64(prior_std) if prior_std is not None else None
        draws = np.zeros((G, n_chains, cfg.num_samples, 7 + self.dim)) if want_draws else None
        status = np.zeros(G * n_chains, dtype=np.int32)
        lib().orc_set_time_budget(C.c_double(time_budget))
        ng = lib().orc_nuts_batch(C.byref(self.c), _ptr(counts), _ptr(pm) if pm is not None else None,
                                  _ptr(ps) if ps is not None else None, C.c_int32(G), C.c_int32(n_chains),
                                  C.byref(cfg), C.c_uint32(gene_id0), C.c_int32(n_threads),
                                  _ptr(draws) if want_draws else None, _ptr(status))
        lib().orc_set_time_budget(C.c_double(0.0))
        return dict(draws=draws, status=status.reshape(G, n_chains), n_grad=int(ng),
                    busy_seconds=float(lib().orc_last_batch_busy_seconds()))

    def log_prob_grad_batch(self, counts, theta, gene_of=None, prior_mean=None, prior_std=None,
                            jacobian=True, n_threads=0):
        counts = np.ascontiguousarray(counts, dtype=np.int32).reshape(-1, self.N)
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1, self.dim)
        B = theta.shape[0]
        go = _i32(gene_of) if gene_of is not None else None
        pm = _f64(prior_mean) if prior_mean is not None else None
        ps = _f64(prior_std) if prior_std is not None else None
        lp = np.empty(B)
        grad = np.empty((B, self.dim))
        lib().orc_log_prob_grad_batch(C.byref(self.c), _ptr(counts), _ptr(pm) if pm is not None else None,
                                      _ptr(ps) if ps is not None else None, C.c_int32(counts.shape[0]),
                                      _ptr(theta), C.c_int32(B), _ptr(go) if go is not None else None,
                                      _ptr(lp), _ptr(grad), C.c_int32(1 if jacobian else 0),
                                      C.c_int32(n_threads))
        return lp, grad


def philox(key0, key1, c0, c1, c2, c3):
    out = (C.c_uint32 * 4)()
    lib().orc_philox(C.c_uint32(key0), C.c_uint32(key1), C.c_uint32(c0), C.c_uint32(c1), C.c_uint32(c2),
                     C.c_uint32(c3), out)
    return tuple(int(x) for x in out)


_DENSITY_FN = C.CFUNCTYPE(C.c_double, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double))


def nuts_generic(pyfn, D, cfg, gene_id=0, chain_id=1, q_init=None, save_warmup=False):
    """Run the oracle's sampler on an arbitrary python density pyfn(q)->(lp, grad) (tests only)."""

    def _cb(_ctx, qp, gp):
        q = np.ctypeslib.as_array(qp, shape=(D,))
        lp, g = pyfn(q.copy())
        np.ctypeslib.as_array(gp, shape=(D,))[:] = g
        return float(lp)

    cb = _DENSITY_FN(_cb)
    rows = cfg.num_samples + (cfg.num_warmup if save_warmup else 0)
    draws = np.zeros((rows, 7 + D))
    inv_m = np.empty(D)
    eps = C.c_double(0)
    ng = C.c_int64(0)
    qi = _f64(q_init) if q_init is not None else None
    st = lib().orc_nuts_generic(cb, None, C.c_int32(D), C.byref(cfg), C.c_uint32(gene_id),
                                C.c_uint32(chain_id), _ptr(qi) if qi is not None else None,
                                C.c_int32(1 if save_warmup else 0), _ptr(draws), _ptr(inv_m), C.byref(eps),
                                C.byref(ng))
    return dict(status=st, draws=draws, inv_metric=inv_m, stepsize=eps.value, n_grad=ng.value)
