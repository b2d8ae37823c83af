"""Output side of the drop-in boundary.

* `write_combined_csv`  — `combined_<g>.csv` exactly as bin/splotch:138-140 leaves it: the header line
  (the one containing `lp__`), then the draws of all chains, chain-major, no `#` comment lines.
* `read_stan_csv`       — the column-name -> ndarray mapping of the reference's `read_stan_csv`
  (/root/reference/splotch/utils.py:139-232): dotted 1-based indices, sample axis first.
* `summarize`           — the dictionary `summarize_stan_hdf5` writes (utils.py:234-257): per variable
  `mean`, `std` (numpy std, ddof=0) over all draws; `samples` for beta_level_*; `lambda` = exp(log_lambda).
* `write_summary`       — HDF5 when h5py imports (same group/dataset names), else an .npz with the same
  keys (`<var>/<stat>`); `read_summary` reads either.
"""
from __future__ import annotations

import gzip
import os

import numpy as np

SAMPLER_COLS = ["lp__", "accept_stat__", "stepsize__", "treedepth__", "n_leapfrog__", "divergent__", "energy__"]
PATHFINDER_COLS = ["lp_approx__", "lp__", "path__"]      # leading columns of `BINARY pathfinder` output
SUMMARY_VARS = ["log_lambda", "beta_level_1", "beta_level_2", "beta_level_3", "psi", "sigma", "sigma_level_2",
                "sigma_level_3", "theta", "tau", "a"]


def write_combined_csvs(paths, columns, diag: np.ndarray, values: np.ndarray, lead_cols=None, sig_figs: int = 0,
                        n_threads: int = 0) -> None:
    """`combined_<g>.csv` of several genes at once through the library's native writer (`csplotch_csv_write`, host
    threads, one file each).  diag: (F, S, 7) sampler columns; values: (F, S, n_outputs) in `columns` order; rows
    are chain-major.  `lead_cols`: names of the leading columns when they are not the sampler's (Pathfinder output:
    lp_approx__, lp__, path__).  CmdStan writes 6 significant digits by default (sig_figs=-1); the default here
    (sig_figs=0) is the shortest text that reads back to the same double, so that a summary made from the CSV equals
    the on-device summary; integral values print without a decimal point."""
    import ctypes as C
    from . import _lib
    lead_cols = SAMPLER_COLS if lead_cols is None else list(lead_cols)
    paths = [os.fspath(p) for p in paths]
    F = len(paths)
    diag = np.ascontiguousarray(diag, dtype=np.float64).reshape(F, -1, len(lead_cols))
    values = np.ascontiguousarray(values, dtype=np.float64).reshape(F, diag.shape[1], len(columns))
    header = ",".join(lead_cols + list(columns)).encode()
    arr = (C.c_char_p * F)(*[p.encode() for p in paths])
    _lib.check(_lib.lib().csplotch_csv_write(arr, C.c_int32(F), header, diag.ctypes.data_as(C.c_void_p),
                                             C.c_int32(diag.shape[2]), values.ctypes.data_as(C.c_void_p),
                                             C.c_int32(values.shape[2]), C.c_int64(diag.shape[1]),
                                             C.c_int32(sig_figs), C.c_int32(n_threads)))


def write_combined_csv(path: str, columns, diag: np.ndarray, values: np.ndarray, lead_cols=None, sig_figs: int = 0) -> None:
    """One file: diag (S, 7), values (S, n_outputs)."""
    assert np.shape(diag)[0] == np.shape(values)[0] and np.shape(values)[1] == len(columns)
    write_combined_csvs([path], columns, np.asarray(diag)[None], np.asarray(values)[None], lead_cols, sig_figs)


def read_stan_csv(filename: str, variables):
    """Same contract as the reference's read_stan_csv: {variable: array(S, d1, d2, ...)}; scalars (S, 1)."""
    opener = gzip.open if filename.endswith(".gz") else open
    with opener(filename, "rt") as f:
        header = None
        rows = []
        for line in f:
            if line.startswith("#") or not line.strip():
                continue
            if header is None:
                header = line.strip().split(",")
                continue
            toks = line.strip().split(",")
            if len(toks) != len(header):
                continue  # the reference skips ragged rows (utils.py:188-192)
            rows.append(toks)
    data = np.array(rows, dtype=np.float64).reshape(len(rows), len(header))
    S = data.shape[0]
    out = {}
    for v in variables:
        idx = [i for i, c in enumerate(header) if c == v or c.startswith(v + ".")]
        if not idx:
            continue
        names = [header[i] for i in idx]
        if len(names) == 1 and names[0] == v:
            out[v] = data[:, idx].reshape(S, 1)
            continue
        subs = np.array([[int(t) for t in n.split(".")[1:]] for n in names])
        shape = tuple(subs.max(0))
        arr = np.zeros((S,) + shape)
        for col, sub in zip(idx, subs):
            arr[(slice(None),) + tuple(int(x) - 1 for x in sub)] = data[:, col]
        out[v] = arr
    return out


def summarize(post: dict) -> dict:
    """{var: {'mean','std'[, 'samples']}} + 'lambda' from a read_stan_csv-style dictionary."""
    out = {}
    for key, arr in post.items():
        d = {"mean": arr.mean(axis=0), "std": arr.std(axis=0)}
        if key.startswith("beta"):
            d["samples"] = arr
        out[key] = d
    if "log_lambda" in post:
        lam = np.exp(post["log_lambda"])
        out["lambda"] = {"mean": lam.mean(axis=0), "std": lam.std(axis=0)}
    return out


def summary_from_sampler(sampler, g: int) -> dict:
    """The same dictionary built from the on-device streaming summaries (no CSV round trip)."""
    # the raw draws of the batch are fetched once; each gene is constrained on its own (bounded host memory)
    raw = sampler.draws() if not hasattr(sampler, "_raw_cache") else sampler._raw_cache
    sampler._raw_cache = raw
    ex = sampler.extract_small(genes=[g], raw=raw)
    sp = sampler.spot_summaries() if not hasattr(sampler, "_sp_cache") else sampler._sp_cache
    sampler._sp_cache = sp
    out = {}
    for key in ("log_lambda", "lambda", "psi"):
        if key + "/mean" in sp:
            out[key] = {"mean": sp[key + "/mean"][g], "std": sp[key + "/std"][g]}
    for key in SUMMARY_VARS:
        if key in ex:
            arr = ex[key][0]
            if arr.ndim == 1:
                arr = arr[:, None]
            d = {"mean": arr.mean(axis=0), "std": arr.std(axis=0)}
            if key.startswith("beta"):
                d["samples"] = arr
            out[key] = d
    return out


def write_summary(path: str, summ: dict) -> str:
    """HDF5 (`combined_<g>.hdf5`) when h5py is importable, else `<path>.npz` with '<var>/<stat>' keys."""
    try:
        import h5py  # noqa: F401
    except ImportError:
        h5py = None
    if h5py is not None:
        tmp = path + ".tmp"
        with h5py.File(tmp, "w") as h:
            for key, d in summ.items():
                for stat, arr in d.items():
                    h.create_dataset("%s/%s" % (key, stat), data=arr)
        os.replace(tmp, path)
        return path
    flat = {"%s/%s" % (k, s): a for k, d in summ.items() for s, a in d.items()}
    out = path if path.endswith(".npz") else path + ".npz"
    tmp = out + ".tmp.npz"
    np.savez(tmp, **flat)
    os.replace(tmp, out)
    return out


def read_summary(path: str) -> dict:
    if os.path.exists(path) and not path.endswith(".npz"):
        import h5py
        out = {}
        with h5py.File(path, "r") as h:
            for key in h:
                out[key] = {stat: h[key][stat][:] for stat in h[key]}
        return out
    p = path if path.endswith(".npz") else path + ".npz"
    z = np.load(p)
    out = {}
    for k in z.files:
        var, stat = k.split("/")
        out.setdefault(var, {})[stat] = z[k]
    return out
