"""R-dump reader / writer for the Stan data files `data_<g>.R`.

Semantics follow the reference's `read_rdump` / `to_rdump`
(/root/reference/splotch/utils.py:78-125; grammar in SURVEY.md App. D):

* scalar        ``key <- v``
* 1-D           ``key <-\\n c(v1,v2,...)``            (empty: ``c()``)
* N-D           ``key <-\\n structure(c(<column-major>), .Dim=c(d1,d2))``

The reader is whitespace tolerant (the Science/human/data_SNAP25.R fixture uses
``", "`` separators and ``.Dim = c(``) and vectorised: the reference parses token
by token in Python, which dominates end-to-end time at 10^5 spots.
`read_rdump_gene_only` is the ingestion fast path: it extracts only the
gene-specific keys (`counts`, `beta_prior_mean`, `beta_prior_std`) from a file
whose shared covariates are already known.
"""
from __future__ import annotations

import os
import re

import numpy as np

_KEY_RE = re.compile(r"^[ \t]*([A-Za-z_][A-Za-z0-9_.]*)[ \t]*<-", re.M)
_STRUCT_RE = re.compile(r"structure\s*\(\s*c\s*\((.*)\)\s*,\s*\.Dim\s*=\s*c\s*\((.*?)\)\s*\)\s*$", re.S)
_VEC_RE = re.compile(r"c\s*\((.*)\)\s*$", re.S)

GENE_KEYS = ("counts", "beta_prior_mean", "beta_prior_std")


def _parse_numbers(txt: str) -> np.ndarray:
    txt = txt.strip()
    if not txt:
        return np.zeros(0, dtype=np.int64)
    if "L" in txt:                      # R's own dump() marks integers as `3L`; to_rdump never does
        txt = txt.replace("L", "")
    integral = not any(ch in txt for ch in ".eEnN")  # no '.', exponent, nan, inf
    toks = txt.split(",")
    if integral:
        return np.array(toks, dtype=np.int64)
    return np.array(toks, dtype=np.float64)


def _parse_value(txt: str):
    txt = txt.strip()
    m = _STRUCT_RE.match(txt)
    if m:
        flat = _parse_numbers(m.group(1))
        shape = tuple(int(t) for t in m.group(2).split(",") if t.strip())
        # R stores column-major
        return flat.reshape(tuple(reversed(shape))).T
    m = _VEC_RE.match(txt)
    if m:
        return _parse_numbers(m.group(1))
    v = _parse_numbers(txt)
    return v[0].item()


def read_rdump(filename: str, keys=None) -> dict:
    """Parse an R-dump file into {key: python scalar | ndarray}. `keys` restricts the keys parsed."""
    with open(filename, "r") as f:
        text = f.read()
    out = {}
    ms = list(_KEY_RE.finditer(text))
    for i, m in enumerate(ms):
        key = m.group(1)
        if keys is not None and key not in keys:
            continue
        end = ms[i + 1].start() if i + 1 < len(ms) else len(text)
        out[key] = _parse_value(text[m.end():end])
    return out


def read_rdump_gene_only(filename: str) -> dict:
    """Fast path: only the keys that differ between genes (bin/splotch_generate_input_files:329-346)."""
    return read_rdump(filename, keys=GENE_KEYS)


def read_genes_native(paths, n_spots: int, n_celltypes: int = 0, n_threads: int = 0):
    """`counts` [G, N] (and `beta_prior_mean`, `beta_prior_std` [G, K] when n_celltypes > 0) of many data_<g>.R
    files through the library's native reader (csplotch_rdump_read_genes, csrc/ingest.cpp): host threads, one
    file each, only the tail of every file is read.  Same values as `read_rdump_gene_only` file by file."""
    import ctypes as C

    from . import _lib
    lib = _lib.lib()
    G = len(paths)
    counts = np.empty((G, n_spots), dtype=np.int32)
    pm = ps = None
    if n_celltypes > 0:
        pm = np.empty((G, n_celltypes))
        ps = np.empty((G, n_celltypes))
    arr = (C.c_char_p * G)(*[os.fsencode(p) for p in paths])
    vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None  # noqa: E731
    _lib.check(lib.csplotch_rdump_read_genes(arr, C.c_int32(G), C.c_int32(n_spots), C.c_int32(n_celltypes),
                                            vp(counts), vp(pm), vp(ps), C.c_int32(n_threads)))
    out = {"counts": counts}
    if pm is not None:
        out["beta_prior_mean"], out["beta_prior_std"] = pm, ps
    return out


def _fmt(v) -> str:
    if isinstance(v, (bool, np.bool_)):
        return str(int(v))
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    return repr(float(v))


def to_rdump(data: dict, filename: str) -> None:
    """Write `data` exactly in the layout of the reference's `to_rdump` (utils.py:113-125)."""
    with open(filename, "w") as f:
        for key, val in data.items():
            arr = np.asarray(val)
            if arr.ndim == 0:
                f.write("%s <- %s\n" % (key, _fmt(arr.item())))
            elif arr.ndim == 1:
                f.write("%s <-\n c(%s)\n" % (key, ",".join(_fmt(x) for x in arr.tolist())))
            else:
                flat = arr.flatten("F").tolist()
                f.write("%s <-\n structure(c(%s), .Dim=c(%s))\n"
                        % (key, ",".join(_fmt(x) for x in flat), ",".join(str(s) for s in arr.shape)))
