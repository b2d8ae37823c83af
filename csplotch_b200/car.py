"""CAR pre-compute on the GPU: eigenvalues of D^-1/2 W D^-1/2 per tissue, D_sparse, CSR U / V.

Stands in for the CAR block of `generate_dictionary` (/root/reference/splotch/utils.py:735-765) — what
`splotch_generate_input_files` spends most of its time on for Visium-sized inputs.  Everything numerical happens in
`csplotch_car_precompute` (csrc/car_precompute.cu): dense matrices assembled on the device, cuSOLVER syevd, identical
tissues / chunks solved once.  `api.Model` calls it when a data dictionary arrives without `eig_values`."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check


def parse_coords(coords_str) -> np.ndarray:
    """"x_y" strings of integer coordinates (the reference's coordinate labels, utils.py:646) -> (N, 2) int32"""
    return np.array([tuple(int(v) for v in str(c).split("_")) for c in coords_str], dtype=np.int32).reshape(-1, 2)


def precompute(N_spots, W_sparse, coords=None, chunk_spots: int = 10000, exact_max_n: int = 9999, want_eig: bool = True):
    """-> dict with the CAR keys of the Stan data dictionary: W_n, W_sparse, D_sparse, eig_values, U, V (+ n_solves).

    N_spots: spots per tissue; W_sparse: (W_n, 2) 1-based pairs, each once; coords: (N, 2) integer coordinates or a
    list of "x_y" strings, only needed for tissues with more than `exact_max_n` spots (the reference chunks tissues of
    10 000 spots and more; raise `exact_max_n`, up to 40 000, to solve larger ones exactly instead)."""
    lib = _lib.lib()
    _lib.require_gpu()
    ns = np.ascontiguousarray(np.atleast_1d(np.asarray(N_spots, dtype=np.int32)))
    W = np.ascontiguousarray(np.asarray(W_sparse, dtype=np.int32).reshape(-1, 2))
    N, W_n = int(ns.sum()), int(W.shape[0])
    cd = None
    if coords is not None:
        cd = np.asarray(coords)
        if cd.dtype.kind in "US" or cd.dtype == object:
            cd = parse_coords(cd)
        cd = np.ascontiguousarray(cd.astype(np.int32).reshape(-1, 2))
        if cd.shape[0] != N:
            raise _lib.CsplotchError("coords must have one row per spot")
    eig = np.empty(N) if want_eig else None
    U = np.empty(N + 1, dtype=np.int32)
    V = np.empty(2 * W_n, dtype=np.int32)
    D = np.empty(N)
    n_solves = C.c_int32(0)

    def p(a):
        return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None

    check(lib.csplotch_car_precompute(C.c_int32(len(ns)), p(ns), C.c_int32(W_n), p(W), p(cd), C.c_int32(chunk_spots),
                                      C.c_int32(exact_max_n), p(eig), p(U), p(V), p(D), C.byref(n_solves)))
    return {"W_n": np.array([W_n], dtype=np.int64), "W_sparse": W.astype(np.int64), "D_sparse": D.astype(np.int64),
            "eig_values": eig, "U": U.astype(np.int64), "V": V.astype(np.int64), "n_solves": int(n_solves.value)}
