"""TEST INFRASTRUCTURE: host build of the product's sampler control logic (nuts_core.h)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIB = os.path.join(_HERE, "libemul.so")


class NutsConfig(C.Structure):
    _fields_ = [("num_warmup", C.c_int32), ("num_samples", C.c_int32), ("max_depth", C.c_int32),
                ("delta", C.c_double), ("gamma", C.c_double), ("kappa", C.c_double), ("t0", C.c_double),
                ("init_buffer", C.c_int32), ("term_buffer", C.c_int32), ("window", C.c_int32),
                ("init_radius", C.c_double), ("stepsize", C.c_double), ("seed", C.c_uint32)]


def build(merge2: bool = True):
    """CSPLOTCH_EMUL_SAN=1: the same translation unit under AddressSanitizer + UndefinedBehaviorSanitizer (the vector
    pool, the binary-counter stack and the window arithmetic of nuts_core.h are product code).
    merge2=False: the U-turn cascade taken one merge per pass (libemul_m1.so), the product's -DCSPLOTCH_NO_MERGE2."""
    san = os.environ.get("CSPLOTCH_EMUL_SAN") == "1"
    out = os.path.join(_HERE, "libemul_san.so") if san else (_LIB if merge2 else os.path.join(_HERE, "libemul_m1.so"))
    srcs = [os.path.join(_HERE, "emul.cpp"), os.path.join(ROOT, "csplotch_b200", "csrc", "nuts_core.h"),
            os.path.join(ROOT, "csplotch_b200", "csrc", "philox.h")]
    if (not os.path.exists(out)) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        flags = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                 "-fno-omit-frame-pointer"] if san else ["-O2"]
        subprocess.check_call(["/usr/bin/g++"] + flags + ([] if merge2 else ["-DEMUL_MERGE2=0"]) + ["-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                               "-I" + os.path.join(ROOT, "csplotch_b200", "csrc"), "-o", out, srcs[0]])
    return out


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.emul_nuts_run.restype = C.c_int
    return _lib


_lib_m1 = None


def lib_m1():
    """the one-merge-per-pass build (tests of the merge2 cascade)"""
    global _lib_m1
    if _lib_m1 is None:
        _lib_m1 = C.CDLL(build(merge2=False))
        _lib_m1.emul_nuts_run.restype = C.c_int
    return _lib_m1


def nuts_run(fn_ptr, ctx_ptr, D, ocfg, gene_id=0, chain_id=1, q_init=None, chunk=0, merge2=True):
    """fn_ptr: C function pointer double(void*, const double*, double*); ocfg: oracle._NutsCfg"""
    cfg = NutsConfig(ocfg.num_warmup, ocfg.num_samples, ocfg.max_depth, ocfg.delta, ocfg.gamma, ocfg.kappa,
                     ocfg.t0, ocfg.init_buffer, ocfg.term_buffer, ocfg.window, ocfg.init_radius,
                     ocfg.stepsize, ocfg.seed)
    draws = np.zeros((ocfg.num_samples, 7 + D))
    inv_m = np.empty(D)
    eps = C.c_double(0)
    nl = C.c_longlong(0)
    ng = C.c_longlong(0)
    qi = np.ascontiguousarray(q_init, dtype=np.float64) if q_init is not None else None
    st = (lib() if merge2 else lib_m1()).emul_nuts_run(fn_ptr, ctx_ptr, C.c_int(D), C.byref(cfg), C.c_uint32(gene_id),
                             C.c_uint32(chain_id), qi.ctypes.data_as(C.c_void_p) if qi is not None else None,
                             C.c_int(chunk), draws.ctypes.data_as(C.c_void_p), inv_m.ctypes.data_as(C.c_void_p),
                             C.byref(eps), C.byref(nl), C.byref(ng))
    return dict(status=st, draws=draws, inv_metric=inv_m, stepsize=eps.value, n_leapfrog=nl.value,
                n_grad=ng.value)
