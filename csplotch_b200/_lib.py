"""ctypes binding of libcsplotch_b200.so (the C ABI in include/csplotch.h).

There is no CPU fallback: if the CUDA library is missing or no device is visible the
product fails loudly."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# CSPLOTCH_LIB: another build of the same library (development builds of csplotch_b200.build --dev)
LIB_PATH = os.environ.get("CSPLOTCH_LIB") or os.path.join(HERE, "libcsplotch_b200.so")


class CsplotchError(RuntimeError):
    pass


class SharedDesc(C.Structure):
    _fields_ = [
        ("family", C.c_int32), ("N_tissues", C.c_int32), ("N_spots", C.c_void_p),
        ("N_covariates", C.c_int32), ("N_celltypes", C.c_int32),
        ("N_level_1", C.c_int32), ("N_level_2", C.c_int32), ("N_level_3", C.c_int32),
        ("zi", C.c_int32), ("car", C.c_int32), ("nb", C.c_int32),
        ("level_2_mapping", C.c_void_p), ("level_3_mapping", C.c_void_p), ("tissue_mapping", C.c_void_p),
        ("size_factors", C.c_void_p), ("D", C.c_void_p), ("E", C.c_void_p),
        ("W_n", C.c_int32), ("W_sparse", C.c_void_p), ("U", C.c_void_p), ("V", C.c_void_p),
        ("D_sparse", C.c_void_p), ("eig_values", C.c_void_p),
    ]


class NutsCfg(C.Structure):
    _fields_ = [
        ("num_warmup", C.c_int32), ("num_samples", C.c_int32), ("num_chains", C.c_int32),
        ("max_depth", C.c_int32), ("delta", C.c_double), ("gamma", C.c_double), ("kappa", C.c_double),
        ("t0", C.c_double), ("init_buffer", C.c_int32), ("term_buffer", C.c_int32), ("window", C.c_int32),
        ("init_radius", C.c_double), ("stepsize", C.c_double), ("seed", C.c_uint32), ("keep_draws", C.c_int32),
    ]


class Dims(C.Structure):
    _fields_ = [("N", C.c_int32), ("dim", C.c_int32), ("n_beta", C.c_int32), ("n_scalar", C.c_int32),
                ("n_small", C.c_int32), ("n_outputs", C.c_int32)]


# every symbol include/csplotch.h declares
SYMBOLS = [
    "csplotch_last_error", "csplotch_abi_version", "csplotch_device_count", "csplotch_set_device",
    "csplotch_model_create", "csplotch_model_destroy", "csplotch_model_dims",
    "csplotch_batch_create", "csplotch_batch_create_dev", "csplotch_batch_destroy", "csplotch_batch_set_gene_ids",
    "csplotch_log_prob_grad", "csplotch_log_prob_grad_dev", "csplotch_write_array", "csplotch_leapfrog_fixed", "csplotch_leapfrog_bench",
    "csplotch_nuts_cfg_default", "csplotch_sampler_create", "csplotch_sampler_destroy", "csplotch_sampler_init",
    "csplotch_sampler_advance", "csplotch_sampler_sync", "csplotch_sampler_last_ms", "csplotch_sampler_counters",
    "csplotch_sampler_status", "csplotch_sampler_slot_ns", "csplotch_sampler_layout", "csplotch_sampler_profile", "csplotch_sampler_get_draws", "csplotch_sampler_get_theta_draws", "csplotch_sampler_get_theta_draws_range",
    "csplotch_sampler_get_state", "csplotch_sampler_get_spot_summaries", "csplotch_rdump_read_genes", "csplotch_host_alloc", "csplotch_host_free",
    "csplotch_csv_write", "csplotch_car_precompute",
]

_lib = None


def lib():
    """Load the CUDA library; raise (never fall back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CsplotchError(
                "csplotch_b200: %s not found. Build it with `python -m csplotch_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        l.csplotch_last_error.restype = C.c_char_p
        for s in SYMBOLS:
            try:
                getattr(l, s)
            except AttributeError:
                # only an explicitly selected build (CSPLOTCH_LIB: development / older variants for A-B runs) may lack
                # entry points; the product library must export every symbol include/csplotch.h declares
                if not os.environ.get("CSPLOTCH_LIB"):
                    raise
        _lib = l
    return _lib


def check(rc: int):
    if rc != 0:
        raise CsplotchError("csplotch error %d: %s" % (rc, lib().csplotch_last_error().decode()))


def require_gpu():
    n = lib().csplotch_device_count()
    if n < 1:
        raise CsplotchError("csplotch_b200 needs a CUDA device (sm_100a); none is visible. There is no CPU fallback.")
    return n
