"""csrc/fastmath.h on the host (same source as the kernels): exp_fast against libm, the running-product log."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emul", "fastmath_host.cpp")
HDR = os.path.join(ROOT, "csplotch_b200", "csrc", "fastmath.h")
LIB = os.path.join(ROOT, "tests", "emul", "libfastmath.so")


def _lib():
    if not os.path.exists(LIB) or max(os.path.getmtime(SRC), os.path.getmtime(HDR)) > os.path.getmtime(LIB):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                               "-I" + os.path.join(ROOT, "csplotch_b200", "csrc"), "-o", LIB, SRC])
    l = C.CDLL(LIB)
    l.fm_logprod.restype = C.c_double
    return l


def _exp_fast(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    _lib().fm_exp(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_long(x.size))
    return y


def test_exp_fast_within_two_ulp():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-708, 709, 2_000_000), rng.uniform(-40, 40, 1_500_000), rng.normal(0, 1, 500_000),
                        np.linspace(-1, 1, 100_001), np.arange(-1000, 1000) * np.log(2) * 0.5,
                        [0.0, -0.0, 707.999, -707.999, 1e-300, -1e-300, 5e-324]])
    got, want = _exp_fast(x), np.exp(x)
    ulp = np.spacing(want)
    err = np.abs(got - want) / ulp
    assert err.max() <= 2.0, (err.max(), x[err.argmax()])
    assert np.mean(err <= 1.0) > 0.999


def test_exp_fast_special_values():
    with np.errstate(over="ignore", under="ignore"):
        x = np.array([710.0, 1e6, np.inf, -746.0, -1e6, -np.inf, -720.0, 708.5, -708.5])
        got, want = _exp_fast(x), np.exp(x)
    assert np.array_equal(got, want)
    assert np.isnan(_exp_fast(np.array([np.nan]))[0])


def test_running_product_log():
    rng = np.random.default_rng(1)
    l = _lib()
    for n, lo in ((1, 0.3), (500, 1e-6), (500, 1e-250), (4000, 0.5)):
        f = np.exp(rng.uniform(np.log(lo), np.log(2.0), n))
        got = l.fm_logprod(f.ctypes.data_as(C.c_void_p), C.c_long(n))
        want = float(np.sum(np.log(f.astype(np.longdouble))))
        assert abs(got - want) <= 1e-13 * max(1.0, abs(want)) + 4e-16 * n, (n, lo, got, want)
