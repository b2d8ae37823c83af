"""In-tree build of the CUDA library (sm_100a only).  `python -m csplotch_b200.build`."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "csplotch.cu")
HOST_SRC = os.path.join(HERE, "csrc", "ingest.cpp")   # host-only part of the ABI (R-dump fast path)
CAR_SRC = os.path.join(HERE, "csrc", "car_precompute.cu")   # CAR pre-compute (cuSOLVER through dlopen)
DEPS = [SRC, HOST_SRC, CAR_SRC] + [os.path.join(HERE, "csrc", f) for f in ("device_ops.cuh", "nuts_core.h", "philox.h", "fastmath.h")] + [
    os.path.join(ROOT, "include", "csplotch.h")]
OUT = os.path.join(HERE, "libcsplotch_b200.so")

NVCC_FLAGS = [
    "-ccbin", "/usr/bin/g++",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # NO --split-compile: with `--split-compile 0` (optimise on all host cores, 2x faster to build) the output for this
    # translation unit is NOT deterministic — three sequential builds of one source gave 944 / 944 / 912-byte frames for
    # k_leapfrog_fixed<10>, and the 944-byte variant runs the c3 leapfrog at 0.27 of the HBM roof against 0.40
    # (profiles/r02_kernel_experiments.md, GPU call 20).  The single-threaded default is reproducible and the fastest.
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc_path() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def is_stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


DEV_OUT = os.path.join(HERE, "libcsplotch_b200_dev.so")


def build(force: bool = False, verbose: bool = False, dev: bool = False, out: str | None = None) -> str:
    """dev=True: K in {1, 10} only, written to libcsplotch_b200_dev.so (or `out`; experiments: load it with
    CSPLOTCH_LIB=<path>, extra -D flags from CSPLOTCH_DEV_DEFS); the product library always holds every instantiation."""
    if not dev and not force and not is_stale():
        return OUT
    out = (out or DEV_OUT) if dev else OUT
    cmd = ([nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + (["-DCSPLOTCH_DEV_KT"] + os.environ.get("CSPLOTCH_DEV_DEFS", "").split() if dev else [])
           + ["-o", out, SRC, HOST_SRC, CAR_SRC, "-ldl"])
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    _out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, dev="--dev" in sys.argv, out=_out))
