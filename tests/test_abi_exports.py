"""The C-ABI library loads and exports every symbol include/csplotch.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "csplotch.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(csplotch_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    from csplotch_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.SYMBOLS) == names
    lib.csplotch_abi_version.restype = ctypes.c_int32
    assert lib.csplotch_abi_version() == 1


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product must refuse loudly rather than fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from csplotch_b200 import api, synth
    from csplotch_b200._lib import CsplotchError
    shared = synth.shared_tiny(seed=1)
    with pytest.raises(CsplotchError):
        api.Model(shared, "comp_splotch_stan_model")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "csplotch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "csplotch_oracle" not in txt, f


def test_build_is_single_threaded_and_targets_sm_100a():
    """`nvcc --split-compile 0` produced two code variants of the tile loop from one source, a third apart in speed
    (profiles/r02_kernel_experiments.md): the product build must stay reproducible."""
    from csplotch_b200 import build as b
    assert not any("split-compile" in f for f in b.NVCC_FLAGS)
    assert "arch=compute_100a,code=sm_100a" in b.NVCC_FLAGS and "-lineinfo" in b.NVCC_FLAGS
