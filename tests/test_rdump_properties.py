"""Property tests of the R-dump boundary: whatever `to_rdump` writes (the reference's layout,
/root/reference/splotch/utils.py:113-125) reads back to the same keys, shapes and values through the Python reader and —
for the per-gene keys — through the native tail reader, for ragged / empty / extreme inputs and for the spelling variants
other writers use (`", "`, `.Dim = c(`, `3L`, trailing blank lines, keys in any order)."""
import os

import numpy as np
from hypothesis import given, settings, strategies as st

from csplotch_b200 import rdump

ints = st.integers(-2**31 + 1, 2**31 - 1)
floats = st.floats(allow_nan=False, allow_infinity=True, width=64)


@st.composite
def dump(draw):
    n = draw(st.integers(1, 40))
    k = draw(st.integers(1, 6))
    d = {
        "N_spots": np.array(draw(st.lists(st.integers(1, 50), min_size=1, max_size=4)), dtype=np.int64),
        "scalar_int": draw(ints),
        "scalar_real": draw(st.floats(allow_nan=False, allow_infinity=False, width=64)),
        "empty": np.zeros(0, dtype=np.int64),
        "W_sparse": np.zeros((0, 0)),
        "E": np.array(draw(st.lists(floats, min_size=n * k, max_size=n * k))).reshape(n, k),
        "pairs": np.array(draw(st.lists(ints, min_size=2 * n, max_size=2 * n)), dtype=np.int64).reshape(n, 2),
        "counts": np.array(draw(st.lists(st.integers(0, 2**31 - 1), min_size=n, max_size=n)), dtype=np.int64),
        "beta_prior_mean": np.array(draw(st.lists(floats, min_size=k, max_size=k))),
        "beta_prior_std": np.array(draw(st.lists(floats, min_size=k, max_size=k))),
    }
    return d


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a, b)


@settings(max_examples=60, deadline=None, derandomize=True)
@given(dump(), st.sampled_from(["as_written", "spaced", "reordered", "Rish"]))
def test_round_trip(tmp_path_factory, d, style):
    p = str(tmp_path_factory.mktemp("rd") / "data_1.R")
    if style == "reordered":
        d = dict(reversed(list(d.items())))              # per-gene keys first: the native reader must still find them
    rdump.to_rdump(d, p)
    txt = open(p).read()
    if style == "spaced":
        txt = txt.replace(",", ", ").replace(".Dim=c(", ".Dim = c(").replace("<-\n", "<-  \n") + "\n\n"
    elif style == "Rish":
        txt = txt.replace("scalar_int <- %d" % d["scalar_int"], "scalar_int <- %dL" % d["scalar_int"])
    open(p, "w").write(txt)
    back = rdump.read_rdump(p)
    assert list(back) == list(d)
    for key in d:
        if np.asarray(d[key]).size == 0:
            assert np.asarray(back[key]).size == 0, key
        else:
            assert _same(back[key], d[key]), key
    assert isinstance(back["scalar_int"], int) and isinstance(back["scalar_real"], float)
    n, k = len(d["counts"]), len(d["beta_prior_mean"])
    nat = rdump.read_genes_native([p, p], n, k, n_threads=2)
    for key in rdump.GENE_KEYS:
        assert _same(nat[key][0], d[key]) and _same(nat[key][1], d[key]), key
    only = rdump.read_rdump_gene_only(p)
    assert sorted(only) == sorted(rdump.GENE_KEYS)
