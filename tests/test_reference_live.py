"""Live checks against the reference's own Python, imported read-only from /root/reference in the build container
(skipped where it is absent, e.g. on the GPU box; the same import recipe as tests/golden/make_golden.py):
a `combined_<g>.csv` written by this repo's native writer is parsed by the reference's `read_stan_csv`
(/root/reference/splotch/utils.py:139-232, both its pandas and its line-by-line branch) and summarised by its
`summarize_stan_hdf5` (utils.py:234-257, h5py replaced by a recording stand-in: h5py is not installed here) to the same
arrays, groups and shapes as `stancsv.read_stan_csv` / `stancsv.summarize` give."""
import os
import sys

import numpy as np
import pytest

from csplotch_b200 import cli, stancsv

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "splotch", "utils.py")),
                                reason="/root/reference not on this box")

from tests.test_vinit_host import host_only, _c1_data_dir  # noqa: E402,F401  (fixture: oracle-backed device classes)


@pytest.fixture(scope="module")
def ref_utils():
    before, path0 = dict(sys.modules), list(sys.path)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    U, _ = make_golden.load_reference()
    yield U
    # the stand-ins for the reference's absent imports (h5py, jax, ...) must not leak into other tests
    for name in list(sys.modules):
        if name not in before:
            del sys.modules[name]
    sys.path[:] = path0


class _H5:
    """Records what summarize_stan_hdf5 writes: {dataset name: array}."""
    files = {}

    def __init__(self, path, mode="r"):
        self.path = path
        _H5.files[path] = {}

    def create_dataset(self, name, data=None):
        _H5.files[self.path][name] = np.array(data)

    def close(self):
        pass


def test_reference_reader_and_summariser_on_our_csv(ref_utils, host_only, golden_dir, tmp_path, monkeypatch):
    data_dir = _c1_data_dir(golden_dir, tmp_path, (5,))
    out = str(tmp_path / "out")
    assert cli.main(["-g", "5", "-d", data_dir, "-o", out, "-b", "comp_splotch_stan_model", "-n", "20", "-c", "2",
                     "--seed", "3"]) == 0
    p = os.path.join(out, "0", "combined_5.csv")
    variables = list(stancsv.SUMMARY_VARS)
    mine = stancsv.read_stan_csv(p, variables)
    for efficient in (False, True):
        theirs = ref_utils.read_stan_csv(p, variables, efficient=efficient)
        assert sorted(theirs) == sorted(mine)
        for k in mine:
            assert theirs[k].shape == mine[k].shape, (k, efficient)
            if efficient:       # float() per token, like this repo's reader: exact
                assert np.array_equal(theirs[k], mine[k]), k
            else:               # pandas' default C float parser is not correctly rounded on 17-digit text (up to 1e-12 relative)
                assert np.allclose(theirs[k], mine[k], rtol=1e-11, atol=1e-300), k
    assert mine["beta_level_1"].shape == (40, 1, 14, 5) and mine["sigma"].shape == (40, 1)
    # the summariser (what `splotch -s` / splotch_summarize_output leave for splotch_compile_lambdas / _betas)
    monkeypatch.setattr(ref_utils, "h5py", type("h5py", (), {"File": _H5}))
    h5 = str(tmp_path / "combined_5.hdf5")
    ref_utils.summarize_stan_hdf5(p, h5)
    theirs = _H5.files[h5]
    want = stancsv.summarize(mine)
    flat = {"%s/%s" % (k, s): a for k, d in want.items() for s, a in d.items()}
    assert sorted(theirs) == sorted(flat)
    for k in flat:
        assert theirs[k].shape == flat[k].shape, k
        assert np.allclose(theirs[k], flat[k], rtol=1e-13, atol=0), k
    assert theirs["sigma/mean"].shape == (1,) and theirs["lambda/mean"].shape == (10,)
    # and the file this repo writes for `-s` holds exactly those datasets
    monkeypatch.delitem(sys.modules, "h5py", raising=False)          # the stand-in above is not an HDF5 library
    path = stancsv.write_summary(str(tmp_path / "mine.hdf5"), want)
    back = stancsv.read_summary(path)
    assert sorted("%s/%s" % (k, s) for k, d in back.items() for s in d) == sorted(theirs)


def test_reference_rdump_reader_on_our_writer(ref_utils, golden_dir, tmp_path):
    """to_rdump of this repo -> read_rdump of the reference (utils.py:78-111), on a reference-generated gene."""
    from csplotch_b200 import rdump
    d = rdump.read_rdump(os.path.join(golden_dir, "c1", "data_3.R"))
    p = str(tmp_path / "data_3.R")
    rdump.to_rdump(d, p)
    theirs = ref_utils.read_rdump(p)
    assert sorted(theirs) == sorted(d)
    for k in d:
        assert np.array_equal(np.asarray(theirs[k]), np.asarray(d[k])), k
    assert open(p).read() == open(os.path.join(golden_dir, "c1", "data_3.R")).read()      # byte-identical files


def test_composition_rounding_matches_the_reference(ref_utils):
    """synth.scale_and_round (all rows at once) against the reference's row loop (utils.py:672-682 after :729-730)."""
    from csplotch_b200 import synth
    rng = np.random.default_rng(0)
    for K in (1, 2, 5, 8, 10, 16):
        raw = rng.dirichlet(np.full(K, 0.3), size=500)
        raw[::50] = np.eye(K)[rng.integers(0, K, 10)]                   # pure spots
        want = np.round(raw, 8)
        want[:, -1] = 1 - want[:, :-1].sum(axis=1)
        want = ref_utils.scale_and_round(want)
        got = synth.scale_and_round(raw)
        assert np.abs(got - want).max() <= 5e-16 and np.abs(got.sum(axis=1) - 1).max() < 1e-15
        if K < 8:                      # numpy sums 8 or more terms in a different order along an axis than over a row
            assert np.array_equal(got, want)


def test_car_helpers_match_the_reference(ref_utils, tmp_path, monkeypatch):
    """The inputs of the synthetic configs are built the way the reference builds them: 4-neighbour adjacency
    (utils.py:530-552), symmetric CSR `U`, `V` (utils.py:592-614) and the chunked eigenvalues of Visium-HD sized
    tissues (utils.py:629-670)."""
    from csplotch_b200 import synth
    monkeypatch.setattr(synth, "_DATA", str(tmp_path))                           # keep the eigenvalue cache out of the tree
    for h, w in ((7, 5), (12, 23)):
        coords = np.array([(i, j) for i in range(h) for j in range(w)])          # row-major, as synth numbers the bins
        W = ref_utils.get_spot_adjacency_matrix_4n(coords).toarray()
        pairs = synth.grid4_pairs(h, w)
        mine = np.zeros_like(W)
        mine[pairs[:, 0], pairs[:, 1]] = 1
        mine[pairs[:, 1], pairs[:, 0]] = 1
        assert np.array_equal(mine, W)
        V, U = synth.pairs_to_csr(pairs, h * w)
        Vr, Ur = ref_utils.sparse_to_csr_indptr(pairs + 1, h * w)
        assert np.array_equal(V, Vr) and np.array_equal(U, Ur)
    # chunk side 10 <-> chunksize 100 spots; sizes whose largest coordinate is not a multiple of the chunk side (the
    # reference's loop would drop the last row / column there and trip its own assertion)
    for h, w in ((23, 17), (32, 26)):
        coords_str = np.array(["%d_%d" % (i, j) for i in range(h) for j in range(w)])
        want = ref_utils.eigs_square_chunking(coords_str, chunksize=100)
        got = synth.grid_chunk_eigs(h, w, chunk=10)
        assert got.shape == want.shape and np.abs(got - want).max() < 1e-12


def test_visium_hex_adjacency_matches_the_reference(ref_utils):
    """synth.hex_pairs (odd-r offsets) against the reference's rule for Visium arrays: true-hex coordinates and
    distance <= 1.2 (utils_visium.py:80-87, 214-217 -> utils.py:511-528), on a 12 x 9 array in pseudo-hex indexing."""
    import importlib
    from csplotch_b200 import synth
    ref_utils_visium = importlib.import_module("splotch.utils_visium")
    rows, cols = 12, 9
    pseudo = [(2 * x + (y % 2), y) for y in range(rows) for x in range(cols)]      # Visium's doubled-x, odd rows offset
    W = np.asarray(ref_utils_visium.get_spot_adjacency_matrix_hex(pseudo))
    pairs = synth.hex_pairs(rows, cols)
    mine = np.zeros_like(W)
    mine[pairs[:, 0], pairs[:, 1]] = 1
    mine[pairs[:, 1], pairs[:, 0]] = 1
    assert np.array_equal(mine, W) and W.sum(axis=1).max() == 6


def test_build_shared_equals_generate_dictionary(ref_utils):
    """The gene-shared dictionary of the synthetic configs (synth.build_shared) key by key against the reference's
    generate_dictionary (utils.py:684-774) on the same small Visium-like design: compositional, 3 levels, CAR + ZI."""
    from scipy.sparse import csr_matrix
    from csplotch_b200 import synth
    rng = np.random.default_rng(3)
    shapes = [(5, 4), (3, 6), (4, 4)]
    C, K = 4, 3
    levels, l2, l3 = (2, 3, 3), [1, 2, 2], [1, 2, 3]
    tm = [1, 3, 2]
    sizes, pairs, eigs, W_list, Wn, coords, sf, aar, comp = [], [], [], [], [], [], [], [], []
    for (h, w) in shapes:
        n = h * w
        p = synth.hex_pairs(h, w)
        W = np.zeros((n, n))
        W[p[:, 0], p[:, 1]] = W[p[:, 1], p[:, 0]] = 1
        sizes.append(n); pairs.append(p); eigs.append(synth.car_eigenvalues(p, n))
        W_list.append(csr_matrix(W)); Wn.append(len(p))
        coords.append(np.array(["%d_%d" % (x, y) for y in range(h) for x in range(w)]))
        sf.append(rng.uniform(0.2, 3.0, n))
        d = rng.integers(0, C, n)
        aar.append(np.eye(C, dtype=int)[d].T)                               # (C, n) one-hot, as the annotation files
        comp.append(rng.dirichlet(np.full(K, 0.5), n).T)                    # (K, n)
    D = np.concatenate([a.argmax(0) + 1 for a in aar])
    E = np.vstack([c.T for c in comp])
    mine = synth.build_shared(sizes, pairs, eigs, D, tm, levels, l2, l3, np.concatenate(sf), C, zi=1, car=1,
                              E=synth.scale_and_round(E))
    ref = ref_utils.generate_dictionary(sizes, len(sizes), C, list(levels), coords, sf, aar, [l2, l3], tm, W_list, Wn,
                                        True, True, False, True, comp)
    assert set(ref) <= set(mine) | {"N_celltypes"}
    for key in ref:
        a, b = np.asarray(ref[key], dtype=np.float64), np.asarray(mine[key], dtype=np.float64)
        assert a.shape == b.shape or a.size == b.size == 0, key
        if a.size:
            assert np.allclose(a, b, rtol=0, atol=1e-12 if key in ("eig_values", "E") else 0), key
