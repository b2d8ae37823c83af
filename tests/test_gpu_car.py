"""the CAR pre-compute on the GPU (csplotch_car_precompute) against the reference's own values.

  * c2 geometry: tests/golden/c2_geometry.npz was produced by the reference's generate_dictionary
    (/root/reference/splotch/utils.py:735-765, tests/golden/make_golden.py): its eig_values (per-tissue
    numpy.linalg.eigvalsh), D_sparse, U, V must come back — eigenvalues to 1e-10, the integer arrays exactly;
  * identical tissues are solved once (c3: 24 identical Visium arrays -> one dense solve);
  * tissues beyond the exact limit: the chunked approximation of eigs_square_chunking (utils.py:629-670), here against
    csplotch_b200.synth.grid_chunk_eigs, which tests/test_reference_live.py pins to the reference's function; and the
    exact option (`exact_max_n`) against numpy.linalg.eigvalsh of the whole tissue."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from csplotch_b200 import car, synth


def test_c2_geometry_equals_reference_values(golden_dir):
    z = np.load(os.path.join(golden_dir, "c2_geometry.npz"))
    out = car.precompute(z["N_spots"], z["W_sparse"])
    assert np.max(np.abs(out["eig_values"] - z["eig_values"])) < 1e-10
    assert np.array_equal(out["D_sparse"], z["D_sparse"].astype(np.int64))
    V, U = synth.pairs_to_csr(z["W_sparse"] - 1, int(np.sum(z["N_spots"])))
    assert np.array_equal(out["U"], U) and np.array_equal(out["V"], V)
    if "U" in z.files:
        assert np.array_equal(out["U"], z["U"]) and np.array_equal(out["V"], z["V"])
    assert out["n_solves"] <= len(z["N_spots"])


def test_identical_tissues_are_solved_once():
    rows, cols, n_arrays = 78, 64, 3
    pairs = synth.hex_pairs(rows, cols)
    n = rows * cols
    W = np.concatenate([pairs + t * n for t in range(n_arrays)]) + 1
    out = car.precompute([n] * n_arrays, W)
    assert out["n_solves"] == 1
    ev = np.sort(np.tile(synth.shared_c3(n_arrays=1)["eig_values"], n_arrays))   # the cached numpy eigvalsh values
    assert np.max(np.abs(out["eig_values"] - ev)) < 1e-10
    assert out["eig_values"].min() >= -1 - 1e-12 and out["eig_values"].max() <= 1 + 1e-12


def test_chunked_and_exact_large_tissue():
    h, w = 46, 50                     # one 2 300-spot grid tissue; the limits are scaled down so that it counts as "large"
    pairs = synth.grid4_pairs(h, w)
    n = h * w
    yy, xx = np.divmod(np.arange(n), w)
    coords = np.stack([yy, xx], 1)
    # chunked: squares of 20 x 20 coordinates (chunk_spots = 400), as eigs_square_chunking does with 100 x 100
    out = car.precompute([n], pairs + 1, coords=coords, chunk_spots=400, exact_max_n=1000)
    want = synth.grid_chunk_eigs(h, w, chunk=20)
    assert np.max(np.abs(out["eig_values"] - want)) < 1e-10
    assert out["n_solves"] <= 6       # 20x20, 20x10, 6x20, 6x10 (+ transposes hash apart): repeated chunks solved once
    # "x_y" coordinate strings like the reference's labels
    out_s = car.precompute([n], pairs + 1, coords=["%d_%d" % (a, b) for a, b in coords], chunk_spots=400, exact_max_n=1000)
    assert np.array_equal(out_s["eig_values"], out["eig_values"])
    # exact: the whole tissue in one dense solve
    exact = car.precompute([n], pairs + 1, exact_max_n=5000)
    assert exact["n_solves"] == 1
    assert np.max(np.abs(exact["eig_values"] - synth.car_eigenvalues(pairs, n))) < 1e-10
    # the approximation differs from the exact spectrum (that is why the option exists)
    assert np.max(np.abs(exact["eig_values"] - out["eig_values"])) > 1e-3


def test_model_without_eig_values_gets_them_from_the_gpu():
    import oracle
    from csplotch_b200 import api
    shared = synth.shared_tiny(seed=3, family="splotch_stan_model", n_levels=2, zi=1, car=1)
    genes = synth.simulate_counts(shared, 1, 4)
    bare = {k: v for k, v in shared.items() if k not in ("eig_values", "U", "V", "D_sparse")}
    model = api.Model(bare, "splotch_stan_model")
    b = api.Batch(model, genes["counts"])
    th = np.random.default_rng(0).normal(0, 1, (1, model.dim))
    lp, grad = b.log_prob_grad(th, gene_of=[0])
    lp0, g0 = oracle.Oracle(synth.gene_dict(shared, genes, 0), "splotch_stan_model").log_prob_grad(th[0])
    assert abs(lp[0] - lp0) < 1e-10 * max(1.0, abs(lp0))
    assert np.max(np.abs(grad[0] - g0) / np.maximum(1.0, np.abs(g0))) < 1e-10


def test_bad_inputs_are_refused():
    from csplotch_b200._lib import CsplotchError
    with pytest.raises(CsplotchError):
        car.precompute([3, 3], np.array([[1, 4]]))          # a pair across two tissues
    with pytest.raises(CsplotchError):
        car.precompute([3], np.array([[1, 2]]))             # spot 3 has no neighbour
    with pytest.raises(CsplotchError):
        car.precompute([2300], synth.grid4_pairs(46, 50) + 1, exact_max_n=100)   # large tissue without coordinates
