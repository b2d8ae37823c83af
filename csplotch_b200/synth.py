"""Deterministic synthetic inputs of the shapes named in BASELINE.json.

Everything here produces *Stan data dictionaries* with exactly the R-dump keys
`generate_dictionary` writes (/root/reference/splotch/utils.py:684-774): 1-based
index arrays, `W_sparse` pairs with i<j, symmetric CSR `U`,`V`, eigenvalues of
D^-1/2 W D^-1/2 per tissue, sorted.  Counts are drawn from the model itself.

    c1  examples_csplotch (real files; fixture under tests/golden/c1, see
        tests/golden/make_golden.py)
    c2  ST-v1 geometry (fixture tests/golden/c2_geometry.npz when present, else a
        synthetic stand-in of the same size), splotch_stan_model, 2-level, Poisson+CAR
    c3  24 Visium arrays x 4992 spots, comp, 3-level (3,6,12), ZIP+CAR, K=10, C=12
    c4  4 Visium-HD arrays of 317x317 bins, splotch_stan_model_csr, 2-level (2,4), ZIP+CAR
    c5  = c3 covariates with 32768 genes
"""
from __future__ import annotations

import os

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_DATA = os.path.join(_HERE, "data")


# ----------------------------------------------------------------------------
# geometry
# ----------------------------------------------------------------------------
def hex_pairs(rows: int = 78, cols: int = 64) -> np.ndarray:
    """Adjacent pairs (i<j, 0-based, row-major spot index y*cols+x) of a full Visium array.

    Odd-r offset rule of the reference (splotch/utils_visium.py:98-106), equivalent to its
    true-hex distance <= 1.2 rule (utils_visium.py:214-217 -> utils.py:511-528)."""
    even = [(-1, -1), (0, -1), (-1, 0), (1, 0), (-1, 1), (0, 1)]
    odd = [(0, -1), (1, -1), (-1, 0), (1, 0), (0, 1), (1, 1)]
    out = []
    for y in range(rows):
        offs = even if y % 2 == 0 else odd
        for x in range(cols):
            i = y * cols + x
            for dx, dy in offs:
                xx, yy = x + dx, y + dy
                if 0 <= xx < cols and 0 <= yy < rows:
                    j = yy * cols + xx
                    if i < j:
                        out.append((i, j))
    return np.array(sorted(out), dtype=np.int64)


def grid4_pairs(h: int, w: int) -> np.ndarray:
    """4-neighbour pairs of an h x w integer grid, row-major (utils.py:530-552)."""
    idx = np.arange(h * w).reshape(h, w)
    a = np.stack([idx[:, :-1].ravel(), idx[:, 1:].ravel()], 1)
    b = np.stack([idx[:-1, :].ravel(), idx[1:, :].ravel()], 1)
    p = np.concatenate([a, b], 0)
    return p[np.lexsort((p[:, 1], p[:, 0]))]


def pairs_to_csr(pairs0: np.ndarray, n: int):
    """Symmetric CSR of the adjacency (utils.py:592-614): returns 1-based (V, U)."""
    r = np.concatenate([pairs0[:, 0], pairs0[:, 1]])
    c = np.concatenate([pairs0[:, 1], pairs0[:, 0]])
    m = sp.csr_matrix((np.ones(len(r), dtype=np.int64), (r, c)), shape=(n, n))
    m.sum_duplicates()
    m.sort_indices()
    return m.indices.astype(np.int64) + 1, m.indptr.astype(np.int64) + 1


def car_eigenvalues(pairs0: np.ndarray, n: int) -> np.ndarray:
    """eigvalsh(D^-1/2 W D^-1/2) of one tissue (utils.py:749-755)."""
    r = np.concatenate([pairs0[:, 0], pairs0[:, 1]])
    c = np.concatenate([pairs0[:, 1], pairs0[:, 0]])
    W = np.zeros((n, n))
    W[r, c] = 1.0
    d = W.sum(1)
    s = 1.0 / np.sqrt(d)
    return np.linalg.eigvalsh(W * s[:, None] * s[None, :])


def _cached_eig(name: str, fn):
    path = os.path.join(_DATA, name)
    if os.path.exists(path):
        return np.load(path)
    ev = fn()
    try:
        os.makedirs(_DATA, exist_ok=True)
        np.save(path, ev)
    except OSError:
        pass
    return ev


def grid_chunk_eigs(h: int, w: int, chunk: int = 100) -> np.ndarray:
    """The reference's chunked approximation for >10k-spot tissues (utils.py:629-670):
    eigenvalues of each (<=chunk x chunk) square block computed separately, concatenated, sorted."""
    evs = []
    for i0 in range(0, h - 1, chunk):       # range(0, xdim, wc) with xdim = max coordinate
        hh = min(chunk, h - i0)
        for j0 in range(0, w - 1, chunk):
            ww = min(chunk, w - j0)
            a, b = max(hh, ww), min(hh, ww)  # spectrum is symmetric under transpose
            evs.append(_cached_eig("eig_grid4_%dx%d.npy" % (a, b),
                                   lambda a=a, b=b: car_eigenvalues(grid4_pairs(a, b), a * b)))
    out = np.concatenate(evs)
    out.sort()
    assert out.size == h * w
    return out


def scale_and_round(mat: np.ndarray) -> np.ndarray:
    """Cell-type fractions as `generate_dictionary` stores them (/root/reference/splotch/utils.py:672-682, 729-731):
    8-decimal rounding with the last column closing the simplex, clipping at 0 and renormalising, 5-decimal rounding,
    and finally the largest entry of every spot absorbing what is missing from 1.  All rows at once."""
    frac = np.round(np.asarray(mat, dtype=np.float64), 8)
    frac[:, -1] = 1 - frac[:, :-1].sum(axis=1)
    frac = np.clip(frac, 0.0, None)
    frac = np.around(frac / frac.sum(axis=1, keepdims=True), decimals=5)
    rows = np.arange(frac.shape[0])
    top = frac.argmax(axis=1)
    rest = frac.copy()
    rest[rows, top] = 0.0
    frac[rows, top] = 1.0 - rest.sum(axis=1)
    assert frac.min() >= 0.0
    return frac


# ----------------------------------------------------------------------------
# shared covariates
# ----------------------------------------------------------------------------
def build_shared(tissue_sizes, tissue_pairs0, tissue_eigs, D, tissue_mapping, levels, level_2_mapping,
                 level_3_mapping, size_factors, n_covariates, zi, car, E=None, nb=0) -> dict:
    """Assemble the gene-independent part of the data dictionary (generate_dictionary)."""
    n_levels = len(levels)
    N = int(np.sum(tissue_sizes))
    d = {
        "N_spots": np.asarray(tissue_sizes, dtype=np.int64),
        "N_tissues": len(tissue_sizes),
        "N_covariates": int(n_covariates),
        "tissue_mapping": np.asarray(tissue_mapping, dtype=np.int64),
        "N_levels": n_levels,
        "zi": int(zi), "nb": int(nb), "car": int(car),
        "N_level_1": int(levels[0]),
        "N_level_2": int(levels[1]) if n_levels > 1 else 0,
        "N_level_3": int(levels[2]) if n_levels > 2 else 0,
        "level_2_mapping": np.asarray(level_2_mapping if n_levels > 1 else [], dtype=np.int64),
        "level_3_mapping": np.asarray(level_3_mapping if n_levels > 2 else [], dtype=np.int64),
        "size_factors": np.asarray(size_factors, dtype=np.float64),
        "D": np.asarray(D, dtype=np.int64),
    }
    if E is not None:
        d["E"] = np.asarray(E, dtype=np.float64)
        d["N_celltypes"] = int(d["E"].shape[1])
    if car:
        offs = np.concatenate([[0], np.cumsum(tissue_sizes)])
        allp = np.concatenate([p + offs[t] for t, p in enumerate(tissue_pairs0)], 0)
        allp = allp[np.lexsort((allp[:, 1], allp[:, 0]))]
        d["W_n"] = np.array([len(allp)], dtype=np.int64)
        d["W_sparse"] = allp + 1
        deg = np.bincount(np.concatenate([allp[:, 0], allp[:, 1]]), minlength=N)
        assert deg.min() > 0, "CAR prior needs every spot to have a neighbour"
        d["D_sparse"] = deg.astype(np.int64)
        d["eig_values"] = np.sort(np.concatenate(tissue_eigs))
        V, U = pairs_to_csr(allp, N)
        d["U"], d["V"] = U, V
    else:
        d["W_n"] = np.zeros(0, dtype=np.int64)
        d["W_sparse"] = np.zeros((0, 0))
        d["D_sparse"] = np.zeros(0)
        d["eig_values"] = np.zeros(0)
        d["U"] = np.zeros(0, dtype=np.int64)
        d["V"] = np.zeros(0, dtype=np.int64)
    return d


def shared_c3(n_arrays: int = 24, rows: int = 78, cols: int = 64, K: int = 10, C: int = 12,
              seed: int = 20243, levels=(3, 6, 12)) -> dict:
    """c3/c5 covariates: full Visium arrays, 3-level hierarchy, cell-type compositions."""
    rng = np.random.default_rng(seed)
    n = rows * cols
    pairs = hex_pairs(rows, cols)
    if (rows, cols) == (78, 64):
        ev = _cached_eig("eig_hex78x64.npy", lambda: car_eigenvalues(pairs, n))
    else:
        ev = car_eigenvalues(pairs, n)
    sizes = [n] * n_arrays
    L1, L2, L3 = levels
    # level-3 units own consecutive arrays; level mappings nest evenly
    tissue_mapping = (np.arange(n_arrays) * L3 // n_arrays) + 1
    level_3_mapping = (np.arange(L3) * L2 // L3) + 1
    level_2_mapping = (np.arange(L2) * L1 // L2) + 1
    band = (np.arange(rows) * C // rows)  # AARs as contiguous row bands
    D = np.tile(np.repeat(band, cols) + 1, n_arrays)
    N = n * n_arrays
    E = scale_and_round(rng.dirichlet(np.full(K, 0.5), size=N))
    sf = np.clip(rng.lognormal(0.0, 0.5, size=N), 0.05, 8.0)
    return build_shared(sizes, [pairs] * n_arrays, [ev] * n_arrays, D, tissue_mapping, levels,
                        level_2_mapping, level_3_mapping, sf, C, zi=1, car=1, E=E)


def shared_c4(n_arrays: int = 4, side: int = 317, C: int = 12, seed: int = 20244, levels=(2, 4)) -> dict:
    """c4 covariates: Visium-HD bins on an integer grid, 4-neighbour CAR, chunked eigenvalues."""
    rng = np.random.default_rng(seed)
    n = side * side
    pairs = grid4_pairs(side, side)
    ev = grid_chunk_eigs(side, side, 100) if n >= 10000 else car_eigenvalues(pairs, n)
    L1, L2 = levels
    tissue_mapping = (np.arange(n_arrays) * L2 // n_arrays) + 1
    level_2_mapping = (np.arange(L2) * L1 // L2) + 1
    band = (np.arange(side) * C // side)
    D = np.tile(np.repeat(band, side) + 1, n_arrays)
    N = n * n_arrays
    sf = np.clip(rng.lognormal(0.0, 0.7, size=N), 0.02, 12.0)
    return build_shared([n] * n_arrays, [pairs] * n_arrays, [ev] * n_arrays, D, tissue_mapping, levels,
                        level_2_mapping, [], sf, C, zi=1, car=1)


def shared_c2(seed: int = 20242, geometry_npz: str | None = None) -> dict:
    """c2 covariates: ST-v1 example geometry, 2-level (2,4), Poisson + CAR, C=11.

    With the committed fixture (made from /root/reference/examples/Annotations by
    tests/golden/make_golden.py through the reference's own tissue detection) the
    geometry is the real one; otherwise 10 synthetic irregular ST arrays of the same size."""
    rng = np.random.default_rng(seed)
    if geometry_npz is None:
        cand = os.path.join(_HERE, "..", "tests", "golden", "c2_geometry.npz")
        geometry_npz = cand if os.path.exists(cand) else None
    if geometry_npz is not None:
        z = np.load(geometry_npz)
        sizes = z["N_spots"].tolist()
        offs = np.concatenate([[0], np.cumsum(sizes)])
        W0 = z["W_sparse"] - 1
        pairs, eigs = [], []
        for t in range(len(sizes)):
            m = (W0[:, 0] >= offs[t]) & (W0[:, 0] < offs[t + 1])
            pairs.append(W0[m] - offs[t])
        eig_all = z["eig_values"]
        D = z["D"]
        C = int(z["N_covariates"])
        tissue_mapping = z["tissue_mapping"]
        level_2_mapping = z["level_2_mapping"]
        levels = (int(z["N_level_1"]), int(z["N_level_2"]))
        N = int(np.sum(sizes))
        sf = np.clip(rng.lognormal(0.0, 0.5, size=N), 0.05, 8.0)
        d = build_shared(sizes, pairs, [eig_all], D, tissue_mapping, levels, level_2_mapping, [], sf, C,
                         zi=0, car=1)
        return d
    # synthetic stand-in: 10 arrays cut from a 33x35 ST grid, ~247 spots each
    sizes, pairs, eigs, Ds = [], [], [], []
    C = 11
    for t in range(10):
        h, w = 13 + (t % 3), 19 - (t % 3)
        n = h * w
        p = grid4_pairs(h, w)
        sizes.append(n)
        pairs.append(p)
        eigs.append(car_eigenvalues(p, n))
        Ds.append(np.repeat(np.arange(h) * C // h, w) + 1)
    tissue_mapping = (np.arange(10) * 4 // 10) + 1
    N = int(np.sum(sizes))
    sf = np.clip(rng.lognormal(0.0, 0.5, size=N), 0.05, 8.0)
    return build_shared(sizes, pairs, eigs, np.concatenate(Ds), tissue_mapping, (2, 4), [1, 1, 2, 2], [],
                        sf, C, zi=0, car=1)


def shared_tiny(seed: int = 0, family: str = "comp_splotch_stan_model", n_levels: int = 3, zi: int = 1,
                car: int = 1, K: int = 3, C: int = 4, tissues=((3, 4), (2, 5), (4, 3), (3, 3))) -> dict:
    """Random tiny design for unit tests (every (zi, car, levels, K) combination)."""
    rng = np.random.default_rng(seed)
    sizes, pairs, eigs = [], [], []
    for (h, w) in tissues:
        n = h * w
        p = grid4_pairs(h, w)
        # shuffle spot order so adjacency is not banded
        perm = rng.permutation(n)
        p = perm[p]
        p = np.sort(p, axis=1)
        p = p[np.lexsort((p[:, 1], p[:, 0]))]
        sizes.append(n); pairs.append(p); eigs.append(car_eigenvalues(p, n))
    T = len(sizes)
    N = int(np.sum(sizes))
    if n_levels == 1:
        levels, l2, l3 = (2,), [], []
        tm = rng.integers(1, 3, size=T)
    elif n_levels == 2:
        levels, l2, l3 = (2, 3), [1, 2, 2], []
        tm = rng.integers(1, 4, size=T)
    else:
        levels, l2, l3 = (2, 3, 4), [1, 2, 2], [1, 2, 3, 3]
        tm = rng.integers(1, 5, size=T)
    D = rng.integers(1, C + 1, size=N)
    sf = np.clip(rng.lognormal(0.0, 0.5, size=N), 0.05, 8.0)
    E = scale_and_round(rng.dirichlet(np.full(K, 0.5), size=N)) if family == "comp_splotch_stan_model" else None
    return build_shared(sizes, pairs, eigs, D, tm, levels, l2, l3, sf, C, zi=zi, car=car, E=E)


# ----------------------------------------------------------------------------
# counts drawn from the model itself
# ----------------------------------------------------------------------------
def _bottom_rows(shared: dict) -> np.ndarray:
    """0-based row (within the stacked level table) of every spot's bottom-level unit."""
    L1, L2, L3 = shared["N_level_1"], shared["N_level_2"], shared["N_level_3"]
    r0 = L1 + L2 if L3 else (L1 if L2 else 0)
    tm = np.asarray(shared["tissue_mapping"]) - 1
    return np.repeat(r0 + tm, np.asarray(shared["N_spots"]))


def simulate_counts(shared: dict, n_genes: int, seed: int, comp: bool | None = None, chunk: int = 16):
    """Draw counts[G,N] (int32) and (for comp) beta_prior_mean/std[G,K] from the generative model."""
    rng = np.random.default_rng(seed)
    N = int(np.sum(shared["N_spots"]))
    C = int(shared["N_covariates"])
    comp = ("E" in shared) if comp is None else comp
    K = int(shared["N_celltypes"]) if comp else 1
    L1, L2, L3 = shared["N_level_1"], shared["N_level_2"], shared["N_level_3"]
    car, zi = shared["car"], shared["zi"]
    rows = _bottom_rows(shared)
    Dm1 = np.asarray(shared["D"]) - 1
    lsf = np.log(np.asarray(shared["size_factors"]))
    E = np.asarray(shared["E"]) if comp else None
    if car:
        W0 = np.asarray(shared["W_sparse"]) - 1
        Wm = sp.csr_matrix((np.ones(len(W0)), (W0[:, 0], W0[:, 1])), shape=(N, N))
        Wm = Wm + Wm.T
        deg = np.asarray(shared["D_sparse"], dtype=np.float64)
    counts = np.empty((n_genes, N), dtype=np.int32)
    for g0 in range(0, n_genes, chunk):
        g1 = min(n_genes, g0 + chunk)
        G = g1 - g0
        raw = rng.standard_normal((G, L1 + L2 + L3, C, K))
        s2 = np.abs(rng.normal(0, 0.5, G)); s3 = np.abs(rng.normal(0, 0.5, G))
        sigma = np.abs(rng.normal(0, 1.0, G))
        mu_g = rng.normal(-1.5, 1.5, G)
        b = np.empty_like(raw)
        b[:, :L1] = 2.0 * raw[:, :L1] + mu_g[:, None, None, None]
        if L2:
            b[:, L1:L1 + L2] = b[:, np.asarray(shared["level_2_mapping"]) - 1] + s2[:, None, None, None] * raw[:, L1:L1 + L2]
        if L3:
            b[:, L1 + L2:] = b[:, L1 + np.asarray(shared["level_3_mapping"]) - 1] + s3[:, None, None, None] * raw[:, L1 + L2:]
        bj = b[:, rows, Dm1, :]                      # (G, N, K)
        ll = np.einsum("gnk,nk->gn", bj, E) if comp else bj[:, :, 0]
        if car:
            tau = np.clip(1.0 / rng.gamma(1.0, 1.0, G), 0.2, 20.0)
            a = rng.uniform(0.2, 0.95, G)
            psi = np.zeros((N, G))
            for _ in range(50):
                psi = a[None, :] * (Wm @ psi) / deg[:, None] + rng.standard_normal((N, G)) / np.sqrt(tau[None, :] * deg[:, None])
            ll = ll + psi.T
        ll = ll + 0.3 * sigma[:, None] * rng.standard_normal((G, N))
        lam = np.exp(np.clip(ll + lsf[None, :], -30, 12))
        y = rng.poisson(lam)
        if zi:
            theta = rng.beta(1.0, 2.0, G)
            y = np.where(rng.random((G, N)) < theta[:, None], 0, y)
        counts[g0:g1] = y.astype(np.int32)
    out = {"counts": counts}
    if comp:
        out["beta_prior_mean"] = np.zeros((n_genes, K))      # utils_sc.py:110-112 defaults
        out["beta_prior_std"] = np.full((n_genes, K), 2.0)
    return out


def gene_dict(shared: dict, genes: dict, g: int) -> dict:
    """The complete Stan data dictionary of gene g (what data_<g+1>.R holds)."""
    d = dict(shared)
    d["counts"] = genes["counts"][g]
    if "beta_prior_mean" in genes:
        d["beta_prior_mean"] = genes["beta_prior_mean"][g]
        d["beta_prior_std"] = genes["beta_prior_std"][g]
    return d


CONFIGS = {
    "c2": dict(family="splotch_stan_model", shared=shared_c2, seed=20242, genes=11313),
    "c3": dict(family="comp_splotch_stan_model", shared=shared_c3, seed=20243, genes=20000),
    "c4": dict(family="splotch_stan_model_csr", shared=shared_c4, seed=20244, genes=2000),
    "c5": dict(family="comp_splotch_stan_model", shared=shared_c3, seed=20245, genes=32768),
}


def make_config(name: str, n_genes: int | None = None):
    """(family, shared, genes) of a named configuration; n_genes caps the gene count."""
    cfg = CONFIGS[name]
    shared = cfg["shared"]()
    G = cfg["genes"] if n_genes is None else int(n_genes)
    genes = simulate_counts(shared, G, cfg["seed"])
    return cfg["family"], shared, genes
