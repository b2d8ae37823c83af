#!/usr/bin/env python
"""Generate the committed fixtures under tests/golden/ by IMPORTING THE REFERENCE (read-only, in the
build container only; /root/reference does not exist on the GPU box).

Absent third-party modules that the reference imports at module level but that the code paths used
here never call (h5py, jax, skimage, matplotlib, scanpy, sklearn.cluster) are stubbed; everything
that runs is the reference's own code:

  c1/        bin/splotch_generate_input_files on examples_csplotch (ST-v1 mode, -d 0, -l 1, -p):
             data_<g>.R for the 7 genes, byte for byte what the reference writes, + the column labels
  c2_geometry.npz   generate_dictionary() on the 10 ST-v1 arrays of examples/Annotations (real
             coordinates + AARs; count tables are missing upstream, so sequencing depth is stubbed)
  io/        to_rdump / read_rdump / read_stan_csv round-trip vectors for the I/O tests
  snap25_head.npz   first tissues of Science/human/data_SNAP25.R parsed with the reference's read_rdump
  snap25_full.npz   the whole file (`python tests/golden/make_golden.py snap25_full`)
"""
import importlib.machinery
import importlib.util
import os
import pickle
import shutil
import sys
import tempfile
import types
from unittest import mock

import numpy as np
import pandas as pd

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def stub_modules():
    import scipy.ndimage
    for name in ["h5py", "jax", "jax.numpy", "jax.example_libraries", "jax.example_libraries.optimizers",
                 "jax.experimental", "skimage", "skimage.segmentation", "skimage.morphology", "skimage.feature",
                 "matplotlib", "matplotlib.pyplot", "scanpy", "scanpy._utils", "anndata"]:
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)
    # `import jax.numpy as np` inside splotch.utils: sparse_to_csr_indptr only needs np.int64,
    # so plain numpy stands in for jax.numpy
    sys.modules["jax.numpy"] = np
    sys.modules["jax"].numpy = np
    for old, new in (("scipy.ndimage.measurements", scipy.ndimage), ("scipy.ndimage.morphology", scipy.ndimage)):
        try:
            __import__(old)
        except Exception:
            sys.modules[old] = new


def load_reference():
    stub_modules()
    sys.path.insert(0, REF)
    import splotch.utils as U
    loader = importlib.machinery.SourceFileLoader("ref_generate", os.path.join(REF, "bin", "splotch_generate_input_files"))
    spec = importlib.util.spec_from_loader("ref_generate", loader)
    gen = importlib.util.module_from_spec(spec)
    loader.exec_module(gen)
    return U, gen


def make_c1(U, gen):
    src = os.path.join(REF, "examples_csplotch")
    tmp = tempfile.mkdtemp()
    cwd = os.getcwd()
    try:
        for f in os.listdir(src):
            shutil.copy(os.path.join(src, f), tmp)
        os.chdir(tmp)
        count_files = ["10005CN36_C1_stdata_under_tissue_IDs.txt.unified.tsv"]
        # celltype_beta_priors without single-cell data: the reference's defaults (utils_sc.py:110-112)
        def priors(celltype_names, genes, **kw):
            return (pd.DataFrame(np.zeros((len(genes), len(celltype_names))), index=genes, columns=celltype_names),
                    pd.DataFrame(2.0 * np.ones((len(genes), len(celltype_names))), index=genes, columns=celltype_names))
        gen.celltype_beta_priors = priors
        scaling = 20.0
        gen.generate_splotch_inputs(count_files, "Metadata_cSplotch.tsv", 2000, scaling, 1, 0.0, "data_c1",
                                    True, True, False, Visium=False, compositional=True)
        dst = os.path.join(OUT, "c1")
        os.makedirs(dst, exist_ok=True)
        for g in range(1, 8):
            shutil.copy(os.path.join("data_c1", "0", "data_%d.R" % g), os.path.join(dst, "data_%d.R" % g))
        info = pickle.load(open(os.path.join("data_c1", "information.p"), "rb"))
        with open(os.path.join(dst, "information.txt"), "w") as f:
            f.write("genes: %s\n" % list(info["genes"]))
            f.write("spots: %s\n" % [c for _, c in info["filenames_and_coordinates"]])
            f.write("annotation_mapping: %s\n" % list(info["annotation_mapping"]))
            f.write("celltype_mapping: %s\n" % list(info["celltype_mapping"]))
            f.write("beta_mapping: %s\n" % info["beta_mapping"])
            f.write("scaling_factor: %s\n" % info["scaling_factor"])
        d = U.read_rdump(os.path.join(dst, "data_1.R"))
        print("c1: N=%d W_n=%s K=%s C=%d" % (np.sum(d["N_spots"]), d["W_n"], d["N_celltypes"], d["N_covariates"]))
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp)


def make_c2_geometry(U, gen):
    """ST-v1 example: annotations are present, count tables are not.  Run the reference's own
    filtering / tissue detection / adjacency / eigenvalue code on the annotated coordinates."""
    meta = pd.read_csv(os.path.join(REF, "examples", "metadata.tsv"), sep="\t")
    n_levels = 2
    levels = {key: {} for key in list(meta["Level 1"].unique())}
    for l1 in levels:
        levels[l1] = list(meta[meta["Level 1"] == l1]["Level 2"].unique())
    count_files = list(meta["Count file"])
    level_mappings, last_ids, cond = U.get_variable_mappings(count_files, meta, levels, n_levels)
    N_levels = [0] * n_levels
    U.n_elements_per_level(levels, 0, N_levels)
    aar_list, tm_list, coords_list, nspots, W_list, Wn_list, sf_list = [], [], [], [], [], [], []
    aar_names = None
    for _, row in meta.iterrows():
        annot = os.path.join(REF, "examples", row["Annotation file"])
        aar_matrix, names = U.read_aar_matrix(annot)
        aar_names = names if aar_names is None else aar_names
        assert np.array_equal(names, aar_names)
        coords_str = np.array(list(aar_matrix.columns))
        good = np.array([aar_matrix[c].sum() > 0 for c in coords_str])
        coords_str = coords_str[good]
        coords = np.array([list(map(float, c.split("_"))) for c in coords_str])
        labels, labeled = U.detect_tissue_sections(coords, True, 2000)
        tm = last_ids[str([row["Level 1"], row["Level 2"]])]
        for t in labels:
            sel = U.get_tissue_section_spots(t, coords, labeled)
            cs, cf = coords_str[sel], coords[sel]
            W = U.get_spot_adjacency_matrix(cf)
            conn = np.array(W.sum(0) > 0).squeeze()
            cs, cf, W = cs[conn], cf[conn], W[conn, :][:, conn]
            aar_list.append(aar_matrix[cs].values)
            tm_list.append(tm)
            coords_list.append(cs)
            nspots.append(len(cs))
            from scipy.sparse import csr_matrix
            W_list.append(csr_matrix(W))
            Wn_list.append(int(W.sum() / 2))
            sf_list.append(np.ones(len(cs)))
    data = U.generate_dictionary(nspots, len(nspots), len(aar_names), N_levels, coords_list, sf_list, aar_list,
                                 level_mappings, tm_list, W_list, Wn_list, True, False, False, False, [])
    np.savez_compressed(os.path.join(OUT, "c2_geometry.npz"),
                        N_spots=np.asarray(data["N_spots"]), N_covariates=data["N_covariates"],
                        tissue_mapping=np.asarray(data["tissue_mapping"]), N_level_1=data["N_level_1"],
                        N_level_2=data["N_level_2"], level_2_mapping=np.asarray(data["level_2_mapping"]),
                        D=np.asarray(data["D"]), W_sparse=np.asarray(data["W_sparse"]),
                        D_sparse=np.asarray(data["D_sparse"]), eig_values=np.asarray(data["eig_values"]),
                        U=np.asarray(data["U"]), V=np.asarray(data["V"]))
    print("c2 geometry: N=%d tissues=%d W_n=%d C=%d levels=%s" % (sum(nspots), len(nspots), sum(Wn_list),
                                                                 len(aar_names), N_levels))


def make_io(U):
    dst = os.path.join(OUT, "io")
    os.makedirs(dst, exist_ok=True)
    rng = np.random.default_rng(0)
    data = {"N_tissues": 2, "scalar_f": 0.125, "neg": -3, "N_spots": np.array([3, 4]),
            "size_factors": rng.lognormal(0, 1, 7), "empty": [], "W_sparse": np.array([[1, 2], [2, 3], [4, 5]]),
            "E": np.round(rng.dirichlet(np.ones(3), 7), 5), "Wempty": np.zeros((0, 0)),
            "counts": rng.poisson(3, 7)}
    U.to_rdump(data, os.path.join(dst, "ref_written.R"))
    back = U.read_rdump(os.path.join(dst, "ref_written.R"))
    np.savez(os.path.join(dst, "ref_written_parsed.npz"), **{k: np.asarray(v) for k, v in back.items() if k not in ("empty", "Wempty")})
    # a small Stan CSV as bin/splotch leaves it (combined_<g>.csv) and what read_stan_csv makes of it
    cols = ["lp__", "accept_stat__", "psi.1", "psi.2", "beta_raw.1.1", "beta_raw.2.1", "beta_raw.1.2", "beta_raw.2.2",
            "sigma", "theta.1", "log_lambda.1", "log_lambda.2", "beta_level_1.1.1", "beta_level_1.2.1",
            "beta_level_1.1.2", "beta_level_1.2.2"]
    vals = rng.normal(0, 1, (6, len(cols)))
    with open(os.path.join(dst, "combined_ref.csv"), "w") as f:
        f.write(",".join(cols) + "\n")
        for r in vals:
            f.write(",".join(repr(float(x)) for x in r) + "\n")
    variables = ["log_lambda", "beta_level_1", "psi", "sigma", "theta"]
    post = U.read_stan_csv(os.path.join(dst, "combined_ref.csv"), variables, efficient=True)
    np.savez(os.path.join(dst, "combined_ref_parsed.npz"), **post)
    print("io fixtures written")


def make_snap25(U):
    path = os.path.join(REF, "Science", "human", "data_SNAP25.R")
    # the reference's own read_rdump cannot parse this file (".Dim = c(" spacing, ", " separators:
    # its regex needs ".Dim=c("), so the whitespace-tolerant product reader is used here
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from csplotch_b200.rdump import read_rdump
    d = read_rdump(path)
    nT = 6
    ns = np.asarray(d["N_spots"])[:nT]
    n = int(ns.sum())
    W = np.asarray(d["W_sparse"])
    W = W[(W[:, 0] <= n) & (W[:, 1] <= n)]
    # eigenvalues of the truncated design must be recomputed for the kept tissues
    np.savez_compressed(os.path.join(OUT, "snap25_head.npz"), N_spots=ns, tissue_mapping=np.asarray(d["tissue_mapping"])[:nT],
                        donor_mapping=np.asarray(d["donor_mapping"]), N_donors=d["N_donors"],
                        N_onsets_tissue_locations=d["N_onsets_tissue_locations"], N_covariates=d["N_covariates"],
                        size_factors=np.asarray(d["size_factors"])[:n], counts=np.asarray(d["counts"])[:n],
                        D=np.asarray(d["D"])[:n], D_sparse=np.asarray(d["D_sparse"])[:n], W_sparse=W,
                        full_N=int(np.sum(d["N_spots"])), full_W_n=int(d["W_n"]),
                        full_counts_sum=int(np.sum(d["counts"])), full_zero_frac=float(np.mean(np.asarray(d["counts"]) == 0)))
    print("snap25 head: N=%d pairs=%d (full N=%d)" % (n, len(W), int(np.sum(d["N_spots"]))))


def make_snap25_full():
    """The WHOLE Science/human/data_SNAP25.R (80 tissue sections, N = 61 031, W_n = 112 012, the file's own eig_values with
    88 entries exactly +1): the shape the full-size parity tests use.  Parsed with the
    whitespace-tolerant product reader for the reason stated in make_snap25()."""
    path = os.path.join(REF, "Science", "human", "data_SNAP25.R")
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from csplotch_b200.rdump import read_rdump
    d = read_rdump(path)
    np.savez_compressed(os.path.join(OUT, "snap25_full.npz"),
                        N_spots=np.asarray(d["N_spots"], dtype=np.int32), tissue_mapping=np.asarray(d["tissue_mapping"], dtype=np.int32),
                        donor_mapping=np.asarray(d["donor_mapping"], dtype=np.int32), N_donors=int(d["N_donors"]),
                        N_onsets_tissue_locations=int(d["N_onsets_tissue_locations"]), N_covariates=int(d["N_covariates"]),
                        size_factors=np.asarray(d["size_factors"]), counts=np.asarray(d["counts"], dtype=np.int32),
                        D=np.asarray(d["D"], dtype=np.int8), D_sparse=np.asarray(d["D_sparse"], dtype=np.int8),
                        W_sparse=np.asarray(d["W_sparse"], dtype=np.int32), eig_values=np.asarray(d["eig_values"]))
    print("snap25 full: N=%d pairs=%d eig==1: %d" % (int(np.sum(d["N_spots"])), len(d["W_sparse"]),
                                                     int(np.sum(np.asarray(d["eig_values"]) == 1.0))))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "snap25_full":
        make_snap25_full()
        sys.exit(0)
    U, gen = load_reference()
    make_io(U)
    make_c1(U, gen)
    make_c2_geometry(U, gen)
    make_snap25(U)
