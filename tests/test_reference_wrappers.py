"""The reference's UNMODIFIED bash wrappers driving this repo's CmdStan shim:
/root/reference/bin/splotch and /root/reference/bin/splotch_vinit are executed as they are, with `-b` pointing at a
stand-in executable named like the model.  On this box (no GPU) the stand-in answers from the CPU oracle
(tests/wrapper/fake_binary.py); the argv parsing, the Pathfinder CSV, the wrapper's own python that turns it into
init_<g>_<c>.json, `init=`, and the grep / sed concatenation are the real thing.  Where /root/reference is absent the
same flow runs through tests/wrapper/run_wrapper.sh (restated)."""
import os
import subprocess

import numpy as np
import pytest

from csplotch_b200 import stancsv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = "/root/reference/bin"


def _setup(golden_dir, tmp_path, g):
    d = tmp_path / "data" / "0"
    d.mkdir(parents=True)
    (d / ("data_%d.R" % g)).write_text(open(os.path.join(golden_dir, "c1", "data_%d.R" % g)).read())
    binary = tmp_path / "bin" / "comp_splotch_stan_model"
    binary.parent.mkdir()
    os.symlink(os.path.join(ROOT, "tests", "wrapper", "fake_binary.py"), binary)
    return str(tmp_path / "data"), str(tmp_path / "out"), str(binary)


def _check(out, g, chains, n):
    p = os.path.join(out, "0", "combined_%d.csv" % g)
    lines = open(p).read().splitlines()
    assert lines[0].startswith("lp__,accept_stat__") and len(lines) == 1 + chains * n
    assert not any(l.startswith(("#", "l")) for l in lines[1:])
    post = stancsv.read_stan_csv(p, stancsv.SUMMARY_VARS + ["lp__"])
    assert post["beta_level_1"].shape == (chains * n, 1, 14, 5) and post["log_lambda"].shape == (chains * n, 10)
    assert np.isfinite(post["lp__"]).all() and (post["theta"] > 0).all() and (post["theta"] < 1).all()
    left = sorted(os.listdir(os.path.join(out, "0")))
    assert "output_%d_1.csv" % g not in left and "combined_%d.csv" % g in left
    return left


@pytest.mark.parametrize("which", ["reference", "restated"])
def test_unmodified_splotch_wrapper(which, golden_dir, tmp_path):
    data, out, binary = _setup(golden_dir, tmp_path, 4)
    if which == "reference":
        if not os.path.exists(os.path.join(REF_BIN, "splotch")):
            pytest.skip("/root/reference not on this box")
        cmd = ["bash", os.path.join(REF_BIN, "splotch"), "-g", "4", "-d", data, "-o", out, "-b", binary, "-n", "12", "-c", "3"]
    else:
        cmd = ["bash", os.path.join(ROOT, "tests", "wrapper", "run_wrapper.sh"), "plain", "4", data, out, binary, "12", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    _check(out, 4, 3, 12)


@pytest.mark.parametrize("which", ["reference", "restated"])
def test_unmodified_splotch_vinit_wrapper(which, golden_dir, tmp_path):
    data, out, binary = _setup(golden_dir, tmp_path, 6)
    if which == "reference":
        if not os.path.exists(os.path.join(REF_BIN, "splotch_vinit")):
            pytest.skip("/root/reference not on this box")
        cmd = ["bash", os.path.join(REF_BIN, "splotch_vinit"), "-g", "6", "-d", data, "-o", out, "-b", binary, "-n", "10", "-c", "2"]
    else:
        cmd = ["bash", os.path.join(ROOT, "tests", "wrapper", "run_wrapper.sh"), "vinit", "6", data, out, binary, "10", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    left = _check(out, 6, 2, 10)
    assert "init_6_1.json" in left and "init_6_2.json" in left and "output_6.csv" not in left
    import json
    init = json.load(open(os.path.join(out, "0", "init_6_1.json")))
    assert "tau.1" in init and "psi.10" in init and "lp__" not in init and "path__" not in init


def test_a_failing_binary_fails_the_wrapper(golden_dir, tmp_path):
    """`|| error_exit "Error: Stan run failed!"` (bin/splotch:131): exit status 1 of the shim reaches the wrapper."""
    data, out, binary = _setup(golden_dir, tmp_path, 4)
    os.remove(os.path.join(data, "0", "data_4.R"))
    r = subprocess.run(["bash", os.path.join(ROOT, "tests", "wrapper", "run_wrapper.sh"), "plain", "4", data, out, binary,
                        "5", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "Error reading data file" in r.stderr


def test_unmodified_wrapper_in_summary_mode(golden_dir, tmp_path):
    """`splotch … -s` (bin/splotch:145-151): the wrapper calls `splotch_summarize_output -p CSV -o HDF5` from PATH and
    removes the CSV; with this repo's bin/ on PATH the summary holds the groups summarize_stan_hdf5 writes."""
    if not os.path.exists(os.path.join(REF_BIN, "splotch")):
        pytest.skip("/root/reference not on this box")
    data, out, binary = _setup(golden_dir, tmp_path, 2)
    env = dict(os.environ, PATH=os.path.join(ROOT, "bin") + os.pathsep + os.environ["PATH"])
    r = subprocess.run(["bash", os.path.join(REF_BIN, "splotch"), "-g", "2", "-d", data, "-o", out, "-b", binary, "-n", "9",
                        "-c", "2", "-s"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert not os.path.exists(os.path.join(out, "0", "combined_2.csv"))                  # removed by the wrapper
    got = stancsv.read_summary(os.path.join(out, "0", "combined_2.hdf5"))
    assert set(got) == {"log_lambda", "lambda", "beta_level_1", "psi", "sigma", "theta", "tau", "a"}
    assert got["beta_level_1"]["samples"].shape == (18, 1, 14, 5) and got["lambda"]["mean"].shape == (10,)
    assert got["sigma"]["mean"].shape == (1,) and set(got["psi"]) == {"mean", "std"}
