"""Host-side pieces of bench.py that the driver's contract depends on, without a GPU: the clocks / throttle-reason
sampler (a stand-in nvidia-smi on PATH), the roofline peak lookup, and the reference arm's JSON line (stdout carries
that line and nothing else)."""
import json
import os
import stat
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_clock_sampler_parses_clocks_and_reasons(tmp_path, monkeypatch):
    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/bin/bash\n"
                    "echo '0, 1965, 1965, 410.2, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active'\n"
                    "echo '0, 1950, 1965, 690.0, 0x0000000000000004, Not Active, Not Active, Not Active, Active'\n"
                    "echo '0, 1200, 1965, 700.1, 0x0000000000000040, Not Active, Active, Not Active, Active'\n"
                    "echo 'garbage line'\n"
                    "sleep 30\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    cs = bench.ClockSampler(0)
    cs.start()
    time.sleep(0.5)
    out = cs.stop()
    assert out["samples"] == 3 and out["sm_mhz"] == 1950.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"]


def test_clock_sampler_without_nvidia_smi(monkeypatch, tmp_path):
    monkeypatch.setenv("PATH", str(tmp_path))
    cs = bench.ClockSampler(0)
    cs.start()
    assert cs.stop()["reasons"] == ["nvidia-smi unavailable"]


def test_roofline_peak_is_the_measured_one():
    peak, src = bench.peaks()
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        assert peak == float(json.load(open(p))["hbm_gbs"]) and src.startswith("measured")
    else:
        assert src.startswith("fallback") and 6000 < peak < 8000


def test_reference_arm_prints_exactly_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-iters", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "grad_evals_per_sec" and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # under torchrun only rank 0 works
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=60, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and r.stdout == ""


def test_ess_leg_with_oracle_backed_device_classes(monkeypatch):
    """`bench.py --ess-genes`: the genes/hour-to-ESS>=400 leg on stand-ins for the device classes (tests only), so the
    bookkeeping — which draws enter which ESS, rates, failed chains — is checked without a GPU."""
    import numpy as np
    from csplotch_b200 import api, rdump
    from tests import test_vinit_host as fakes

    class Sampler(fakes._Sampler):
        def run(self, chunk=None):
            super().run(chunk)
            self.total_ms = 1800.0               # 1.8 s of "device time"
            return self

    monkeypatch.setattr(api, "Model", fakes._Model)
    monkeypatch.setattr(api, "Batch", fakes._Batch)
    monkeypatch.setattr(api, "Sampler", Sampler)
    gdir = os.path.join(ROOT, "tests", "golden", "c1")
    datas = [rdump.read_rdump(os.path.join(gdir, "data_%d.R" % g)) for g in (1, 2, 3)]
    shared = {k: v for k, v in datas[0].items() if k not in rdump.GENE_KEYS}
    genes = {k: np.stack([d[k] for d in datas]) for k in rdump.GENE_KEYS}
    model = api.Model(shared, "comp_splotch_stan_model")
    out = bench.ess_leg(api, model, genes, 3, 60, chains=4, ess_genes=2)
    assert out["genes"] == 3 and out["genes_evaluated_for_ess"] == 2 and out["num_warmup"] == out["num_samples"] == 60
    assert abs(out["genes_per_hour"] - 3 / 1.8 * 3600) < 1e-6 and out["grad_evals"] > 0
    assert 0 <= out["genes_per_hour_to_ess400"] <= out["beta_only"]["genes_per_hour_to_ess400"] <= out["genes_per_hour"]
    assert 0 < out["median_min_bulk_ess"] <= out["beta_only"]["median_min_bulk_ess"] < 4 * 60 * 3
    assert out["beta_only"]["median_max_rhat"] > 0.95


def test_gpu_arm_control_flow_with_mocked_device(tmp_path):
    """bench.py's own arm end to end on a box without a GPU (tests/wrapper/mock_bench.py: torch.cuda mocked, device
    classes replaced by oracle-backed stand-ins): phases, counters, roofline arithmetic, the opt-in ESS leg and the JSON
    line — which must be the only thing on stdout and carry every key of the driver's contract."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "wrapper", "mock_bench.py")], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"):
        assert key in d, key
    assert d["metric"] == "grad_evals_per_sec" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["steps"] == 2 and d["gpu_launches"] == 2 and d["n_gpus"] == 1 and d["scaling"] == "strong"
    # 2 steps x 1 transition x 100 "gradients" in 2 x 5 ms of "device time" (the mock's counters)
    assert abs(d["value"] - 2 * 100 * d["config"]["nuts_transitions_per_chain_per_step"] / 0.010) < 1e-6
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    D, N = 94, 10
    assert rf["algorithmic_bytes_per_leapfrog"] == 56 * D + 4 * N and rf["algorithmic_bytes_per_gradient"] == 16 * D + 4 * N
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["unit"] == d["unit"]
    assert d["ess"]["genes"] == 2 and d["ess"]["num_samples"] == 10 and "beta_only" in d["ess"]
    assert d["genes_per_hour_to_ess400"] == d["ess"]["genes_per_hour_to_ess400"] and d["ess"]["transcriptome_20k_genes_minutes"] > 0
    assert abs(rf["launch_tail_frac"] - 0.1) < 1e-9 and abs(sum(rf["time_shares"].values()) - 1) < 1e-9
    assert "EXTRAPOLATED" in (rf["traffic_source"] or "EXTRAPOLATED")
    assert "ESS leg done" in r.stderr and "e2e arm" in r.stderr
