#!/bin/bash
# round 2, GPU call 6 (product library): whole GPU suite, smoke(), the driver-style bench line on c3
mkdir -p gpurun_out
(timeout -s KILL 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15) | tee gpurun_out/r02_pytest_gpu_full.log
(timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) | tee gpurun_out/r02_smoke.log
timeout -s KILL 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; tail -8 gpurun_out/r02_bench_c3.err
python - <<'PY'
import json
for f in ("r02_bench_c3",):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value", round(d["value"]), "frac", round(d["roofline"]["frac"], 4), "tail", round(d["roofline"]["launch_tail_frac"], 4), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"]), d.get("ess"), d.get("secondary"), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "FAILED", e)
PY
