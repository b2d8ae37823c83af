#!/bin/bash
# round 2, GPU call 21 (product library built without --split-compile): the whole GPU suite, smoke(), the fused leapfrog on c2 / c3 / c4,
# the driver-style bench line
mkdir -p gpurun_out
t0=$SECONDS
(timeout -s KILL 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15) | tee gpurun_out/r02_pytest_gpu_full.log
(timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) | tee gpurun_out/r02_smoke.log
echo "== tests + smoke: $((SECONDS-t0)) s"; t0=$SECONDS
CSPLOTCH_CLUSTER=1 timeout -s KILL 300 python scripts/leap_bench.py c2 c3 c4 --waves 1 --steps 6 2>&1 | grep '^{' | sed "s/^{/{\"v\": \"product\", /" | tee gpurun_out/r02_leap_bench_product.jsonl
echo "== leap bench: $((SECONDS-t0)) s"; t0=$SECONDS
timeout -s KILL 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; tail -8 gpurun_out/r02_bench_c3.err
echo "== bench: $((SECONDS-t0)) s"; t0=$SECONDS
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_bench_c3.json"))
    print("r02_bench_c3 value", round(d["value"]), "frac", round(d["roofline"]["frac"], 4), "tail", round(d["roofline"]["launch_tail_frac"], 4), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"]), d.get("ess"), d.get("secondary"), d.get("cpu_baseline"))
except Exception as e:
    print("bench FAILED", e)
PY
