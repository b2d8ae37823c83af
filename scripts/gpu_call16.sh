#!/bin/bash
# round 2, GPU call 16 (product library): the whole GPU suite, smoke(), the driver-style bench lines (ours + reference arm), the ncu launch
# list of the bench command, one `ncu --set full` capture of a c3 sampler launch at normal tree depth (DRAM traffic), vinit re-measurement
mkdir -p gpurun_out
t0=$SECONDS
(timeout -s KILL 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15) | tee gpurun_out/r02_pytest_gpu_full.log
(timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) | tee gpurun_out/r02_smoke.log
echo "== tests + smoke: $((SECONDS-t0)) s"; t0=$SECONDS
timeout -s KILL 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; tail -8 gpurun_out/r02_bench_c3.err
echo "== bench: $((SECONDS-t0)) s"; t0=$SECONDS
timeout -s KILL 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref_c3.json 2> gpurun_out/r02_bench_ref_c3.err; cut -c1-500 gpurun_out/r02_bench_ref_c3.json
echo "== reference arm: $((SECONDS-t0)) s"; t0=$SECONDS
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_bench_c3.json"))
    print("r02_bench_c3 value", round(d["value"]), "frac", round(d["roofline"]["frac"], 4), "tail", round(d["roofline"]["launch_tail_frac"], 4), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"]), d.get("ess"), d.get("secondary"), d.get("cpu_baseline"))
except Exception as e:
    print("bench FAILED", e)
PY
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench_c3.csv python bench.py --genes 74 --steps 2 --warmup 3 --no-cpu --ess-genes 0 --no-secondary --e2e-genes 8 > gpurun_out/r02_ncu_bench.log 2>&1
echo "== launch list: $((SECONDS-t0)) s"; t0=$SECONDS
NCU_SKIP=1 timeout -s KILL 1200 bash scripts/ncu_capture.sh nuts_c3 k_nuts_advance scripts/profile_nuts.py c3 74 1 8
python scripts/traffic_from_ncu.py nuts_c3 c3 && cp profiles/traffic_c3.json gpurun_out/traffic_c3.json
head -30 gpurun_out/lines_nuts_c3.txt
echo "== traffic capture: $((SECONDS-t0)) s"; t0=$SECONDS
timeout -s KILL 600 python scripts/vinit_compare.py 16 200 200 > gpurun_out/r02_vinit_compare_c2.json 2> gpurun_out/r02_vinit_compare_c2.err; cut -c1-600 gpurun_out/r02_vinit_compare_c2.json
echo "== vinit: $((SECONDS-t0)) s"
