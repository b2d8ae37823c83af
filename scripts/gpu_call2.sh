#!/bin/bash
# round 2, GPU call 2: new parity tests + whole GPU suite, the driver-style bench line on c3, tail before/after the longest-first queue
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_longchain.py tests/test_stan_vectors.py -m gpu -q 2>&1 | tail -15) | tee gpurun_out/r02_pytest_new.log
(timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py --deselect tests/test_gpu_longchain.py 2>&1 | tail -8) | tee gpurun_out/r02_pytest_gpu.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; tail -12 gpurun_out/r02_bench_c3.err; cut -c1-600 gpurun_out/r02_bench_c3.json
CSPLOTCH_QUEUE_ORDER=index timeout 600 python bench.py --steps 8 --warmup 5 --no-cpu --ess-genes 0 --no-secondary --e2e-genes 74 --e2e-iters 4 > gpurun_out/r02_bench_c3_indexorder.json 2> gpurun_out/r02_bench_c3_indexorder.err
python - <<'PY'
import json
for f in ("r02_bench_c3", "r02_bench_c3_indexorder"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value", round(d["value"]), "frac", round(d["roofline"]["frac"], 4), "tail", round(d["roofline"]["launch_tail_frac"], 4), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"]), d.get("ess", {}).get("genes_per_hour"), d.get("secondary"))
    except Exception as e:
        print(f, "FAILED", e)
PY
