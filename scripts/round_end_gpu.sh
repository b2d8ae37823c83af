#!/bin/bash
# Round-end measurement set on one B200 box (round 2): GPU tests, smoke, fused-leapfrog microbenchmark, the driver-style bench
# lines (ours + reference arm), ncu launch list of the bench command, DRAM traffic of one sampler launch per shape, line
# profile of the fused pass.  ~25 GPU-minutes.  Outputs under gpurun_out/; copy what should be judged into profiles/.
mkdir -p gpurun_out
(timeout -s KILL 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15) | tee gpurun_out/pytest_gpu_full.log
(timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) | tee gpurun_out/smoke.log
CSPLOTCH_CLUSTER=1 timeout -s KILL 300 python scripts/leap_bench.py c2 c3 c4 --waves 1 --steps 6 2>&1 | grep '^{' | tee gpurun_out/leap_bench_product.jsonl
timeout -s KILL 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -8 gpurun_out/bench_c3.err
timeout -s KILL 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_c3.json 2> gpurun_out/bench_ref_c3.err; cut -c1-600 gpurun_out/bench_ref_c3.json
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_bench_c3.csv python bench.py --genes 74 --steps 2 --warmup 3 --no-cpu --ess-genes 0 --no-secondary --e2e-genes 8 > gpurun_out/ncu_bench.log 2>&1
bash scripts/traffic_quick.sh c3 74 8; bash scripts/traffic_quick.sh c2 1184 20
bash scripts/gpu_call23.sh      # ncu --set full of k_leapfrog_fixed at c3: per source line + raw page
