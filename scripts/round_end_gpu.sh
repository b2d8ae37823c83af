#!/bin/bash
# Round-end measurement set on one B200 box: GPU tests, default bench + reference arm, ncu launch list of the bench
# command, one `ncu --set full` capture of the sampler kernel per shape (DRAM traffic), line profile of the fused pass.
mkdir -p gpurun_out
(python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/pytest_gpu.log; tail -1 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cut -c1-400 gpurun_out/bench_c2.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_c2.json 2>> gpurun_out/bench_c2.err; cut -c1-300 gpurun_out/bench_ref_c2.json
python bench.py --workload c3 --no-cpu --steps 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cut -c1-400 gpurun_out/bench_c3.json
python bench.py --no-cpu --steps 2 --e2e-genes 4736 > gpurun_out/bench_c2_e2e4736.json 2>> gpurun_out/bench_c2.err; python -c "import json;d=json.load(open('gpurun_out/bench_c2_e2e4736.json'));print('e2e 4736 genes', d['e2e']['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_c2.csv python bench.py --genes 592 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
for shape in c2 c3; do
  G=592; [ $shape = c3 ] && G=148
  bash scripts/ncu_capture.sh nuts_$shape k_nuts_advance scripts/profile_nuts.py $shape $G 2
  python scripts/traffic_from_ncu.py nuts_$shape $shape
done
# splotch_vinit against plain splotch on c2 genes: re-measure after the Pathfinder line-search fix,
# with CmdStan's iteration cap (1000) replaced by 200 to bound the host time
python scripts/vinit_compare.py 16 200 200 > gpurun_out/vinit_compare_c2.json 2> gpurun_out/vinit_compare_c2.err; cut -c1-400 gpurun_out/vinit_compare_c2.json
# second half of the BASELINE metric inside the bench line itself (opt-in leg): genes / hour to bulk-ESS >= 400
python bench.py --no-cpu --steps 2 --ess-genes 592 --ess-n 200 > gpurun_out/bench_c2_ess.json 2> gpurun_out/bench_c2_ess.err; python -c "import json;d=json.load(open('gpurun_out/bench_c2_ess.json'));print('ess', d['ess'])"
