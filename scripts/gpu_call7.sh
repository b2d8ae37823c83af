#!/bin/bash
# round 2, GPU call 7 (development library): warp-per-chain path (tests, c2 A/B), then the per-line ncu profile of the c3 leapfrog
mkdir -p gpurun_out
export CSPLOTCH_LIB=$PWD/csplotch_b200/libcsplotch_b200_dev.so
(timeout -s KILL 900 python -m pytest tests/test_gpu_warp.py -m gpu -q 2>&1 | tail -15) | tee gpurun_out/r02_pytest_warp.log
for sc in cta warp; do
  CSPLOTCH_SCOPE=$sc timeout -s KILL 300 python scripts/leap_bench.py c2 --waves 4 2>&1 | grep '^{' | sed "s/^{/{\"scope\": \"$sc\", /" | tee -a gpurun_out/r02_leap_bench_call7.jsonl
done
for sc in cta warp; do
  CSPLOTCH_SCOPE=$sc PROBE_G=1184 timeout -s KILL 600 python scripts/perf_probe.py c2 > gpurun_out/r02_probe_c2_$sc.log 2>&1
  echo "== probe c2 scope=$sc"; grep -h '"iters_done"\|profile\|layout' gpurun_out/r02_probe_c2_$sc.log | tail -n 5
done
unset CSPLOTCH_SCOPE
export CSPLOTCH_CLUSTER=1
timeout -s KILL 300 python scripts/leap_bench.py c3 --waves 1 --steps 2 > gpurun_out/plain_leap_c3.log 2>&1 || { tail -5 gpurun_out/plain_leap_c3.log; exit 1; }
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:k_leapfrog_fixed -s 1 -c 1 -o /tmp/prof_leap_c3 python scripts/leap_bench.py c3 --waves 1 --steps 2 > gpurun_out/ncu_leap_c3.log 2>&1
python scripts/ncu_lines.py /tmp/prof_leap_c3.ncu-rep 70 > gpurun_out/r02_ncu_leapfrog_c3_lines.txt 2>&1
ncu -i /tmp/prof_leap_c3.ncu-rep --page raw --csv > gpurun_out/r02_ncu_leapfrog_c3_raw.csv 2>/dev/null
head -40 gpurun_out/r02_ncu_leapfrog_c3_lines.txt
