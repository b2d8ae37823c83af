#!/bin/bash
# round 2, GPU call 8 (development library): TMA-staged merge pass; per-line ncu profile of the c3 leapfrog
mkdir -p gpurun_out
export CSPLOTCH_LIB=$PWD/csplotch_b200/libcsplotch_b200_dev.so
(timeout -s KILL 900 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_fullsize.py tests/test_gpu_warp.py -m gpu -q 2>&1 | tail -6) | tee gpurun_out/r02_pytest_call8.log
(timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | grep -v "development build" | tail -6) | tee gpurun_out/r02_pytest_parity_call8.log
PROBE_G=1184 timeout -s KILL 600 python scripts/perf_probe.py c2 > gpurun_out/r02_probe_c2_call8.log 2>&1; grep -h '"iters_done"\|profile\|layout' gpurun_out/r02_probe_c2_call8.log | tail -n 5
for cl in 1 4; do
  CSPLOTCH_CLUSTER=$cl timeout -s KILL 600 python scripts/perf_probe.py c3 > gpurun_out/r02_probe_c3_call8_cl$cl.log 2>&1
  echo "== probe c3 cl=$cl"; grep -h '"iters_done"\|profile\|layout' gpurun_out/r02_probe_c3_call8_cl$cl.log | tail -n 5
done
export CSPLOTCH_CLUSTER=1
timeout -s KILL 300 python scripts/leap_bench.py c3 --waves 1 --steps 2 > gpurun_out/plain_leap_c3.log 2>&1 || { tail -5 gpurun_out/plain_leap_c3.log; exit 1; }
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:k_leapfrog_fixed -s 1 -c 1 -o /tmp/prof_leap_c3 python scripts/leap_bench.py c3 --waves 1 --steps 2 > gpurun_out/ncu_leap_c3.log 2>&1
python scripts/ncu_lines.py /tmp/prof_leap_c3.ncu-rep 70 > gpurun_out/r02_ncu_leapfrog_c3_lines.txt 2>&1
ncu -i /tmp/prof_leap_c3.ncu-rep --page raw --csv > gpurun_out/r02_ncu_leapfrog_c3_raw.csv 2>/dev/null
head -45 gpurun_out/r02_ncu_leapfrog_c3_lines.txt
