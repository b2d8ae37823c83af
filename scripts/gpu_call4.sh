#!/bin/bash
# round 2, GPU call 4 (development library): staged small block + fused epilogue reduction, clusters with per-launch size
mkdir -p gpurun_out
export CSPLOTCH_LIB=$PWD/csplotch_b200/libcsplotch_b200_dev.so
(timeout -s KILL 900 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_car.py -m gpu -q 2>&1 | tail -15) | tee gpurun_out/r02_pytest_cluster2.log
(timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -15) | tee gpurun_out/r02_pytest_parity_dev.log
(timeout -s KILL 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -5) | tee gpurun_out/r02_pytest_fullsize2.log
for cl in 1 2 4; do
  CSPLOTCH_CLUSTER=$cl timeout -s KILL 300 python scripts/leap_bench.py c3 2>&1 | grep '^{' | sed "s/^{/{\"cl\": $cl, /" | tee -a gpurun_out/r02_leap_bench_call4.jsonl
done
CSPLOTCH_CLUSTER=1 timeout -s KILL 300 python scripts/leap_bench.py c2 2>&1 | grep '^{' | sed "s/^{/{\"cl\": 1, /" | tee -a gpurun_out/r02_leap_bench_call4.jsonl
for cl in 1 2 4; do
  CSPLOTCH_CLUSTER=$cl timeout -s KILL 300 python scripts/leap_bench.py c4 --steps 4 2>&1 | grep '^{' | sed "s/^{/{\"cl\": $cl, /" | tee -a gpurun_out/r02_leap_bench_call4.jsonl
done
PROBE_G=1184 timeout -s KILL 600 python scripts/perf_probe.py c2 > gpurun_out/r02_probe_c2_call4.log 2>&1; grep -h '"iters_done"\|profile\|layout' gpurun_out/r02_probe_c2_call4.log | tail -n 5
for cl in 1 4; do
  CSPLOTCH_CLUSTER=$cl timeout -s KILL 600 python scripts/perf_probe.py c3 > gpurun_out/r02_probe_c3_call4_cl$cl.log 2>&1
  echo "== probe c3 cl=$cl"; grep -h '"iters_done"\|profile\|layout' gpurun_out/r02_probe_c3_call4_cl$cl.log | tail -n 5
done
