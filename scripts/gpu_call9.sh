#!/bin/bash
# round 2, GPU call 9: same-box A/B of the library at the start of this session's kernel work (base) against the current one (dev)
mkdir -p gpurun_out
L=$PWD/csplotch_b200
export CSPLOTCH_CLUSTER=1
for v in base dev base dev; do
  CSPLOTCH_LIB=$L/libcsplotch_b200_$v.so timeout -s KILL 300 python scripts/leap_bench.py c2 c3 2>&1 | grep '^{' | tee -a gpurun_out/r02_ab_call9.jsonl
done
for v in base dev; do
  CSPLOTCH_LIB=$L/libcsplotch_b200_$v.so timeout -s KILL 300 python scripts/leap_bench.py c4 --steps 4 2>&1 | grep '^{' | tee -a gpurun_out/r02_ab_call9.jsonl
done
for v in base dev; do
  CSPLOTCH_LIB=$L/libcsplotch_b200_$v.so PROBE_G=1184 timeout -s KILL 900 python scripts/perf_probe.py c2 c3 > gpurun_out/r02_ab_probe_$v.log 2>&1
  echo "== probe $v"; grep -h '"iters_done"\|profile' gpurun_out/r02_ab_probe_$v.log | tail -n 8
done
unset CSPLOTCH_CLUSTER
export CSPLOTCH_LIB=$L/libcsplotch_b200_dev.so
(timeout -s KILL 900 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_fullsize.py tests/test_gpu_warp.py -m gpu -q 2>&1 | tail -4) | tee gpurun_out/r02_pytest_call9.log
