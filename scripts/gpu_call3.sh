#!/bin/bash
# round 2, GPU call 3: the cluster-per-chain path (development library: K in {1, 10}): parity, leapfrog microbenchmark and sampler probe
# against CSPLOTCH_CLUSTER=1.  Every step under its own timeout (a cluster-barrier mismatch would hang).
mkdir -p gpurun_out
export CSPLOTCH_LIB=$PWD/csplotch_b200/libcsplotch_b200_dev.so
(timeout -s KILL 900 python -m pytest tests/test_gpu_cluster.py -m gpu -q -x 2>&1 | tail -15) | tee gpurun_out/r02_pytest_cluster.log
(timeout -s KILL 600 python -m pytest tests/test_gpu_car.py -m gpu -q 2>&1 | tail -15) | tee gpurun_out/r02_pytest_car.log
for cl in 1 2 4; do
  CSPLOTCH_CLUSTER=$cl timeout -s KILL 300 python scripts/leap_bench.py c3 2>&1 | grep '^{' | sed "s/^{/{\"cl\": $cl, /" | tee -a gpurun_out/r02_leap_bench_cluster.jsonl
done
for cl in 1 2; do
  CSPLOTCH_CLUSTER=$cl timeout -s KILL 300 python scripts/leap_bench.py c4 --steps 4 2>&1 | grep '^{' | sed "s/^{/{\"cl\": $cl, /" | tee -a gpurun_out/r02_leap_bench_cluster.jsonl
done
(timeout -s KILL 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "c3 or c4" 2>&1 | tail -5) | tee gpurun_out/r02_pytest_fullsize_cluster.log
for cl in 1 4; do
  CSPLOTCH_CLUSTER=$cl timeout -s KILL 600 python scripts/perf_probe.py c3 > gpurun_out/r02_probe_c3_cl$cl.log 2>&1
  echo "== probe c3 cl=$cl"; grep -h '"iters_done"\|profile' gpurun_out/r02_probe_c3_cl$cl.log | tail -n 5
done
