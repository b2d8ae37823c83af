#!/bin/bash
# round 2, GPU call 13: neighbour positions in scalar registers (no local-memory round trip across the ring wait), 12-exchange butterfly for
# K = 10, running bin totals (dev) against a5e9c90 (head)
mkdir -p gpurun_out
L=$PWD/csplotch_b200
export CSPLOTCH_LIB=$L/libcsplotch_b200_dev.so
(timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cluster.py tests/test_gpu_fullsize.py -m gpu -q --deselect tests/test_gpu_cluster.py::test_cluster_size_follows_the_number_of_chains 2>&1 | tail -40) > gpurun_out/r02_pytest_call13.log
grep -c "development build" gpurun_out/r02_pytest_call13.log; grep "FAILED" gpurun_out/r02_pytest_call13.log | grep -v "development build" | head; tail -1 gpurun_out/r02_pytest_call13.log
export CSPLOTCH_CLUSTER=1
for v in head dev head dev; do
  export CSPLOTCH_LIB=$L/libcsplotch_b200_$v.so
  timeout -s KILL 300 python scripts/leap_bench.py c2 c3 c4 --waves 1 --steps 6 2>&1 | grep '^{' | sed "s/^{/{\"v\": \"$v\", /" | tee -a gpurun_out/r02_ab_call13.jsonl
done
for v in head dev; do
  CSPLOTCH_LIB=$L/libcsplotch_b200_$v.so timeout -s KILL 600 python scripts/perf_probe.py c3 > gpurun_out/r02_ab13_probe_$v.log 2>&1
  echo "== probe $v"; grep -h '"iters_done"\|profile' gpurun_out/r02_ab13_probe_$v.log | tail -n 4
done
