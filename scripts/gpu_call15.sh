#!/bin/bash
# round 2, GPU call 15: two merges of a cascade per pass (dev) against one by one (devm1: the same source with -DCSPLOTCH_NO_MERGE2)
mkdir -p gpurun_out
L=$PWD/csplotch_b200
export CSPLOTCH_LIB=$L/libcsplotch_b200_dev.so
(timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cluster.py tests/test_gpu_fullsize.py tests/test_gpu_longchain.py -m gpu -q --deselect tests/test_gpu_cluster.py::test_cluster_size_follows_the_number_of_chains 2>&1 | tail -60) > gpurun_out/r02_pytest_call15.log
grep -c "development build" gpurun_out/r02_pytest_call15.log; grep FAILED gpurun_out/r02_pytest_call15.log | cut -c1-150; tail -1 gpurun_out/r02_pytest_call15.log
for v in devm1 dev devm1 dev; do
  CSPLOTCH_LIB=$L/libcsplotch_b200_$v.so PROBE_G=1184 timeout -s KILL 600 python scripts/perf_probe.py c2 c3 > gpurun_out/r02_ab15_probe_$v.log 2>&1
  echo "== probe $v"; grep -h '"iters_done"\|profile' gpurun_out/r02_ab15_probe_$v.log | cut -c1-330
done
