#!/bin/bash
# A/B on the GPU box: the product library (A) against the development build (B, csplotch_b200.build --dev)
# usage: scripts/exp_ab.sh <tag> [shapes...]     results under gpurun_out/ab_<tag>_{A,B}.log
tag=$1; shift
shapes=${@:-c2}
mkdir -p gpurun_out
DEV=$PWD/csplotch_b200/libcsplotch_b200_dev.so
(CSPLOTCH_LIB=$DEV timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -25) > gpurun_out/ab_${tag}_pytestB.log
for side in A B; do
  if [ $side = B ]; then export CSPLOTCH_LIB=$DEV; else unset CSPLOTCH_LIB; fi
  PROBE_G=${PROBE_G:-1184} timeout 900 python scripts/perf_probe.py $shapes > gpurun_out/ab_${tag}_$side.log 2>&1
  grep -h '"iters_done"' gpurun_out/ab_${tag}_$side.log | tail -n 6
done
grep -c "development build" gpurun_out/ab_${tag}_pytestB.log; tail -n 3 gpurun_out/ab_${tag}_pytestB.log
