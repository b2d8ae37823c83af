#!/bin/bash
# round 2, GPU call 1: fp64 roofs, barrier-variant A/B of the fused leapfrog (microbenchmark + sampler), parity of the new default
mkdir -p gpurun_out
tools/fp64_peak > gpurun_out/fp64_peak.json 2>&1; cat gpurun_out/fp64_peak.json
L=$PWD/csplotch_b200
for v in r1 elect nohint both; do
  CSPLOTCH_LIB=$L/libcsplotch_b200_$v.so timeout 600 python scripts/leap_bench.py c2 c3 2>&1 | grep '^{' | tee -a gpurun_out/leap_bench_call1.jsonl
done
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -8) | tee gpurun_out/parity_product.log
for v in r1 both; do
  CSPLOTCH_LIB=$L/libcsplotch_b200_$v.so PROBE_G=1184 timeout 900 python scripts/perf_probe.py c2 c3 > gpurun_out/probe_$v.log 2>&1
  echo "== probe $v"; grep -h '"iters_done"\|profile' gpurun_out/probe_$v.log | tail -n 10
done
