#!/bin/bash
# round 2, GPU call 20: nvcc --split-compile 0 is NOT deterministic for this translation unit: three sequential builds of one source gave
# 944 / 944 / 912-byte frames for k_leapfrog_fixed<10>.  Same source: s1 (944), s3 (912), sc1 (--split-compile 1, deterministic, 920)
mkdir -p gpurun_out
L=$PWD/csplotch_b200
export CSPLOTCH_CLUSTER=1
for v in s1 s3 sc1 s1 s3 sc1; do
  export CSPLOTCH_LIB=$L/libcsplotch_b200_$v.so
  timeout -s KILL 300 python scripts/leap_bench.py c3 --waves 1 --steps 6 2>&1 | grep '^{' | sed "s/^{/{\"v\": \"$v\", /" | tee -a gpurun_out/r02_ab_call20.jsonl
done
export CSPLOTCH_LIB=$L/libcsplotch_b200_sc1.so
timeout -s KILL 300 python scripts/leap_bench.py c2 c4 --waves 1 --steps 6 2>&1 | grep '^{' | sed "s/^{/{\"v\": \"sc1\", /" | tee -a gpurun_out/r02_ab_call20.jsonl
for v in sc1 s3; do
  CSPLOTCH_LIB=$L/libcsplotch_b200_$v.so timeout -s KILL 600 python scripts/perf_probe.py c3 > gpurun_out/r02_ab20_probe_$v.log 2>&1
  echo "== probe $v"; grep -h '"iters_done"\|profile' gpurun_out/r02_ab20_probe_$v.log | cut -c1-330 | tail -3
done
