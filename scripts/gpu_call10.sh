#!/bin/bash
# round 2, GPU call 10: which change after the cluster commit slowed the fused leapfrog? head (a5e9c90) / dev (working tree) / devp (dev + neighbour
# positions fetched before the ring barrier) / base (pre-cluster) / the product library of call 6
mkdir -p gpurun_out
L=$PWD/csplotch_b200
export CSPLOTCH_CLUSTER=1
for r in 1 2; do
for v in base head dev devp prod; do
  if [ $v = prod ]; then unset CSPLOTCH_LIB; else export CSPLOTCH_LIB=$L/libcsplotch_b200_$v.so; fi
  timeout -s KILL 300 python scripts/leap_bench.py c2 c3 --waves 1 2>&1 | grep '^{' | sed "s/^{/{\"v\": \"$v\", /" | tee -a gpurun_out/r02_ab_call10.jsonl
done
done
