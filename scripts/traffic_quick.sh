#!/bin/bash
# usage: scripts/traffic_quick.sh <shape> <genes> [warm transitions]   (GPU box)
# DRAM bytes of ONE sampler launch at normal tree depth with one-pass ncu metrics (no --set full: the kernel runs for
# seconds and owns tens of GB, a full capture replays it ~45 times) -> profiles/traffic_<shape>.json
shape=$1; G=$2; warm=${3:-8}
mkdir -p gpurun_out
export CSPLOTCH_CLUSTER=1      # the scope of the bench's resident arm (thousands of chains: plain CTAs)
timeout 600 python scripts/profile_nuts.py $shape $G 1 $warm > gpurun_out/plain_tq_$shape.log 2>&1 || { tail -3 gpurun_out/plain_tq_$shape.log; exit 1; }
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
  --clock-control none -k regex:k_nuts_advance -s 1 -c 1 --csv --page raw --log-file gpurun_out/raw_tq_$shape.csv python scripts/profile_nuts.py $shape $G 1 $warm > gpurun_out/ncu_tq_$shape.log 2>&1
python scripts/traffic_from_ncu.py tq_$shape $shape && cp profiles/traffic_$shape.json gpurun_out/traffic_$shape.json
