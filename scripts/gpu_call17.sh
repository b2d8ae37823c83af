#!/bin/bash
# round 2, GPU call 17: DRAM traffic of a sampler launch at normal depth (plain CTAs, c3 and c2); the non-compositional kernels at
# 2 CTAs/SM x 128 registers (devs2) against 3 x 80 (dev); the reference arm with the bounded per-step sample
mkdir -p gpurun_out
t0=$SECONDS
bash scripts/traffic_quick.sh c3 74 8; bash scripts/traffic_quick.sh c2 1184 20
echo "== traffic: $((SECONDS-t0)) s"; t0=$SECONDS
L=$PWD/csplotch_b200
for v in dev devs2 dev devs2; do
  export CSPLOTCH_LIB=$L/libcsplotch_b200_$v.so
  per=3; [ $v = devs2 ] && per=2
  CSP_CTAS_PER_SM=$per timeout -s KILL 300 python scripts/leap_bench.py c2 c4 --waves 1 --steps 6 2>&1 | grep '^{' | sed "s/^{/{\"v\": \"$v\", /" | tee -a gpurun_out/r02_ab_call17.jsonl
  PROBE_G=1184 timeout -s KILL 600 python scripts/perf_probe.py c2 > gpurun_out/r02_ab17_probe_$v.log 2>&1
  echo "== probe $v"; grep -h '"iters_done"\|profile' gpurun_out/r02_ab17_probe_$v.log | cut -c1-330
done
unset CSPLOTCH_LIB
echo "== c2 A/B: $((SECONDS-t0)) s"; t0=$SECONDS
timeout -s KILL 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref_c3.json 2> gpurun_out/r02_bench_ref_c3.err; cut -c1-900 gpurun_out/r02_bench_ref_c3.json
echo "== reference arm: $((SECONDS-t0)) s"
