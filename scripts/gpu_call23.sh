#!/bin/bash
# round 2, GPU call 23 (final product library): `ncu --set full` of the fused leapfrog at c3, per source line + raw page
mkdir -p gpurun_out
export CSPLOTCH_CLUSTER=1
timeout -s KILL 200 python scripts/leap_bench.py c3 --waves 1 --steps 2 > gpurun_out/plain_leap_c3.log 2>&1 || { tail -5 gpurun_out/plain_leap_c3.log; exit 1; }
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_leapfrog_fixed -s 1 -c 1 -o /tmp/prof_leap_c3 python scripts/leap_bench.py c3 --waves 1 --steps 2 > gpurun_out/ncu_leap_c3.log 2>&1
python scripts/ncu_lines.py /tmp/prof_leap_c3.ncu-rep 70 > gpurun_out/r02_ncu_leapfrog_c3_lines.txt 2>&1
ncu -i /tmp/prof_leap_c3.ncu-rep --page raw --csv > gpurun_out/r02_ncu_leapfrog_c3_raw.csv 2>/dev/null
head -12 gpurun_out/r02_ncu_leapfrog_c3_lines.txt
