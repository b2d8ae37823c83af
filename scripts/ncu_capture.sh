#!/bin/bash
# usage: scripts/ncu_capture.sh <tag> <kernel regex> <python driver args...>
# Profiles one launch with ncu --set full on the GPU box and keeps only text summaries under gpurun_out/
# (the .ncu-rep of this kernel is ~30 MB; gpurun_out/ is capped at 64 MiB).
set -u
tag=$1; shift
kre=$1; shift
mkdir -p gpurun_out
timeout 300 python "$@" > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -n 5 gpurun_out/plain_$tag.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$kre -s ${NCU_SKIP:-1} -c 1 -o /tmp/prof_$tag python "$@" > gpurun_out/ncu_$tag.log 2>&1
python scripts/ncu_lines.py /tmp/prof_$tag.ncu-rep 60 > gpurun_out/lines_$tag.txt 2>&1
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/raw_$tag.csv 2>/dev/null
tail -n 2 gpurun_out/plain_$tag.log
