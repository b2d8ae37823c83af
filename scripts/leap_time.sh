#!/bin/bash
# kernel durations of the stand-alone fused leapfrog (k_leapfrog_fixed) for several batch sizes: ncu duration metric only
# usage: scripts/leap_time.sh <shape> <B1> [B2 ...]
shape=$1; shift
mkdir -p gpurun_out
for B in "$@"; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_leapfrog_fixed --csv --log-file /tmp/lt_$B.csv python scripts/profile_lpgrad.py $shape $B 3 leap > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('/tmp/lt_$B.csv')) if len(r)>5 and r[0].isdigit()]
d=[float(r[-1].replace(',','')) for r in rows]
print('$shape B=$B launches', len(d), 'duration(last)', d[-1], 'per-traj-eval', d[-1]/($B*5))
PY
done
