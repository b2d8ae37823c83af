#!/bin/bash
# round 2, GPU call 24: the e2e job (74 c3 genes x 4 chains, 20 transitions from initialisation) with the cluster size the library picks
# (4 CTAs per chain at 296 gene-chains) against plain CTAs and 2-CTA clusters
mkdir -p gpurun_out
for cl in auto 1 2; do
  if [ $cl = auto ]; then unset CSPLOTCH_CLUSTER; else export CSPLOTCH_CLUSTER=$cl; fi
  timeout -s KILL 170 python bench.py --genes 74 --steps 1 --warmup 3 --ess-genes 0 --no-secondary --no-cpu > gpurun_out/r02_e2e_cl_$cl.json 2> gpurun_out/r02_e2e_cl_$cl.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_e2e_cl_$cl.json')); print('cl=$cl', 'e2e', round(d['e2e']['value']), 'value', round(d['value']), 'tail', round(d['roofline']['launch_tail_frac'],3))"
done
