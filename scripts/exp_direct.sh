#!/bin/bash
# one development library, two runs: staged pipeline (CSPLOTCH_FORCE_ORDER=nodirect) vs eval_direct, + GPU parity
mkdir -p gpurun_out
export CSPLOTCH_LIB=$PWD/csplotch_b200/libcsplotch_b200_dev.so
(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -40) > gpurun_out/direct_pytest.log
grep -v "development build" gpurun_out/direct_pytest.log | grep -i "assert\|error\|FAILED" | grep -v "csplotch_b2\.\.\." | head -10
tail -n 1 gpurun_out/direct_pytest.log
export PROBE_G=1184
CSPLOTCH_FORCE_ORDER=nodirect timeout 200 python scripts/perf_probe.py c2 2>&1 | grep -h 'iters_done\|profile' | tee gpurun_out/direct_A.log
timeout 200 python scripts/perf_probe.py c2 2>&1 | grep -h 'iters_done\|profile' | tee gpurun_out/direct_B.log
