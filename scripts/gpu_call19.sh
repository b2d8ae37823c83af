#!/bin/bash
# round 2, GPU call 19 (2 GPUs): the sharded (c5, strong scaling) bench path end to end at reduced size, the way the driver launches it
mkdir -p gpurun_out
nvidia-smi -L
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --genes 148 --ess-genes 4 --ess-n 30 --e2e-genes 16 > gpurun_out/r02_bench_c5_n2_small.json 2> gpurun_out/r02_bench_c5_n2_small.err
echo "rc=$?"; tail -6 gpurun_out/r02_bench_c5_n2_small.err; cut -c1-1500 gpurun_out/r02_bench_c5_n2_small.json
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_n2.json 2> gpurun_out/r02_bench_ref_n2.err
echo "rc=$?"; cut -c1-400 gpurun_out/r02_bench_ref_n2.json
