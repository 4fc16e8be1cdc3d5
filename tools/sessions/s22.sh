#!/bin/bash
# GPU session 22 (round 2, 8 GPUs): the N = 8 bench line with the final code.
mkdir -p gpurun_out
LBM_BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 > gpurun_out/s22_scale_n8.json 2> gpurun_out/s22_scale_n8.err
echo "rc=$?" >> gpurun_out/s22_scale_n8.err
echo done
