#!/bin/bash
# GPU session 3 (round 2, 2 GPUs): multi-GPU tests (flags, events, IPC, time-out), bench N=2.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/s3_topo.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_tb2.py -q --timeout 900 > gpurun_out/s3_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s3_tests.log
LBM_BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/s3_bench_n2.json 2> gpurun_out/s3_bench_n2.err
echo "rc=$?" >> gpurun_out/s3_bench_n2.err
timeout 600 python tools/quick_bench.py --gpus 2 --steps 100 --reps 3 --kernel tb2 --ny 32768 > gpurun_out/s3_inproc_n2.log 2>&1
echo done
