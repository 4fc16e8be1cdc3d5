#!/bin/bash
# GPU session 6 (round 2, 8 GPUs): the driver's scaling commands at N = 8 and N = 4, e2e breakdown per rank.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/s6_topo.txt 2>&1
nproc >> gpurun_out/s6_topo.txt
run() {  # n, tag, extra args...
  n=$1; tag=$2; shift 2
  LBM_BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n "$@" > gpurun_out/s6_bench_$tag.json 2> gpurun_out/s6_bench_$tag.err
  echo "rc=$?" >> gpurun_out/s6_bench_$tag.err
}
run 8 n8
LBM_BENCH_BIND=0 run 8 n8_nobind --no-parity --no-strong --steps 3
run 4 n4
timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 > gpurun_out/s6_tests_multi.log 2>&1
echo "rc=$?" >> gpurun_out/s6_tests_multi.log
echo done
