#!/bin/bash
# GPU session 10 (round 2, 8 GPUs): the driver's scaling sequence N = 1, 2, 4, 8 back to back on one box, final code.
mkdir -p gpurun_out
run() {  # n, tag, extra args...
  n=$1; tag=$2; shift 2
  if [ $n = 1 ]; then
    LBM_BENCH_VERBOSE=1 timeout 900 python bench.py "$@" > gpurun_out/s10_scale_$tag.json 2> gpurun_out/s10_scale_$tag.err
  else
    LBM_BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $n "$@" > gpurun_out/s10_scale_$tag.json 2> gpurun_out/s10_scale_$tag.err
  fi
  echo "rc=$?" >> gpurun_out/s10_scale_$tag.err
}
run 1 n1 --no-cpu-baseline
run 2 n2
run 4 n4
run 8 n8
timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 > gpurun_out/s10_tests_multi.log 2>&1
echo "rc=$?" >> gpurun_out/s10_tests_multi.log
echo done
