#!/bin/bash
# GPU session 5 (round 2, 1 GPU): K9 (pairs kernel) correctness + timings, full suite.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pairs.py -q --timeout 300 -x > gpurun_out/s5_tests_pairs.log 2>&1
echo "rc=$?" >> gpurun_out/s5_tests_pairs.log
{
for shape in "128 128" "128 256" "256 256"; do
  set -- $shape
  for k in persistent pairs; do
    echo "== $1x$2 $k"; timeout 200 python tools/quick_bench.py --nx $1 --ny $2 --steps 40000 --reps 3 --kernel $k | grep MLUPS | tail -1
  done
  for tr in 1 2 3 4; do echo "== $1x$2 pairs tile_rows $tr"; LBM_PAIRS_TILE_ROWS=$tr timeout 100 python tools/quick_bench.py --nx $1 --ny $2 --steps 40000 --reps 3 --kernel pairs | grep MLUPS | tail -1; done
done
} > gpurun_out/s5_bench.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/s5_tests_all.log 2>&1
echo "rc=$?" >> gpurun_out/s5_tests_all.log
echo done
