#!/bin/bash
# GPU session 11 (round 2, 1 GPU): K9 with depth > 2: correctness and depth sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pairs.py tests/test_gpu_parity.py -q --timeout 600 -x > gpurun_out/s11_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s11_tests.log
{
for shape in "128 128" "128 256" "256 256"; do
  set -- $shape
  for dp in 2 3 4 5 6; do
    for tr in 1 2; do
      echo "== $1x$2 pairs depth $dp tile_rows $tr"; LBM_PAIRS_DEPTH=$dp LBM_PAIRS_TILE_ROWS=$tr timeout 100 python tools/quick_bench.py --nx $1 --ny $2 --steps 40000 --reps 3 --kernel pairs 2>&1 | grep -E "MLUPS|Error" | tail -1
    done
  done
  echo "== $1x$2 pairs default"; timeout 100 python tools/quick_bench.py --nx $1 --ny $2 --steps 40000 --reps 3 --kernel pairs | grep MLUPS | tail -1
done
} > gpurun_out/s11_bench.log 2>&1
echo done
