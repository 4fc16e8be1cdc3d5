#!/bin/bash
# GPU session 13 (round 2, 1 GPU): L2 persisting access-policy window for K5 (measurement).
mkdir -p gpurun_out
{
for n in 512 768 1024; do
  echo "== ${n}^2 persistent, no window"; timeout 100 python tools/quick_bench.py --nx $n --ny $n --steps 20000 --reps 3 --kernel persistent | grep MLUPS | tail -1
  echo "== ${n}^2 persistent, persisting window"; LBM_GPU_L2_PERSIST=1 LBM_GPU_L2_PERSIST_VERBOSE=1 timeout 100 python tools/quick_bench.py --nx $n --ny $n --steps 20000 --reps 3 --kernel persistent 2>&1 | grep -E "MLUPS|L2 persistence" | tail -2
done
} > gpurun_out/s13_l2persist.log 2>&1
echo done
