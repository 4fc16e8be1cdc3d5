#!/bin/bash
# GPU session 20 (round 2, 1 GPU): K5 / K7 crossover between 1024^2 and 2048^2.
mkdir -p gpurun_out
{
for n in 1152 1280 1408 1536 1792; do
  for k in persistent tb2 vec4; do
    echo "== ${n}^2 $k"; timeout 100 python tools/quick_bench.py --nx $n --ny $n --steps 4000 --reps 3 --kernel $k | grep MLUPS | tail -1
  done
done
} > gpurun_out/s20_bench.log 2>&1
echo done
