#!/bin/bash
# GPU session 17 (round 2, 1 GPU): K7 strip width variants across grid sizes.
mkdir -p gpurun_out
{
for shape in "1024 1024" "1536 1536" "2048 2048" "3072 3072" "8192 8192" "16384 2048" "16384 16384"; do
  set -- $shape
  echo "== $1 x $2"; LBM_VARIANTS=base,tb2_t96,tb2_t64 timeout 900 python tools/build_variants.py --run --nx $1 --ny $2 --steps 400 --reps 3 --kernel tb2
done
} > gpurun_out/s17_bench.log 2>&1
echo done
