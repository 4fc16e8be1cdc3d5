#!/bin/bash
# GPU session 15 (round 2, 1 GPU): K7 with L2 prefetch-ahead variants.
mkdir -p gpurun_out
{
echo "== 16384^2 tb2 L2 prefetch variants"; LBM_VARIANTS=base,tb2_pf2,tb2_pf4,tb2_pf8 timeout 900 python tools/build_variants.py --run --steps 200 --reps 3 --kernel tb2
echo "== 8192^2"; LBM_VARIANTS=base,tb2_pf4 timeout 600 python tools/build_variants.py --run --nx 8192 --ny 8192 --steps 200 --reps 3 --kernel tb2
} > gpurun_out/s15_bench.log 2>&1
echo done
