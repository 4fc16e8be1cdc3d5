#!/bin/bash
# GPU session 16 (round 2, 1 GPU): K7 strip width variants (128 / 96 / 64 threads per block).
mkdir -p gpurun_out
{
echo "== 16384^2 tb2 strip variants"; LBM_VARIANTS=base,tb2_t96,tb2_t64 timeout 900 python tools/build_variants.py --run --steps 200 --reps 3 --kernel tb2
echo "== 4096^2"; LBM_VARIANTS=base,tb2_t96,tb2_t64 timeout 600 python tools/build_variants.py --run --nx 4096 --ny 4096 --steps 400 --reps 3 --kernel tb2
echo "== 2048^2"; LBM_VARIANTS=base,tb2_t96,tb2_t64 timeout 600 python tools/build_variants.py --run --nx 2048 --ny 2048 --steps 400 --reps 3 --kernel tb2
} > gpurun_out/s16_bench.log 2>&1
echo done
