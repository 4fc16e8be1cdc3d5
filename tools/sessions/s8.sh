#!/bin/bash
# GPU session 8 (round 2, 1 GPU): K7 with the split barrier: correctness + timing.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tb2.py tests/test_gpu_multi.py tests/test_gpu_fullsize.py -q --timeout 600 > gpurun_out/s8_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s8_tests.log
{
for n in 2048 4096 8192 16384; do echo "== ${n}^2 tb2"; timeout 200 python tools/quick_bench.py --nx $n --ny $n --steps 200 --reps 3 --kernel tb2 | grep MLUPS | tail -1; done
for h in 32 128; do echo "== tb2 seg_rows $h"; LBM_TB2_SEG_ROWS=$h timeout 120 python tools/quick_bench.py --steps 100 --reps 3 --kernel tb2 | grep MLUPS | tail -1; done
} > gpurun_out/s8_bench.log 2>&1
echo done
