#!/bin/bash
# GPU session 21 (round 2, 1 GPU): K7 with short segments at the end of a launch.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tb2.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py -q --timeout 600 > gpurun_out/s21_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s21_tests.log
{
for tr in 0 8 16 32; do echo "== 16384^2 tail rows $tr"; LBM_TB2_TAIL_ROWS=$tr timeout 200 python tools/quick_bench.py --steps 200 --reps 3 --kernel tb2 | grep MLUPS | tail -1; done
for tr in 0 16; do echo "== 8192^2 tail rows $tr"; LBM_TB2_TAIL_ROWS=$tr timeout 200 python tools/quick_bench.py --nx 8192 --ny 8192 --steps 400 --reps 3 --kernel tb2 | grep MLUPS | tail -1; done
for tr in 0 16; do echo "== 16384x2048 tail rows $tr"; LBM_TB2_TAIL_ROWS=$tr timeout 200 python tools/quick_bench.py --nx 16384 --ny 2048 --steps 400 --reps 3 --kernel tb2 | grep MLUPS | tail -1; done
} > gpurun_out/s21_bench.log 2>&1
echo done
