#!/bin/bash
# GPU session 1 (round 2): correctness of the diet + K7 + protocol refactor, first timings.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/s1_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_tb2.py tests/test_gpu_parity.py tests/test_gpu_multi.py -q -x --timeout 600 > gpurun_out/s1_tests_a.log 2>&1
echo "rc=$?" >> gpurun_out/s1_tests_a.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_gpu_tb2.py --deselect tests/test_gpu_parity.py --deselect tests/test_gpu_multi.py > gpurun_out/s1_tests_b.log 2>&1
echo "rc=$?" >> gpurun_out/s1_tests_b.log
{
echo "== 16384^2 vec4 variants"; timeout 600 python tools/build_variants.py --run --steps 40 --reps 3 --kernel vec4
echo "== 16384^2 tb2 variants"; LBM_VARIANTS=base,tb2_mb2,scalar timeout 600 python tools/build_variants.py --run --steps 40 --reps 3 --kernel tb2
for h in 16 32 128 256; do echo "== tb2 seg_rows $h"; LBM_TB2_SEG_ROWS=$h timeout 120 python tools/quick_bench.py --steps 40 --reps 3 --kernel tb2 | grep MLUPS | tail -1; done
for n in 128 256 1024; do
  echo "== ${n}^2 persistent variants"; LBM_VARIANTS=base,mb6,mb5,mb4,scalar timeout 600 python tools/build_variants.py --run --nx $n --ny $n --steps 20000 --reps 3 --kernel persistent
done
echo "== 128^2 cluster"; timeout 120 python tools/quick_bench.py --nx 128 --ny 128 --steps 20000 --reps 3 --kernel cluster | grep MLUPS
echo "== 128x256 cluster"; timeout 120 python tools/quick_bench.py --nx 128 --ny 256 --steps 20000 --reps 3 --kernel cluster | grep MLUPS
echo "== 4096^2 tb2 / vec4"; timeout 120 python tools/quick_bench.py --nx 4096 --ny 4096 --steps 200 --reps 3 --kernel tb2 | grep MLUPS | tail -1; timeout 120 python tools/quick_bench.py --nx 4096 --ny 4096 --steps 200 --reps 3 --kernel vec4 | grep MLUPS | tail -1
} > gpurun_out/s1_bench.log 2>&1
echo done
