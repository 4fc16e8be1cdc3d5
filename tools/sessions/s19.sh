#!/bin/bash
# GPU session 19 (round 2, 1 GPU): grid barrier with one release-reduction (K5, K9).
mkdir -p gpurun_out
{
for shape in "128 128 pairs" "128 256 pairs" "256 256 pairs" "512 512 persistent" "1024 1024 persistent"; do
  set -- $shape
  echo "== $1 x $2 $3"; LBM_VARIANTS=base,barrier1 timeout 300 python tools/build_variants.py --run --nx $1 --ny $2 --steps 40000 --reps 3 --kernel $3
done
} > gpurun_out/s19_bench.log 2>&1
LBM_B200_LIB=$PWD/advanced-hpc-lbm_b200/variants/liblbm_barrier1.so timeout 600 python -m pytest tests/test_gpu_pairs.py tests/test_gpu_parity.py -q --timeout 300 > gpurun_out/s19_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s19_tests.log
echo done
