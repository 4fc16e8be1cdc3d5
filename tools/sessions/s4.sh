#!/bin/bash
# GPU session 4 (round 2, 1 GPU): K8 (persistent two-step kernel) correctness and timings.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tb2.py tests/test_gpu_parity.py -q --timeout 900 -x > gpurun_out/s4_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s4_tests.log
{
for shape in "128 128" "128 256" "256 256" "512 512" "1024 1024"; do
  set -- $shape
  for k in persistent tb2p cluster; do
    if [ $k = cluster ] && [ $1 -gt 128 ]; then continue; fi
    echo "== $1x$2 $k"; timeout 200 python tools/quick_bench.py --nx $1 --ny $2 --steps 20000 --reps 3 --kernel $k | grep MLUPS | tail -1
  done
done
for sr in 1 2 4; do echo "== 128x128 tb2p seg $sr"; LBM_TB2P_SEG_ROWS=$sr timeout 100 python tools/quick_bench.py --nx 128 --ny 128 --steps 20000 --reps 3 --kernel tb2p | grep MLUPS | tail -1; done
for sr in 1 2 4; do echo "== 256x256 tb2p seg $sr"; LBM_TB2P_SEG_ROWS=$sr timeout 100 python tools/quick_bench.py --nx 256 --ny 256 --steps 20000 --reps 3 --kernel tb2p | grep MLUPS | tail -1; done
for sr in 4 7 12 16; do echo "== 1024x1024 tb2p seg $sr"; LBM_TB2P_SEG_ROWS=$sr timeout 100 python tools/quick_bench.py --nx 1024 --ny 1024 --steps 20000 --reps 3 --kernel tb2p | grep MLUPS | tail -1; done
} > gpurun_out/s4_bench.log 2>&1
echo done
