#!/bin/bash
# GPU session 14 (round 2, 4 GPUs): full-size equality 1 GPU vs 4 GPUs with the two-step kernel.
mkdir -p gpurun_out
timeout 900 python tools/validate_weak_scaling.py > gpurun_out/s14_weak_validation.log 2>&1
echo "rc=$?" >> gpurun_out/s14_weak_validation.log
echo done
