#!/bin/bash
# GPU session 2 (round 2): full GPU test suite, the new bench line, K7/K5 crossover, ncu of K7.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/s2_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s2_tests.log
timeout 900 python bench.py > gpurun_out/s2_bench_n1.json 2> gpurun_out/s2_bench_n1.err
echo "rc=$?" >> gpurun_out/s2_bench_n1.err
{
for n in 1024 1536 2048 4096 8192; do
  for k in persistent tb2 vec4; do
    if [ $k = persistent ] && [ $n -gt 2048 ]; then continue; fi
    echo "== ${n}^2 $k"; timeout 200 python tools/quick_bench.py --nx $n --ny $n --steps 400 --reps 3 --kernel $k | grep MLUPS | tail -1
  done
done
echo "== 16384^2 tb2"; timeout 200 python tools/quick_bench.py --steps 100 --reps 3 --kernel tb2 | grep MLUPS
echo "== 16384^2 vec4"; timeout 200 python tools/quick_bench.py --steps 100 --reps 3 --kernel vec4 | grep MLUPS
} > gpurun_out/s2_sizes.log 2>&1
CMD="python tools/quick_bench.py --steps 6 --reps 1 --kernel tb2"
$CMD > gpurun_out/s2_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lbm_step2_tb -s 3 -c 2 -o gpurun_out/prof_tb2 $CMD > gpurun_out/s2_ncu_full.log 2>&1
BCMD="python bench.py --steps 2 --warmup 3 --timesteps 20 --no-cpu-baseline --no-shipped"
$BCMD > gpurun_out/s2_bench_short.json 2> gpurun_out/s2_bench_short.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/s2_launches.csv $BCMD > gpurun_out/s2_ncu_launches.log 2>&1
echo done
