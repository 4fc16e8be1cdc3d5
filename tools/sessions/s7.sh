#!/bin/bash
# GPU session 7 (round 2, 1 GPU): K7 barrier experiment, K1a ncu counters, bench launch list.
mkdir -p gpurun_out
{
echo "== 16384^2 tb2 variants"; LBM_VARIANTS=base,tb2_x_nobarrier timeout 600 python tools/build_variants.py --run --steps 100 --reps 3 --kernel tb2
echo "== 16384^2 vec4 base"; LBM_VARIANTS=base timeout 600 python tools/build_variants.py --run --steps 100 --reps 3 --kernel vec4
} > gpurun_out/s7_bench.log 2>&1
CMD="python tools/quick_bench.py --steps 6 --reps 1 --kernel vec4"
$CMD > gpurun_out/s7_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lbm_step_vec4 -s 3 -c 2 -o gpurun_out/prof_vec4_r02 $CMD > gpurun_out/s7_ncu_vec4.log 2>&1
BCMD="python bench.py --steps 2 --warmup 3 --timesteps 20 --no-cpu-baseline"
$BCMD > gpurun_out/s7_bench_short.json 2> gpurun_out/s7_bench_short.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s7_launches.csv $BCMD > gpurun_out/s7_ncu_launches.log 2>&1
CMD2="python tools/quick_bench.py --nx 128 --ny 128 --steps 2000 --reps 1 --kernel pairs"
$CMD2 > gpurun_out/s7_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lbm_steps_pairs -c 1 -o gpurun_out/prof_pairs_r02 $CMD2 > gpurun_out/s7_ncu_pairs.log 2>&1
echo done
