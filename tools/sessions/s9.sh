#!/bin/bash
# GPU session 9 (round 2, 1 GPU): final single-GPU evidence: full suite, bench line, ncu of K7, launch list.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/s9_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s9_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s9_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/s9_bench_n1.json 2> gpurun_out/s9_bench_n1.err
echo "rc=$?" >> gpurun_out/s9_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s9_bench_ref_n1.json 2> gpurun_out/s9_bench_ref_n1.err
{
for n in 2048 4096 8192 16384; do echo "== ${n}^2 tb2"; timeout 200 python tools/quick_bench.py --nx $n --ny $n --steps 200 --reps 3 --kernel tb2 | grep MLUPS | tail -1; done
echo "== 16384x2048 tb2 (strong-scaling slab)"; timeout 200 python tools/quick_bench.py --nx 16384 --ny 2048 --steps 400 --reps 3 --kernel tb2 | grep MLUPS | tail -1
} > gpurun_out/s9_sizes.log 2>&1
CMD="python tools/quick_bench.py --steps 6 --reps 1 --kernel tb2"
$CMD > gpurun_out/s9_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lbm_step2_tb -s 3 -c 2 -o gpurun_out/prof_tb2_final $CMD > gpurun_out/s9_ncu_full.log 2>&1
BCMD="python bench.py --steps 2 --warmup 3 --timesteps 20 --no-cpu-baseline"
$BCMD > gpurun_out/s9_bench_short.json 2> gpurun_out/s9_bench_short.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s9_launches.csv $BCMD > gpurun_out/s9_ncu_launches.log 2>&1
echo done
