#!/bin/bash
# GPU session 12 (round 2, 1 GPU): final regression of the committed tree: full suite, smoke, default bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/s12_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s12_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s12_smoke.log 2>&1
( time timeout 900 python bench.py > gpurun_out/s12_bench_n1.json 2> gpurun_out/s12_bench_n1.err ) 2> gpurun_out/s12_bench_time.txt
echo done
