#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/q4_tests.out 2>&1; tail -n 12 gpurun_out/q4_tests.out
( time timeout 1200 python bench.py ) > gpurun_out/q4_bench_n1.json 2> gpurun_out/q4_bench.err; tail -n 6 gpurun_out/q4_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/q4_smoke.out 2>&1; tail -n 6 gpurun_out/q4_smoke.out
