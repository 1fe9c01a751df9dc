#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_fullsize.py tests/test_golden.py -m gpu -q -x -k "fp16x3" > gpurun_out/j_tests.out 2>&1; tail -n 8 gpurun_out/j_tests.out
timeout 600 python bench.py --shape wn18rr --prec fp16x3 --steps 20 --warmup 5 --no-cpu-baseline --num-labels 0 --no-alt > gpurun_out/j_bench_fp16x3.json 2> gpurun_out/j_bench.err; tail -n 3 gpurun_out/j_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file gpurun_out/j_launches_fp16x3.csv \
  python bench.py --shape wn18rr --prec fp16x3 --steps 4 --warmup 3 --no-cpu-baseline --num-labels 0 --no-alt --no-breakdown > /dev/null 2> gpurun_out/j_ncu.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:RankEpiT -s 2 -c 1 -o gpurun_out/j_rank \
  python tools/microbench.py rank bf16 > gpurun_out/j_rank.out 2>&1
tail -n 3 gpurun_out/j_rank.out
