#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py tests/test_golden.py -m gpu -q -x > gpurun_out/x_tests.out 2>&1; tail -n 8 gpurun_out/x_tests.out
timeout 600 python bench.py --shape synth-10m --prec bf16 --steps 8 --warmup 3 --no-cpu-baseline --num-labels 0 --no-alt --no-breakdown > gpurun_out/x_bench_10m.json 2> gpurun_out/x_bench.err; tail -n 3 gpurun_out/x_bench.err
timeout 600 python bench.py --shape wn18rr --prec fp16x3 --steps 20 --warmup 5 --no-cpu-baseline --num-labels 0 --no-alt --no-breakdown > gpurun_out/x_bench_wn.json 2>> gpurun_out/x_bench.err; tail -n 3 gpurun_out/x_bench.err
timeout 600 python bench.py --shape fb15k-237 --prec fp16x3 --steps 20 --warmup 5 --no-cpu-baseline --num-labels 0 --no-alt --no-breakdown > gpurun_out/x_bench_fb.json 2>> gpurun_out/x_bench.err; tail -n 3 gpurun_out/x_bench.err
