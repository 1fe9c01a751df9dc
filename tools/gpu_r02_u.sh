#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/u_tests.out 2>&1; tail -n 8 gpurun_out/u_tests.out
timeout 600 python bench.py --shape synth-10m --prec bf16 --steps 8 --warmup 3 --no-cpu-baseline --num-labels 0 --no-alt > gpurun_out/u_bench_10m.json 2> gpurun_out/u_bench.err; tail -n 3 gpurun_out/u_bench.err
