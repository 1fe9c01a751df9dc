#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_golden.py -m gpu -q -x -k "concat" > gpurun_out/r_concat.out 2>&1; tail -n 15 gpurun_out/r_concat.out
timeout 1200 python -m pytest tests/test_gpu_umma.py tests/test_gpu_fullsize.py tests/test_gpu_model.py -m gpu -q -x -k "fp16x3" > gpurun_out/r_fp16.out 2>&1; tail -n 8 gpurun_out/r_fp16.out
timeout 600 python bench.py --shape wn18rr --prec fp16x3 --steps 20 --warmup 5 --no-cpu-baseline --num-labels 0 --no-alt > gpurun_out/r_bench_wn.json 2> gpurun_out/r_bench.err; tail -n 3 gpurun_out/r_bench.err
timeout 300 python tools/microbench.py rank fp16x3 > gpurun_out/r_mb.out 2>&1
timeout 300 python tools/microbench.py prof fp16x3 >> gpurun_out/r_mb.out 2>&1
cat gpurun_out/r_mb.out
