#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py -m gpu -q -x -k "fp16x3" > gpurun_out/g_umma.out 2>&1; tail -n 25 gpurun_out/g_umma.out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "cpg" > gpurun_out/g_cpg.out 2>&1; tail -n 25 gpurun_out/g_cpg.out
timeout 900 python -m pytest tests/test_golden.py -m gpu -q > gpurun_out/g_golden.out 2>&1; tail -n 25 gpurun_out/g_golden.out
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q > gpurun_out/g_model.out 2>&1; tail -n 25 gpurun_out/g_model.out
timeout 300 python tools/microbench.py prof fp16x3 > gpurun_out/g_mb.out 2>&1
timeout 300 python tools/microbench.py prof tf32x3 >> gpurun_out/g_mb.out 2>&1
timeout 300 python tools/microbench.py cpg >> gpurun_out/g_mb.out 2>&1
timeout 300 python tools/microbench.py rank >> gpurun_out/g_mb.out 2>&1
cat gpurun_out/g_mb.out
timeout 600 python bench.py --shape wn18rr --prec fp16x3 --steps 20 --warmup 5 --no-cpu-baseline --num-labels 0 --no-alt > gpurun_out/g_bench_fp16x3.json 2> gpurun_out/g_bench.err
tail -n 3 gpurun_out/g_bench.err
