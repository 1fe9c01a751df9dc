#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_umma.py tests/test_gpu_fullsize.py tests/test_golden.py -m gpu -q -x -k "rank or full_size or eval" > gpurun_out/w_tests.out 2>&1; tail -n 8 gpurun_out/w_tests.out
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "eval" >> gpurun_out/w_tests.out 2>&1; tail -n 4 gpurun_out/w_tests.out
timeout 300 python tools/microbench.py rank > gpurun_out/w_mb.out 2>&1; cat gpurun_out/w_mb.out
