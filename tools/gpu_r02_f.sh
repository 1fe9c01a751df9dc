#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tests/diag_bf16_grad_error.py > gpurun_out/f_diag.out 2>&1
cat gpurun_out/f_diag.out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
tail -n 5 gpurun_out/f_bench.err
timeout 900 python -m pytest tests/test_golden.py tests/test_gpu_model.py -m gpu -q -x > gpurun_out/f_tests.out 2>&1
tail -n 15 gpurun_out/f_tests.out
