#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/i_multi.out 2>&1; tail -n 30 gpurun_out/i_multi.out
