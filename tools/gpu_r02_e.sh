#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/fused_trace.py > gpurun_out/e_trace.out 2>&1
cat gpurun_out/e_trace.out
timeout 300 python tools/microbench.py bcebig bf16 > gpurun_out/e_mb.out 2>&1
timeout 300 python tools/microbench.py prof bf16 >> gpurun_out/e_mb.out 2>&1
cat gpurun_out/e_mb.out
timeout 300 python tests/diag_bf16_grad_error.py > gpurun_out/e_diag.out 2>&1
cat gpurun_out/e_diag.out
