#!/bin/bash
# round-2 GPU call B: fused scorer kernel — toy parity first, then the full BCE test matrix, then timings on / off
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_umma.py -m gpu -q -x -k "bce and bf16 and 7-97-40" > gpurun_out/b_toy.out 2>&1
echo "toy rc=$?" >> gpurun_out/b_toy.out
tail -n 30 gpurun_out/b_toy.out
timeout 900 python -m pytest tests/test_gpu_umma.py -m gpu -q -k "bce" > gpurun_out/b_bce.out 2>&1
echo "bce rc=$?" >> gpurun_out/b_bce.out
tail -n 30 gpurun_out/b_bce.out
for f in 1 0; do
  COPER_FUSED_SCORER=$f timeout 300 python tools/microbench.py bcebig bf16 > gpurun_out/b_mb_fused$f.out 2>&1
  COPER_FUSED_SCORER=$f timeout 300 python tools/microbench.py prof bf16 >> gpurun_out/b_mb_fused$f.out 2>&1
  COPER_FUSED_SCORER=$f timeout 300 python tools/microbench.py prof tf32x3 >> gpurun_out/b_mb_fused$f.out 2>&1
  cat gpurun_out/b_mb_fused$f.out
done
