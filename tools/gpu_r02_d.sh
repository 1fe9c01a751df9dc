#!/bin/bash
# round-2 GPU call D: warp-uniform issue loops — parity over the tcgen05 test files, then timings
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_umma.py tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/d_tests.out 2>&1
echo "tests rc=$?" >> gpurun_out/d_tests.out
tail -n 8 gpurun_out/d_tests.out
for f in 1 0; do
  COPER_FUSED_SCORER=$f timeout 300 python tools/microbench.py bcebig bf16 > gpurun_out/d_mb_fused$f.out 2>&1
  COPER_FUSED_SCORER=$f timeout 300 python tools/microbench.py prof bf16 >> gpurun_out/d_mb_fused$f.out 2>&1
  cat gpurun_out/d_mb_fused$f.out
done
timeout 300 python tools/microbench.py prof tf32x3 > gpurun_out/d_mb.out 2>&1
timeout 300 python tools/microbench.py cpg >> gpurun_out/d_mb.out 2>&1
timeout 300 python tools/microbench.py rank >> gpurun_out/d_mb.out 2>&1
cat gpurun_out/d_mb.out
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --num-labels 0 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/d_launches.csv \
  python tools/microbench.py bcebig bf16 > gpurun_out/d_launches.out 2>&1
