#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_golden.py -m gpu -q > gpurun_out/h_golden.out 2>&1; tail -n 8 gpurun_out/h_golden.out
for s in wn18rr fb15k-237; do for p in fp16x3 tf32x3; do
timeout 600 python bench.py --shape $s --prec $p --steps 20 --warmup 5 --no-cpu-baseline --num-labels 0 --no-alt > gpurun_out/h_bench_${s}_$p.json 2> gpurun_out/h_bench.err
tail -n 3 gpurun_out/h_bench.err
done; done
