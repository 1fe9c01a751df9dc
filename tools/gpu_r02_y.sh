#!/bin/bash
set -x
mkdir -p gpurun_out
B="--no-cpu-baseline --num-labels 0 --no-alt --no-breakdown --no-extra"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/y_launches_fb.csv \
  python bench.py --shape fb15k-237 --prec fp16x3 --steps 3 --warmup 3 $B > /dev/null 2> gpurun_out/y1.err
