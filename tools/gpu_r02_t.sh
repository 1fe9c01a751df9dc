#!/bin/bash
set -x
mkdir -p gpurun_out
B="--no-cpu-baseline --num-labels 0 --no-alt --no-breakdown --no-extra"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/t_launches_10m.csv \
  python bench.py --shape synth-10m --prec bf16 --steps 3 --warmup 3 $B > /dev/null 2> gpurun_out/t1.err
tail -n 3 gpurun_out/t1.err
