#!/bin/bash
# round-2 GPU call C: launch list + full ncu capture of the fused scorer kernel at the 1.25M-row bf16 shard
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c_launches.csv \
  python tools/microbench.py bcebig bf16 > gpurun_out/c_launches.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bce_dq_fused -s 1 -c 1 -o gpurun_out/c_fused \
  python tools/microbench.py bcebig bf16 > gpurun_out/c_full.out 2>&1
tail -n 5 gpurun_out/c_full.out
