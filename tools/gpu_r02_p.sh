#!/bin/bash
# round-2 profile call: launch lists (train + eval, both headline workloads) and --set full captures of the top kernels
set -x
mkdir -p gpurun_out
B="--no-cpu-baseline --num-labels 0 --no-alt --no-breakdown --no-extra"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/p_launches_wn18rr_fp16x3.csv \
  python bench.py --shape wn18rr --prec fp16x3 --steps 3 --warmup 3 $B > /dev/null 2> gpurun_out/p1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/p_launches_bcebig_bf16.csv \
  python tools/microbench.py bcebig bf16 > /dev/null 2> gpurun_out/p2.err
NB="--kernel-name-base demangled --set full --clock-control none --import-source on"
timeout 900 ncu $NB -k regex:bce_dq_fused -s 1 -c 1 -o gpurun_out/p_fused python tools/microbench.py bcebig bf16 > gpurun_out/p3.out 2>&1
timeout 900 ncu $NB -k regex:"GemmCfg<1, 256, 4, 4, 0, 1, 0, 0>.*StoreEpi" -s 1 -c 1 -o gpurun_out/p_dE python tools/microbench.py bcebig bf16 > gpurun_out/p4.out 2>&1
timeout 900 ncu $NB -k regex:RankEpiT -s 3 -c 1 -o gpurun_out/p_rank python tools/microbench.py rank bf16 > gpurun_out/p5.out 2>&1
timeout 900 ncu $NB -k regex:BceEpiT -s 1 -c 1 -o gpurun_out/p_bce_fp16x3 python tools/microbench.py prof fp16x3 > gpurun_out/p6.out 2>&1
timeout 900 ncu $NB -k regex:"StoreEpi" -s 2 -c 2 -o gpurun_out/p_dqdE_fp16x3 python tools/microbench.py prof fp16x3 > gpurun_out/p7.out 2>&1
timeout 900 ncu $NB -k regex:"CpgFwdEpi|CpgBwdTEpi" -s 2 -c 2 -o gpurun_out/p_cpg_fp16x3 python tools/microbench.py cpg > gpurun_out/p8.out 2>&1
ls -la gpurun_out/*.ncu-rep
