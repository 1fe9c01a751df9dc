#!/bin/bash
set -x
mkdir -p gpurun_out
NB="--kernel-name-base demangled --set full --clock-control none --import-source on"
timeout 900 ncu $NB -k "regex:RankEpiT<.int.2>" -s 2 -c 1 -o gpurun_out/v_rank1m python tools/microbench.py rank bf16 > gpurun_out/v1.out 2>&1
tail -n 3 gpurun_out/v1.out
