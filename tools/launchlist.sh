#!/bin/bash
# usage (on the GPU box): tools/launchlist.sh <tag> [bench args...]  -> gpurun_out/launches_<tag>.csv + summary on stdout
tag=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-breakdown "$@" > gpurun_out/ncu_bench_$tag.log 2>&1
python tools/ncu_summary.py gpurun_out/launches_$tag.csv 50 | cut -c1-150
