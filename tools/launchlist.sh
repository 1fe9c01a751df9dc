#!/bin/bash
# usage (on the GPU box): tools/launchlist.sh <tag> [bench args...]  -> gpurun_out/launches_<tag>.csv + summary on stdout
# COPER_NCU_WARM=1 keeps caches warm between kernels (closer to the in-graph timings; default = ncu's cold-cache mode)
tag=$1; shift
extra=""
if [ -n "$COPER_NCU_WARM" ]; then extra="--cache-control none"; fi
ncu --metrics gpu__time_duration.sum --clock-control none $extra -c 900 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-breakdown "$@" > gpurun_out/ncu_bench_$tag.log 2>&1
python tools/ncu_summary.py gpurun_out/launches_$tag.csv 60 | cut -c1-150
