#!/bin/bash
# after the scatter / statistics rework: tests, WN18RR line, and the 10 M-entity step under three switches + its launch list
set -x
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/z4_tests.out 2>&1; tail -n 8 gpurun_out/z4_tests.out
B="--no-cpu-baseline --num-labels 0 --no-alt --no-extra --no-breakdown"
timeout 600 python bench.py --shape wn18rr --prec fp16x3 $B > gpurun_out/z4_bench_wn.json 2>> gpurun_out/z4_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/z4_bench_wn.json').read().strip().splitlines()[-1]); print('wn', d['ms_per_step'], d['e2e']['ms_per_step'], d['eval']['ms_per_batch'], d['e2e']['eval_ms_per_batch'], d.get('gpu_launches_per_step'))"
run10m() {
  env "$@" timeout 900 python bench.py --shape synth-10m --prec bf16 --steps 8 --warmup 3 $B > gpurun_out/z4_bench_10m_$TAG.json 2>> gpurun_out/z4_bench.err
  python -c "
import json
d=json.loads(open('gpurun_out/z4_bench_10m_$TAG.json').read().strip().splitlines()[-1]); print('10m $TAG', d['ms_per_step'], d['e2e']['ms_per_step'], d['eval']['ms_per_batch'])"
}
TAG=default run10m COPER_PDL=1
TAG=nopdl run10m COPER_PDL=0
TAG=nosplit run10m COPER_SPLIT_CPG_BWD=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/z4_launches_10m.csv \
  python bench.py --shape synth-10m --prec bf16 --steps 2 --warmup 3 $B > /dev/null 2> gpurun_out/z4_1.err
tail -n 2 gpurun_out/z4_1.err
