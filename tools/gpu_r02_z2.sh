#!/bin/bash
# merged small launches + generator weight-gradient half on the side stream: new tests first, then the model-level
# suites, a WN18RR bench line, and a warm-cache launch list of the step
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_merged_launches.py -x -q -m gpu ) > gpurun_out/z2_tests_new.out 2>&1; tail -n 15 gpurun_out/z2_tests_new.out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu --deselect tests/test_gpu_merged_launches.py ) > gpurun_out/z2_tests.out 2>&1; tail -n 12 gpurun_out/z2_tests.out
B="--no-cpu-baseline --num-labels 0 --no-alt --no-extra"
( timeout 600 python bench.py --shape wn18rr --prec fp16x3 $B ) > gpurun_out/z2_bench_wn.json 2> gpurun_out/z2_bench.err; tail -n 3 gpurun_out/z2_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/z2_bench_wn.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e'], d.get('gpu_launches_per_step'), d.get('eval'))
print(d.get('kernel_ms'))
P
( timeout 600 python bench.py --shape fb15k-237 --prec fp16x3 $B --no-breakdown ) > gpurun_out/z2_bench_fb.json 2>> gpurun_out/z2_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/z2_bench_fb.json').read().strip().splitlines()[-1]); print('fb', d['ms_per_step'], d['value'], d.get('eval'))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1500 --csv --log-file gpurun_out/z2_launches_wn18rr_warm.csv \
  python bench.py --shape wn18rr --prec fp16x3 --steps 3 --warmup 3 $B --no-breakdown > /dev/null 2> gpurun_out/z2_1.err
tail -n 2 gpurun_out/z2_1.err
