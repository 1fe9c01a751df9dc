#!/bin/bash
# SM share of the side stream's GEMMs (coper_set_sm_budget) at the WN18RR / FB15k-237 shapes; 10 M step with the per-model PDL rule
set -x
mkdir -p gpurun_out
B="--no-cpu-baseline --num-labels 0 --no-alt --no-extra --no-breakdown"
for sms in 0 48 74 100 0; do
  COPER_SIDE_SMS=$sms timeout 600 python bench.py --shape wn18rr --prec fp16x3 $B > gpurun_out/z5_bench_wn_s$sms.json 2>> gpurun_out/z5_bench.err
  python -c "
import json
d=json.loads(open('gpurun_out/z5_bench_wn_s$sms.json').read().strip().splitlines()[-1]); print('wn side_sms=$sms', d['ms_per_step'], d['e2e']['ms_per_step'], d['eval']['ms_per_batch'])"
done
for sms in 0 74; do
  COPER_SIDE_SMS=$sms timeout 600 python bench.py --shape fb15k-237 --prec fp16x3 $B > gpurun_out/z5_bench_fb_s$sms.json 2>> gpurun_out/z5_bench.err
  python -c "
import json
d=json.loads(open('gpurun_out/z5_bench_fb_s$sms.json').read().strip().splitlines()[-1]); print('fb side_sms=$sms', d['ms_per_step'], d['e2e']['ms_per_step'], d['eval']['ms_per_batch'])"
done
timeout 900 python bench.py --shape synth-10m --prec bf16 --steps 8 --warmup 3 $B > gpurun_out/z5_bench_10m.json 2>> gpurun_out/z5_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/z5_bench_10m.json').read().strip().splitlines()[-1]); print('10m', d['ms_per_step'], d['e2e']['ms_per_step'], d['eval']['ms_per_batch'])"
