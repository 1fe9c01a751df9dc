#!/bin/bash
# A/B in one box: activation + operand form in one launch (COPER_FUSE_ACT_PREPARE), Conv1BN backward folded into the conv
# backward (COPER_FOLD_CONV_BN)
set -x
mkdir -p gpurun_out
B="--no-cpu-baseline --num-labels 0 --no-alt --no-extra --no-breakdown"
for shape in wn18rr fb15k-237; do
for cfg in 11 01 10 00 11; do
  COPER_FUSE_ACT_PREPARE=${cfg:0:1} COPER_FOLD_CONV_BN=${cfg:1:1} timeout 300 python bench.py --shape $shape --prec fp16x3 $B > gpurun_out/z6_$shape-$cfg.json 2>> gpurun_out/z6_bench.err
  python -c "
import json
d=json.loads(open('gpurun_out/z6_$shape-$cfg.json').read().strip().splitlines()[-1]); print('$shape fuse_act,fold_conv=$cfg', round(d['ms_per_step'],4), round(d['eval']['ms_per_batch'],4), d['gpu_launches_per_step'])"
done
done
