#!/bin/bash
# programmatic dependent launch on every kernel of the library: tests, then A/B (COPER_PDL=1 / 0) bench lines in the
# same box, then ncu --set full of a few of the short kernels of the chain
set -x
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/z3_tests.out 2>&1; tail -n 12 gpurun_out/z3_tests.out
B="--no-cpu-baseline --num-labels 0 --no-alt --no-extra --no-breakdown"
for pdl in 1 0 1; do
  COPER_PDL=$pdl timeout 600 python bench.py --shape wn18rr --prec fp16x3 $B > gpurun_out/z3_bench_wn_pdl$pdl.json 2>> gpurun_out/z3_bench.err
  python - <<P
import json
d=json.loads(open('gpurun_out/z3_bench_wn_pdl$pdl.json').read().strip().splitlines()[-1])
print('pdl=$pdl wn', d['ms_per_step'], d['e2e']['ms_per_step'], d['eval']['ms_per_batch'], d['e2e']['eval_ms_per_batch'])
P
done
for pdl in 1 0; do
  COPER_PDL=$pdl timeout 600 python bench.py --shape fb15k-237 --prec fp16x3 $B > gpurun_out/z3_bench_fb_pdl$pdl.json 2>> gpurun_out/z3_bench.err
  python - <<P
import json
for n in ('fb',):
    d=json.loads(open('gpurun_out/z3_bench_%s_pdl$pdl.json'%n).read().strip().splitlines()[-1])
    print('pdl=$pdl', n, d['ms_per_step'], d['e2e']['ms_per_step'], d['eval']['ms_per_batch'])
P
done
COPER_PDL=1 timeout 900 python bench.py --shape synth-10m --prec bf16 --steps 10 --warmup 3 $B > gpurun_out/z3_bench_10m_pdl1.json 2>> gpurun_out/z3_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/z3_bench_10m_pdl1.json').read().strip().splitlines()[-1]); print('10m', d['ms_per_step'], d['e2e']['ms_per_step'], d['eval']['ms_per_batch'])"
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled \
  -k regex:'segscatter_small_kernel|cpg_fwd_finalize_kernel4|prepare_fp16x3_fused_kernel|colstats_kernel' -c 12 \
  -o gpurun_out/z3_small python bench.py --shape wn18rr --prec fp16x3 --steps 2 --warmup 1 $B > /dev/null 2> gpurun_out/z3_ncu.err
tail -n 3 gpurun_out/z3_ncu.err
