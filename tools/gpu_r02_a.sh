#!/bin/bash
# round-2 GPU call A: compute-sanitizer over the tcgen05 kernels at toy shapes + baseline bench lines of the other shapes
set -x
mkdir -p gpurun_out
K='7-97-40 or 5-8-8 or 130-77-301'
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck.log \
  python -m pytest tests/test_gpu_umma.py -m gpu -q -x -k "$K" > gpurun_out/sanitizer_memcheck.out 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.out
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/sanitizer_racecheck.log \
  python -m pytest tests/test_gpu_umma.py -m gpu -q -x -k "$K" > gpurun_out/sanitizer_racecheck.out 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.out
for s in fb15k-237 nell-995 yago3-10; do
  timeout 600 python bench.py --shape $s --steps 20 --warmup 5 --no-cpu-baseline --num-labels 0 > gpurun_out/base_$s.json 2> gpurun_out/base_$s.err
done
tail -3 gpurun_out/sanitizer_*.out
