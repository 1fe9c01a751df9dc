#!/bin/bash
# final multi-GPU lines: N=2, 4, 8 (10M strong headline + wn18rr weak block + 1-GPU base)
set -x
mkdir -p gpurun_out
for n in 2 4 8; do
  devs=$(seq -s, 0 $((n-1)))
  CUDA_VISIBLE_DEVICES=$devs timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/m8b_bench$n.json 2> gpurun_out/m8b_bench$n.err
  tail -n 2 gpurun_out/m8b_bench$n.err
done
