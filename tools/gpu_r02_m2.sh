#!/bin/bash
# 2-GPU call: NCCL parity tests + the N=2 bench line (10M strong headline, wn18rr weak block, 1-GPU base)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/m2_multi.out 2>&1; tail -n 8 gpurun_out/m2_multi.out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/m2_bench2.json 2> gpurun_out/m2_bench2.err
tail -n 5 gpurun_out/m2_bench2.err
cat gpurun_out/m2_bench2.json | head -c 3000
