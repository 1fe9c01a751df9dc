#!/bin/bash
# 8-GPU call: the N=4 and N=8 bench lines (10M strong headline + wn18rr weak block + 1-GPU base)
set -x
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/m8_bench4.json 2> gpurun_out/m8_bench4.err
tail -n 3 gpurun_out/m8_bench4.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/m8_bench8.json 2> gpurun_out/m8_bench8.err
tail -n 3 gpurun_out/m8_bench8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 8 --steps 5 --warmup 3 > gpurun_out/m8_ref8.json 2> gpurun_out/m8_ref8.err
tail -n 3 gpurun_out/m8_ref8.err
