#!/bin/bash
# call 25 (8 GPUs): final bench line at N=8 (transport auto)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r3_c25_bench_n8.json 2> gpurun_out/r3_c25_bench_n8.err
tail -c 1000 gpurun_out/r3_c25_bench_n8.json; tail -2 gpurun_out/r3_c25_bench_n8.err
