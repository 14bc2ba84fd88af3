#!/bin/bash
# call 21 (4 GPUs): the N=4 line
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r3_c21_bench_n4.json 2> gpurun_out/r3_c21_bench_n4.err
tail -c 1500 gpurun_out/r3_c21_bench_n4.json; tail -2 gpurun_out/r3_c21_bench_n4.err
