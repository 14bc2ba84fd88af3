#!/bin/bash
# call 10 (8 GPUs): bench N=8, peer transport, two-stream schedule
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r3_c10_bench_n8_peer.json 2> gpurun_out/r3_c10_bench_n8_peer.err
tail -c 3000 gpurun_out/r3_c10_bench_n8_peer.json; tail -3 gpurun_out/r3_c10_bench_n8_peer.err
