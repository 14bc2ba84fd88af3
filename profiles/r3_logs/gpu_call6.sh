#!/bin/bash
# call 6 (2 GPUs): bench N=2, peer transport after moving the flag handshake into a one-thread kernel
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r3_c6_bench_n2_peer.json 2> gpurun_out/r3_c6_bench_n2_peer.err
tail -c 2500 gpurun_out/r3_c6_bench_n2_peer.json; tail -3 gpurun_out/r3_c6_bench_n2_peer.err
