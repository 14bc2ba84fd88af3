#!/bin/bash
# call 9 (2 GPUs): bench N=2 on the two-stream schedule
mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r3_c9_bench_n2_peer_$i.json 2> gpurun_out/r3_c9_bench_n2_peer_$i.err
tail -c 1800 gpurun_out/r3_c9_bench_n2_peer_$i.json; tail -2 gpurun_out/r3_c9_bench_n2_peer_$i.err
done
