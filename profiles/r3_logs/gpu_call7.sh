#!/bin/bash
# call 7 (2 GPUs): bench N=2 with per-rank kernel times: nccl, peer, nccl, peer
mkdir -p gpurun_out
for i in 1 2; do for t in nccl peer; do
  WSB_EXCHANGE=$t timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r3_c7_bench_n2_${t}_$i.json 2> gpurun_out/r3_c7_bench_n2_${t}_$i.err
  tail -c 1500 gpurun_out/r3_c7_bench_n2_${t}_$i.json
done; done
