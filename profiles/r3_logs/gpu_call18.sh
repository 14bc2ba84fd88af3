#!/bin/bash
# call 18 (8 GPUs): quick A/B of the push CTA count at N=8, then the full N=8 line with the default
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 8 --steps 20 --warmup 3 --quick > gpurun_out/r3_c18_$name.json 2> gpurun_out/r3_c18_$name.err
  echo "== $name"; tail -c 600 gpurun_out/r3_c18_$name.json | head -c 600; echo; }
run push16 WSB_DBG_PUSH_BLOCKS=16
run push37 WSB_DBG_PUSH_BLOCKS=37
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r3_c18_bench_n8.json 2> gpurun_out/r3_c18_bench_n8.err
tail -c 1200 gpurun_out/r3_c18_bench_n8.json
