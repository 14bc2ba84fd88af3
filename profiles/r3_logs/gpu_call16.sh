#!/bin/bash
# call 16 (2 GPUs): strips after the one-fence-per-CTA push kernel: quick A/B over the number of push CTAs, then the full N=2 line
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 20 --warmup 3 --quick > gpurun_out/r3_c16_$name.json 2> gpurun_out/r3_c16_$name.err
  echo "== $name"; tail -c 1200 gpurun_out/r3_c16_$name.json; }
run push37 WSB_DBG_PUSH_BLOCKS=37
run push148 WSB_DBG_PUSH_BLOCKS=148
run push8 WSB_DBG_PUSH_BLOCKS=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r3_c16_bench_n2.json 2> gpurun_out/r3_c16_bench_n2.err
tail -c 1500 gpurun_out/r3_c16_bench_n2.json
