#!/bin/bash
# call 17 (2 GPUs): push kernel with contiguous per-CTA runs (page locality of the peer stores): quick A/B over CTA counts
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 20 --warmup 3 --quick > gpurun_out/r3_c17_$name.json 2> gpurun_out/r3_c17_$name.err
  echo "== $name"; tail -c 900 gpurun_out/r3_c17_$name.json; }
run push37 WSB_DBG_PUSH_BLOCKS=37
run push16 WSB_DBG_PUSH_BLOCKS=16
run push8 WSB_DBG_PUSH_BLOCKS=8
run push74 WSB_DBG_PUSH_BLOCKS=74
