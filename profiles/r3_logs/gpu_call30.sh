#!/bin/bash
# call 30 (2 GPUs): is the slow push at 4 GPUs a property of the strip width? 2 GPUs, total width 8192 / 4096 -> strips of 4096 / 2048 columns
mkdir -p gpurun_out
run() { name=$1; w=$2; shift; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 20 --warmup 3 --quick --width $w > gpurun_out/r3_c30_$name.json 2> gpurun_out/r3_c30_$name.err
  echo "== $name"; tail -c 700 gpurun_out/r3_c30_$name.json; echo; }
run w8192_peer 8192 WSB_EXCHANGE=peer
run w8192_nopush 8192 WSB_EXCHANGE=peer WSB_DBG_NOPUSH=1
run w8192_nccl 8192 WSB_EXCHANGE=nccl
run w4096_peer 4096 WSB_EXCHANGE=peer
run w4096_nopush 4096 WSB_EXCHANGE=peer WSB_DBG_NOPUSH=1
