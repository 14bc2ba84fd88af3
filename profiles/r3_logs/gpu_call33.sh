#!/bin/bash
# call 33 (4 GPUs): what makes the peer push slow on 4 GPUs — the number of ranks or the strip width?
mkdir -p gpurun_out
run() { name=$1; np=$2; w=$3; shift; shift; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $np --steps 20 --warmup 3 --quick --width $w > gpurun_out/r3_c33_$name.json 2> gpurun_out/r3_c33_$name.err
  echo "== $name"; tail -c 300 gpurun_out/r3_c33_$name.json; echo; }
run n3_w12288_peer 3 12288 WSB_EXCHANGE=peer
run n3_w12288_nopush 3 12288 WSB_EXCHANGE=peer WSB_DBG_NOPUSH=1
run n4_w8192_peer 4 8192 WSB_EXCHANGE=peer
run n4_w8192_nopush 4 8192 WSB_EXCHANGE=peer WSB_DBG_NOPUSH=1
run n4_w32768_peer 4 32768 WSB_EXCHANGE=peer
run n4_w32768_nopush 4 32768 WSB_EXCHANGE=peer WSB_DBG_NOPUSH=1
