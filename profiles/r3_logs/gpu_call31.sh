#!/bin/bash
# call 31 (8 GPUs): how far is N=8 from the no-exchange bound?  quick runs: peer / no push at all / NCCL
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 8 --steps 20 --warmup 3 --quick > gpurun_out/r3_c31_$name.json 2> gpurun_out/r3_c31_$name.err
  echo "== $name"; tail -c 400 gpurun_out/r3_c31_$name.json; echo; }
run nopush WSB_EXCHANGE=peer WSB_DBG_NOPUSH=1
run peer WSB_EXCHANGE=peer
