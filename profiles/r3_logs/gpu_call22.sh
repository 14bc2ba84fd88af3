#!/bin/bash
# call 22 (4 GPUs): why is N=4 slower than its share? topology + quick A/B
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r3_c22_topo.txt 2>&1
run() { name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 4 --steps 20 --warmup 3 --quick > gpurun_out/r3_c22_$name.json 2> gpurun_out/r3_c22_$name.err
  echo "== $name"; tail -c 300 gpurun_out/r3_c22_$name.json; echo; }
run shipped WSB_DBG_NOPUSH=0
run nopush WSB_DBG_NOPUSH=1
run push37 WSB_DBG_PUSH_BLOCKS=37
run nccl WSB_EXCHANGE=nccl
head -12 gpurun_out/r3_c22_topo.txt
