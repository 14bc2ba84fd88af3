#!/bin/bash
# call 37 (4 GPUs): transport "auto" now also times the landing-zone push ("peerc") on rings of >= 3 strips of >= 4096
# columns: one --quick run at the BASELINE width — calibration of all three transports on the live state + the chosen one
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus 4 --steps 20 --warmup 3 --quick > gpurun_out/r3_c37_n4_auto_quick.json 2> gpurun_out/r3_c37_n4_auto_quick.err
tail -c 1800 gpurun_out/r3_c37_n4_auto_quick.json; tail -3 gpurun_out/r3_c37_n4_auto_quick.err
