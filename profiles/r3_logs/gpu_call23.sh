#!/bin/bash
# call 23 (4 GPUs): transport "auto" (both set up, calibrated on the live state): multi-GPU tests of the switch, N=4 bench line
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=500 -k "auto" ) > gpurun_out/r3_c23_pytest.log 2>&1
tail -4 gpurun_out/r3_c23_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r3_c23_bench_n4.json 2> gpurun_out/r3_c23_bench_n4.err
tail -c 1200 gpurun_out/r3_c23_bench_n4.json; tail -2 gpurun_out/r3_c23_bench_n4.err
