#!/bin/bash
# call 24 (2 GPUs): final bench line at N=2 (transport auto) + the multi-GPU tests at world 2
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=500 ) > gpurun_out/r3_c24_pytest.log 2>&1
tail -3 gpurun_out/r3_c24_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r3_c24_bench_n2.json 2> gpurun_out/r3_c24_bench_n2.err
tail -c 1000 gpurun_out/r3_c24_bench_n2.json; tail -2 gpurun_out/r3_c24_bench_n2.err
