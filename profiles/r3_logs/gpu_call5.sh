#!/bin/bash
# call 5 (2 GPUs): real 2-GPU strips vs 1 GPU (both transports), bench N=2 with the peer and the NCCL transport
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=500 ) > gpurun_out/r3_c5_pytest.log 2>&1
tail -5 gpurun_out/r3_c5_pytest.log
for t in peer nccl; do
  WSB_EXCHANGE=$t timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r3_c5_bench_n2_$t.json 2> gpurun_out/r3_c5_bench_n2_$t.err
  tail -c 2500 gpurun_out/r3_c5_bench_n2_$t.json; tail -3 gpurun_out/r3_c5_bench_n2_$t.err
done
nvidia-smi topo -m > gpurun_out/r3_c5_topo.txt 2>&1
