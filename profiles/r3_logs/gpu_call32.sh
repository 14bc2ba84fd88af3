#!/bin/bash
# call 32 (4 GPUs): P2P attributes between ring neighbours on the 4-GPU box (NVLink or PCIe?) + the peer transport once more
mkdir -p gpurun_out
WSB_EXCHANGE=peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 4 --steps 20 --warmup 3 --quick > gpurun_out/r3_c32_n4_peer.json 2> gpurun_out/r3_c32_n4_peer.err
tail -c 1500 gpurun_out/r3_c32_n4_peer.json
nvidia-smi nvlink --status -i 0 2>&1 | head -12 > gpurun_out/r3_c32_nvlink.txt; nvidia-smi -q -i 0 2>&1 | grep -i -A6 "fabric" >> gpurun_out/r3_c32_nvlink.txt; cat gpurun_out/r3_c32_nvlink.txt
