#!/bin/bash
# call 12 (2 GPUs): what slows the interior kernels of the two-stream strip schedule? quick bench: as shipped / no push at all /
# slow polling; then particle-pass timing on GPU 0 (register-blocked box sums)
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 20 --warmup 3 --quick > gpurun_out/r3_c12_$name.json 2> gpurun_out/r3_c12_$name.err
  echo "== $name"; tail -c 1500 gpurun_out/r3_c12_$name.json
}
run shipped WSB_DBG_NOPUSH=0
run nopush WSB_DBG_NOPUSH=1
run poll2000 WSB_DBG_POLL_NS=2000
run nccl WSB_EXCHANGE=nccl
( python profiles/quick_particles.py 20 1000000; python profiles/quick_particles.py 20 2684354 ) > gpurun_out/r3_c12_particles.log 2>&1
cat gpurun_out/r3_c12_particles.log
