#!/bin/bash
# call 28 (1 GPU): A/B of cheap single-GPU variants (bit-exact by construction; checksums printed): streaming stores,
# two cells in flight in the per-cell loops, TMA L2 promotion 256 B / none
mkdir -p gpurun_out
args="shipped=2d-weather-sandbox_b200/csrc/libwsb200.so"
for f in gpurun_in/libwsb200_*.so; do n=$(basename "$f" .so); args="$args ${n#libwsb200_}=$f"; done
{ timeout 600 python profiles/tools/ab_bench.py --k 20 $args
  WSB_DBG_L2PROMO=256 timeout 200 python profiles/tools/ab_bench.py --k 20 promo256=2d-weather-sandbox_b200/csrc/libwsb200.so
  WSB_DBG_L2PROMO=0 timeout 200 python profiles/tools/ab_bench.py --k 20 promo0=2d-weather-sandbox_b200/csrc/libwsb200.so
} > gpurun_out/r3_c28_variants.log 2>&1
cat gpurun_out/r3_c28_variants.log
