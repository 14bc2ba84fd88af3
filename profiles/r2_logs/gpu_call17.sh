#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --particles > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -c 1500 gpurun_out/r2c_bench.json
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_fused_dry --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2c_dry \
   python profiles/prof_target.py dry 16384 4096 5 > gpurun_out/r2c_ncu_dry.log 2>&1
