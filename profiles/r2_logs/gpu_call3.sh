#!/bin/bash
mkdir -p gpurun_out
python profiles/tools/dry_ramp.py c5e61e37=gpurun_in/libwsb200_5e61e37.so c4e41718=gpurun_in/libwsb200_4e41718.so r1=gpurun_in/libwsb200_r1.so > gpurun_out/c3_ramp.log 2>&1
grep "batch  [23]" gpurun_out/c3_ramp.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_fused_dry --launch-skip 3 --launch-count 1 -f -o gpurun_out/c3_dry_new \
   python profiles/prof_target.py dry 16384 4096 5 > gpurun_out/c3_ncu_dry.log 2>&1
WSB200_LIB=$PWD/gpurun_in/libwsb200_5e61e37.so timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_fused_dry --launch-skip 3 --launch-count 1 -f -o gpurun_out/c3_dry_5e61 \
   python profiles/prof_target.py dry 16384 4096 5 > gpurun_out/c3_ncu_dry2.log 2>&1
ls -la gpurun_out/
