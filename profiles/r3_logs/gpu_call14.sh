#!/bin/bash
# call 14 (1 GPU): ncu --set full captures: particle kernels, dry sweep (wall tile map), the two full-physics kernels
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none -f"
timeout 300 $NCU -k regex:'k_boxsum|k_precipitation|k_clear_origins' --launch-skip 60 --launch-count 3 -o gpurun_out/r3_particles \
   python profiles/quick_particles.py 2 1000000 > gpurun_out/r3_c14_ncu_particles.log 2>&1
timeout 300 $NCU -k regex:k_fused_dry --launch-skip 3 --launch-count 1 -o gpurun_out/r3_dry \
   python profiles/prof_target.py dry 16384 4096 5 > gpurun_out/r3_c14_ncu_dry.log 2>&1
timeout 300 $NCU -k regex:k_fused --launch-skip 2 --launch-count 2 -o gpurun_out/r3_full \
   python profiles/prof_target.py full 16384 4096 3 > gpurun_out/r3_c14_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/r3_c14_ncu_*.log
