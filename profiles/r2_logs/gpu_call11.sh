#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -5
python profiles/tools/dry_probe2.py 2d-weather-sandbox_b200/csrc/libwsb200.so
python profiles/tools/dry_probe2.py gpurun_in/libwsb200_fma.so
python profiles/tools/dry_probe2.py 2d-weather-sandbox_b200/csrc/libwsb200.so 16448
} > gpurun_out/c11.log 2>&1
cat gpurun_out/c11.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_fused_dry --launch-skip 3 --launch-count 1 -f -o gpurun_out/c11_dry_persist \
   python profiles/prof_target.py dry 16384 4096 5 > gpurun_out/c11_ncu_dry.log 2>&1
