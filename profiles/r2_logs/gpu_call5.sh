#!/bin/bash
mkdir -p gpurun_out
{
for i in 1 2 3; do python profiles/tools/dry_probe2.py gpurun_in/libwsb200_r1.so; done
for i in 1 2 3; do python profiles/tools/dry_probe2.py 2d-weather-sandbox_b200/csrc/libwsb200.so; done
python profiles/tools/dry_probe2.py gpurun_in/libwsb200_r1.so 16448
python profiles/tools/dry_probe2.py 2d-weather-sandbox_b200/csrc/libwsb200.so 16448
python profiles/tools/dry_probe2.py gpurun_in/libwsb200_r1.so 16448
python profiles/tools/dry_probe2.py 2d-weather-sandbox_b200/csrc/libwsb200.so 16448
} > gpurun_out/c5_probe2.log 2>&1
cat gpurun_out/c5_probe2.log
