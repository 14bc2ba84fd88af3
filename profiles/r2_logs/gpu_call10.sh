#!/bin/bash
mkdir -p gpurun_out
{
for v in notmast nonear fma; do python profiles/tools/dry_probe2.py gpurun_in/libwsb200_$v.so; done
python profiles/tools/dry_probe2.py 2d-weather-sandbox_b200/csrc/libwsb200.so
python profiles/tools/dry_probe2.py gpurun_in/libwsb200_notmast.so 16448
python profiles/tools/ab_bench.py --k 20 r1=gpurun_in/libwsb200_r1.so new=2d-weather-sandbox_b200/csrc/libwsb200.so fma=gpurun_in/libwsb200_fma.so
timeout 600 python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -3
} > gpurun_out/c10.log 2>&1
cat gpurun_out/c10.log
