#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -3
python profiles/tools/dry_probe2.py 2d-weather-sandbox_b200/csrc/libwsb200.so
for v in quad4 pair5 quad5 pair6; do python profiles/tools/dry_probe2.py gpurun_in/libwsb200_$v.so; done
python profiles/tools/ab_bench.py --k 20 new=2d-weather-sandbox_b200/csrc/libwsb200.so
} > gpurun_out/c12.log 2>&1
cat gpurun_out/c12.log
