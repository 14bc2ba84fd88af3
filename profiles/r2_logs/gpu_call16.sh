#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests -m gpu -x -q -k "dry or moderate or fast_flow" 2>&1 | tail -3
python profiles/tools/dry_probe2.py 2d-weather-sandbox_b200/csrc/libwsb200.so
} > gpurun_out/c16.log 2>&1
cat gpurun_out/c16.log
