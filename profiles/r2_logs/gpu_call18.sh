#!/bin/bash
mkdir -p gpurun_out
{
timeout 200 python -m pytest tests -m gpu -x -q -k "not full_size and not dry" 2>&1 | tail -3
timeout 200 python profiles/tools/ab_bench.py --k 20 prev=gpurun_in/libwsb200_prev.so new=2d-weather-sandbox_b200/csrc/libwsb200.so 2>&1 | grep -v "dry :"
} > gpurun_out/c18.log 2>&1
cat gpurun_out/c18.log
