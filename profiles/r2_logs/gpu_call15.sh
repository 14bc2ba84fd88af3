#!/bin/bash
mkdir -p gpurun_out
{
for v in ty28 ty20; do python profiles/tools/dry_probe2.py gpurun_in/libwsb200_$v.so; done
} > gpurun_out/c15.log 2>&1
cat gpurun_out/c15.log
