#!/bin/bash
mkdir -p gpurun_out
{
for v in ty24 ty48 ty48r; do python profiles/tools/dry_probe2.py gpurun_in/libwsb200_$v.so; done
} > gpurun_out/c14.log 2>&1
cat gpurun_out/c14.log
