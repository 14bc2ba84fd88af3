#!/bin/bash
mkdir -p gpurun_out
{
for v in r1novmax novmax notmast r1novmax novmax notmast; do python profiles/tools/dry_probe2.py gpurun_in/libwsb200_$v.so; done
} > gpurun_out/c9_probe.log 2>&1
cat gpurun_out/c9_probe.log
