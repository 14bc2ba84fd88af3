#!/bin/bash
# round 2 (files r3_*), call 1: GPU tests of the inherited build, A/B timings of the prepared variants,
# compute-sanitizer memcheck / racecheck / synccheck on the smoke configurations
mkdir -p gpurun_out
bash profiles/tools/time_variants.sh > /dev/null 2>&1
cp gpurun_out/variants.log gpurun_out/r3_c1_variants.log
for tool in memcheck racecheck synccheck; do
  ( time timeout 400 compute-sanitizer --tool $tool --print-limit 20 python profiles/tools/sanitize_target.py 2 ) > gpurun_out/r3_sanitizer_$tool.log 2>&1
  tail -5 gpurun_out/r3_sanitizer_$tool.log
done
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv >> gpurun_out/r3_c1_variants.log
cat gpurun_out/r3_c1_variants.log
