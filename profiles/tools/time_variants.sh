#!/bin/bash
# One GPU call: parity of the shipped build, then dry sweep + full step timings of it and of every variant in gpurun_in/.
#   gpurun --timeout 900 -- 'bash profiles/tools/time_variants.sh'
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -2
python profiles/tools/dry_probe2.py 2d-weather-sandbox_b200/csrc/libwsb200.so
for f in gpurun_in/libwsb200_*.so; do python profiles/tools/dry_probe2.py "$f"; done
args="shipped=2d-weather-sandbox_b200/csrc/libwsb200.so"
for f in gpurun_in/libwsb200_*.so; do n=$(basename "$f" .so); args="$args ${n#libwsb200_}=$f"; done
timeout 400 python profiles/tools/ab_bench.py --k 20 $args
} > gpurun_out/variants.log 2>&1
cat gpurun_out/variants.log
