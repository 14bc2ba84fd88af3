#!/bin/bash
mkdir -p gpurun_out
python profiles/tools/dry_probe.py r1=gpurun_in/libwsb200_r1.so new=2d-weather-sandbox_b200/csrc/libwsb200.so > gpurun_out/c4_probe.log 2>&1
cat gpurun_out/c4_probe.log
for v in r1 notmast nonear; do
WSB200_LIB=$PWD/gpurun_in/libwsb200_$v.so timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_fused_dry --launch-skip 3 --launch-count 2 --csv --log-file gpurun_out/c4_ncu_$v.csv python profiles/prof_target.py dry 16384 4096 5 > /dev/null 2>&1
grep -h "k_fused" gpurun_out/c4_ncu_$v.csv | cut -d, -f1,5,12- | sed "s/^/$v /"
done
