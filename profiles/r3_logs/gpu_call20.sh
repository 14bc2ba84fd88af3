#!/bin/bash
# call 20 (1 GPU): ncu launch list of the bench command (shares of the step must agree with the CUDA-event spans of the un-profiled run)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-prewarm > gpurun_out/r3_c20_launches_run.log 2>&1
tail -c 400 gpurun_out/r3_c20_launches_run.log; wc -l gpurun_out/r3_launches.csv
