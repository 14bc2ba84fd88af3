#!/bin/bash
# round-end style run: all GPU tests, the bench line, the ncu launch list of the bench command and
# one ncu --set full capture per fused kernel
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest.log 2>&1
tail -4 gpurun_out/r2_pytest.log
( time timeout 400 python bench.py ) > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -c 3000 gpurun_out/r2_bench.json
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_fused_dry --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_dry \
   python profiles/prof_target.py dry 16384 4096 5 > gpurun_out/r2_ncu_dry.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_fused --launch-skip 2 --launch-count 2 -f -o gpurun_out/r2_full \
   python profiles/prof_target.py full 16384 4096 3 > gpurun_out/r2_ncu_full.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-prewarm > gpurun_out/r2_launches_run.log 2>&1
ls -la gpurun_out | tail -12
