#!/bin/bash
# first GPU call of the session: parity of the new kernels, A/B timings, instruction counts
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size" ) > gpurun_out/c1_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c1_pytest.log
( time timeout 600 python profiles/tools/ab_bench.py --k 20 r1=gpurun_in/libwsb200_r1.so new=2d-weather-sandbox_b200/csrc/libwsb200.so \
    nonear=gpurun_in/libwsb200_nonear.so notmast=gpurun_in/libwsb200_notmast.so fma=gpurun_in/libwsb200_fma.so ) > gpurun_out/c1_ab.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:k_fused --launch-skip 48 --launch-count 12 --csv --log-file gpurun_out/c1_ncu_inst.csv \
    python profiles/tools/ab_bench.py --small --k 4 new=2d-weather-sandbox_b200/csrc/libwsb200.so > gpurun_out/c1_ncu_run.log 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q -k "full_size" ) > gpurun_out/c1_pytest_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/c1_pytest_full.log
tail -5 gpurun_out/c1_pytest.log; cat gpurun_out/c1_ab.log; tail -3 gpurun_out/c1_pytest_full.log
