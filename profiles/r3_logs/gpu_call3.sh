#!/bin/bash
# call 3 (1 GPU): one-GPU strip tests after the kernel preload fix + the oracle-window tests at BASELINE sizes
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_strips.py tests/test_gpu_parity.py -m gpu -q --timeout=900 -k "strips or full_size" ) > gpurun_out/r3_c3_pytest.log 2>&1
tail -40 gpurun_out/r3_c3_pytest.log
