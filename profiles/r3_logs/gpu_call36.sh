#!/bin/bash
# call 36 (1 GPU): the landing-zone push (transport "peerc") with every rank on one GPU, beside the existing strip tests
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_strips.py -m gpu -q --timeout=500 ) > gpurun_out/r3_c36_pytest_strips.log 2>&1
tail -15 gpurun_out/r3_c36_pytest_strips.log
