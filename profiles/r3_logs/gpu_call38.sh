#!/bin/bash
# call 38 (2 GPUs): the landing-zone push ("peerc") between two REAL GPUs over NVLink, bit-identical to the single-GPU run
mkdir -p gpurun_out
( time timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=180 -k "peerc and 2-" ) > gpurun_out/r3_c38_pytest_multi_peerc_n2.log 2>&1
tail -6 gpurun_out/r3_c38_pytest_multi_peerc_n2.log
