#!/bin/bash
# call 11 (1 GPU): full GPU suite (box-filter particle pass, N-API addon) + bench N=1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout=900 ) > gpurun_out/r3_c11_pytest.log 2>&1
tail -15 gpurun_out/r3_c11_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/r3_c11_bench.json 2> gpurun_out/r3_c11_bench.err
tail -c 3000 gpurun_out/r3_c11_bench.json; tail -5 gpurun_out/r3_c11_bench.err
