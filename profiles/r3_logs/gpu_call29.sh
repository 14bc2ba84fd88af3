#!/bin/bash
# call 29 (1 GPU): final build — whole GPU suite, the strip tests three more times (flakiness), smoke(), bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout=900 ) > gpurun_out/r3_c29_pytest.log 2>&1
tail -4 gpurun_out/r3_c29_pytest.log
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_strips.py -m gpu -q --timeout=500 2>&1 | tail -1; done > gpurun_out/r3_c29_strips_repeat.log
cat gpurun_out/r3_c29_strips_repeat.log
( time python __graft_entry__.py smoke ) > gpurun_out/r3_c29_smoke.log 2>&1; tail -4 gpurun_out/r3_c29_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/r3_c29_bench.json 2> gpurun_out/r3_c29_bench.err
tail -c 800 gpurun_out/r3_c29_bench.json; tail -3 gpurun_out/r3_c29_bench.err
