#!/bin/bash
# call 35 (1 GPU): final tree — whole GPU suite, smoke(), the default bench line, the reference arm (oracle/_ref on the host cores)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q --timeout=900 ) > gpurun_out/r3_c35_pytest.log 2>&1
tail -4 gpurun_out/r3_c35_pytest.log
( time python __graft_entry__.py smoke ) > gpurun_out/r3_c35_smoke.log 2>&1; tail -5 gpurun_out/r3_c35_smoke.log
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/r3_c35_bench_reference_arm.json 2> gpurun_out/r3_c35_bench_reference_arm.err
tail -c 600 gpurun_out/r3_c35_bench_reference_arm.json; tail -3 gpurun_out/r3_c35_bench_reference_arm.err
( time timeout 900 python bench.py ) > gpurun_out/r3_c35_bench.json 2> gpurun_out/r3_c35_bench.err
tail -c 1500 gpurun_out/r3_c35_bench.json; tail -3 gpurun_out/r3_c35_bench.err
