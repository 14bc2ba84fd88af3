#!/bin/bash
# call 19 (1 GPU): the whole GPU suite, smoke(), and the full bench line of the final build
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout=900 ) > gpurun_out/r3_c19_pytest.log 2>&1
tail -6 gpurun_out/r3_c19_pytest.log
( time python __graft_entry__.py smoke ) > gpurun_out/r3_c19_smoke.log 2>&1; tail -4 gpurun_out/r3_c19_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/r3_c19_bench.json 2> gpurun_out/r3_c19_bench.err
tail -c 1500 gpurun_out/r3_c19_bench.json; tail -4 gpurun_out/r3_c19_bench.err
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/r3_c19_bench_ref.json 2>&1; tail -c 600 gpurun_out/r3_c19_bench_ref.json
