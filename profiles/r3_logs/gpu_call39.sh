#!/bin/bash
# call 39 (1 GPU): final build of libwsb200.so (landing-zone transport added) — smoke() and the reference-shader vectors
mkdir -p gpurun_out
( time python __graft_entry__.py smoke ) > gpurun_out/r3_c39_smoke.log 2>&1; tail -4 gpurun_out/r3_c39_smoke.log
( time timeout 200 python -m pytest tests/test_ref_shader_golden.py -m gpu -q --timeout=150 ) > gpurun_out/r3_c39_pytest_golden.log 2>&1; tail -3 gpurun_out/r3_c39_pytest_golden.log
