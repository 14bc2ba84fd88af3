#!/bin/bash
# call 34 (1 GPU): the CUDA path against the vectors the reference's own shaders produced (tests/test_ref_shader_golden.py),
# smoke() with its new reference-shader check
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_ref_shader_golden.py -m gpu -q --timeout=500 ) > gpurun_out/r3_c34_pytest.log 2>&1
tail -5 gpurun_out/r3_c34_pytest.log
( time timeout 300 python __graft_entry__.py smoke ) > gpurun_out/r3_c34_smoke.log 2>&1; tail -5 gpurun_out/r3_c34_smoke.log
