#!/bin/bash
# call 2 (1 GPU): all GPU tests incl. the new one-GPU strip tests (peer-memory ghost exchange)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x --timeout=900 ) > gpurun_out/r3_c2_pytest.log 2>&1
tail -30 gpurun_out/r3_c2_pytest.log
