#!/bin/bash
# call 8 (1 GPU): the whole GPU suite with all 14 reference saves available (copied to gpurun_in/saves for this call only):
# one-GPU strips on the two-stream schedule, per-pass + fused parity on every save, lightning bolt, drift curve
mkdir -p gpurun_out
export WSB_REFERENCE_SAVES=$PWD/gpurun_in/saves
( time timeout 2400 python -m pytest tests -m gpu -q --timeout=1200 ) > gpurun_out/r3_c8_pytest.log 2>&1
tail -40 gpurun_out/r3_c8_pytest.log
cat gpurun_out/drift_curve.json
