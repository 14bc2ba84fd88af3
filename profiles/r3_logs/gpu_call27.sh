#!/bin/bash
# call 27 (1 GPU): contraction-allowed build (-DWSB_EXP_FMAMIX, not shipped) against the oracle: drift after 1 / 10 / 100 / 1000 iterations
mkdir -p gpurun_out
timeout 600 python profiles/tools/fma_tolerance.py gpurun_in/libwsb200_fma.so > gpurun_out/r3_c27_fma_tolerance.log 2>&1
cat gpurun_out/r3_c27_fma_tolerance.log
