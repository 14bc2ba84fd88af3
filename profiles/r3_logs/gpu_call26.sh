#!/bin/bash
# call 26 (1 GPU): microbenchmark — does the row stride decide how long the ghost push takes?
mkdir -p gpurun_out
./gpurun_in/push_stride > gpurun_out/r3_c26_push_stride.log 2>&1; cat gpurun_out/r3_c26_push_stride.log
