#!/bin/bash
# call 4 (1 GPU): the new bench line (all BASELINE configs at N=1, sustained batches, e2e at K and K=200)
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/r3_c4_bench.json 2> gpurun_out/r3_c4_bench.err
tail -c 6000 gpurun_out/r3_c4_bench.json; tail -5 gpurun_out/r3_c4_bench.err
