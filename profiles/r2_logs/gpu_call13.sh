#!/bin/bash
mkdir -p gpurun_out
{
WSB200_LIB=$PWD/gpurun_in/libwsb200_ty32.so timeout 600 python -m pytest tests -m gpu -x -q -k "dry or moderate or fast_flow" 2>&1 | tail -3
python profiles/tools/dry_probe2.py 2d-weather-sandbox_b200/csrc/libwsb200.so
for v in ty32 ty32q ty32fma ty32c4; do python profiles/tools/dry_probe2.py gpurun_in/libwsb200_$v.so; done
} > gpurun_out/c13.log 2>&1
cat gpurun_out/c13.log
