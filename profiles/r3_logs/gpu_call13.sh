#!/bin/bash
# call 13 (1 GPU): parity of the dirty-list particle pass and the wall-tile-map dry sweep, then their timings
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=900 ) > gpurun_out/r3_c13_pytest.log 2>&1
tail -5 gpurun_out/r3_c13_pytest.log
( python profiles/quick_particles.py 20 1000000; python profiles/quick_particles.py 20 2684354
  python profiles/tools/dry_probe2.py 2d-weather-sandbox_b200/csrc/libwsb200.so
  python profiles/tools/ab_bench.py --k 20 shipped=2d-weather-sandbox_b200/csrc/libwsb200.so ) > gpurun_out/r3_c13_timings.log 2>&1
cat gpurun_out/r3_c13_timings.log
