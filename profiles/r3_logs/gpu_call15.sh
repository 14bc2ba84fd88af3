#!/bin/bash
# call 15 (1 GPU): dry sweep with the advection loop templated on the tile map, conflict-free box sums: parity, timings,
# compute-sanitizer on the final kernels
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=900 -k "dry or particle or precip or golden or drift or lightning or saves or per_pass" ) > gpurun_out/r3_c15_pytest.log 2>&1
tail -4 gpurun_out/r3_c15_pytest.log
( python profiles/tools/dry_probe2.py 2d-weather-sandbox_b200/csrc/libwsb200.so
  python profiles/tools/ab_bench.py --k 20 shipped=2d-weather-sandbox_b200/csrc/libwsb200.so
  python profiles/quick_particles.py 20 1000000; python profiles/quick_particles.py 20 2684354 ) > gpurun_out/r3_c15_timings.log 2>&1
cat gpurun_out/r3_c15_timings.log
for tool in memcheck racecheck synccheck; do
  ( time timeout 400 compute-sanitizer --tool $tool --print-limit 20 python profiles/tools/sanitize_target.py 2 ) > gpurun_out/r3_sanitizer_$tool.log 2>&1
  tail -4 gpurun_out/r3_sanitizer_$tool.log
done
