#!/bin/bash
# Builds A/B variants of libwsb200 into gpurun_in/ (git-ignored, travels with gpurun) — run on the CPU
# container, then time them in ONE gpu call with profiles/tools/time_variants.sh.
#   bash profiles/tools/build_variants.sh
set -e
cd "$(dirname "$0")/../.."
mkdir -p gpurun_in
CSRC=2d-weather-sandbox_b200/csrc
build() {  # name, flags
  make -C $CSRC -s variant OUT="$PWD/gpurun_in/libwsb200_$1.so" EXTRA="$2" > /dev/null
  printf "%-12s %s | " "$1" "$2"
  grep -A2 "k_fused_dry" "gpurun_in/libwsb200_$1.so.log" | grep -E "Used|spill" | tr -s ' ' | tr '\n' ' '
  echo
}
# (round 2: the two-cells-per-thread advection variants (WSB_DRY_PAIRADV) and the 16-byte sweeps (WSB_SWEEP_QUAD) were timed —
#  profiles/r3_logs/c1_variants.log: 27-66 % and 2 % slower — and removed from the sources)
build ty24 "-DWSB_DRY_TY=24"
build fma "-DWSB_EXP_FMAMIX"   # timing only: NOT the frozen arithmetic
