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
build pairadv4 "-DWSB_DRY_PAIRADV=1"
build pairadv3 "-DWSB_DRY_PAIRADV=1 -DWSB_DRY_CTAS=3"
build pairadv3t32 "-DWSB_DRY_PAIRADV=1 -DWSB_DRY_CTAS=3 -DWSB_DRY_TY=32"
build ty24 "-DWSB_DRY_TY=24"
build quad "-DWSB_SWEEP_QUAD=1"
build fma "-DWSB_EXP_FMAMIX"   # timing only: NOT the frozen arithmetic
