#!/bin/bash
# times the Schur stages for every kernel-variant library under build/variants (tools/build_variant.sh) and the default one
# usage: tools/syrk_variants.sh <cfg> <scale> <iters>
cd "$(dirname "$0")/.."
cfg=${1:-4}; scale=${2:-0.3}; iters=${3:-5}
echo "default: $(python tools/schur_time.py $cfg $scale $iters 2>/dev/null | tail -1)"
for lib in robot_camera_calibration_b200/build/variants/librcc_ba_*.so; do
  echo "$(basename $lib .so): $(RCC_BA_LIB=$PWD/$lib python tools/schur_time.py $cfg $scale $iters 2>/dev/null | tail -1)"
done
