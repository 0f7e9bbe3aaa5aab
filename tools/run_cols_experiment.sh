# SYRK v2 with 1 / 2 / 3 columns of a partner block per lane item (RCC_SY_COLS): parity of the 2-column build, then timings
set -x
cd $GRAFT_REPO_ROOT
V=$PWD/robot_camera_calibration_b200/build/variants
RCC_BA_LIB=$V/librcc_ba_cols2.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -3
echo "cfg4 0.2 default: $(python tools/schur_time.py 4 0.2 3 2>/dev/null | tail -1)"
for v in cols2 cols2c4 cols3c4; do echo "cfg4 0.2 $v: $(RCC_BA_LIB=$V/librcc_ba_$v.so python tools/schur_time.py 4 0.2 3 2>/dev/null | tail -1)"; done
echo "cfg2 default: $(python tools/schur_time.py 2 1.0 10 2>/dev/null | tail -1)"
for v in cols2 cols2c4 cols3c4; do echo "cfg2 $v: $(RCC_BA_LIB=$V/librcc_ba_$v.so python tools/schur_time.py 2 1.0 10 2>/dev/null | tail -1)"; done
