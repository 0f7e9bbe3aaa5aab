# SYRK v2 with one warp per CTA and wider warp sub-tiles (Y_ef is re-read from shared memory once per (pair, sub-tile):
# a broadcast LDS.128 costs two wavefronts, 36 of the ~95 per visit)
cd $GRAFT_REPO_ROOT
V=$PWD/robot_camera_calibration_b200/build/variants
echo "cfg4 0.2 default: $(python tools/schur_time.py 4 0.2 3 2>/dev/null | tail -1)"
for v in w1s64b16 w1s64b8 w2s64b16 w1s96b16; do echo "cfg4 0.2 $v: $(RCC_BA_LIB=$V/librcc_ba_$v.so python tools/schur_time.py 4 0.2 3 2>/dev/null | tail -1)"; done
echo "cfg2 default: $(python tools/schur_time.py 2 1.0 10 2>/dev/null | tail -1)"
for v in w1s64b16 w1s64b8 w2s64b16 w1s96b16; do echo "cfg2 $v: $(RCC_BA_LIB=$V/librcc_ba_$v.so python tools/schur_time.py 2 1.0 10 2>/dev/null | tail -1)"; done
