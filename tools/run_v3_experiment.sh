set -x
cd $GRAFT_REPO_ROOT
RCC_SYRK=v3 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -5
for v in v2 v3; do echo "cfg4 0.2 $v: $(RCC_SYRK=$v python tools/schur_time.py 4 0.2 3 2>/dev/null | tail -1)"; done
echo "cfg4 0.2 v3 pass6: $(RCC_SYRK=v3 RCC_BA_LIB=$PWD/robot_camera_calibration_b200/build/variants/librcc_ba_pass6.so python tools/schur_time.py 4 0.2 3 2>/dev/null | tail -1)"
for v in v2 v3; do echo "cfg2 $v: $(RCC_SYRK=$v python tools/schur_time.py 2 1.0 10 2>/dev/null | tail -1)"; done
