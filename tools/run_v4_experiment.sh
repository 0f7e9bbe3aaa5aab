# SYRK v4 (tensor-core product, accumulators in shared memory) against v2: timings first; the full GPU suite under
# RCC_SYRK=v4 only if v4 is faster on both workloads, else the parity subset (to record that it is correct)
cd $GRAFT_REPO_ROOT
t() { RCC_SYRK=$1 python tools/schur_time.py $2 $3 $4 2>/dev/null | tail -1 | python -c "import sys,json; print(json.load(sys.stdin)['us']['schur_syrk'])"; }
a2=$(t v2 4 0.2 3); a4=$(t v4 4 0.2 3); b2=$(t v2 2 1.0 10); b4=$(t v4 2 1.0 10)
echo "cfg4/5 v2 $a2 us, v4 $a4 us; cfg2 v2 $b2 us, v4 $b4 us"
if python -c "import sys; sys.exit(0 if ($a4 < 0.97*$a2 and $b4 < 1.0*$b2) else 1)"; then
  echo "v4 faster: full GPU suite under RCC_SYRK=v4"
  RCC_SYRK=v4 timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
else
  echo "v4 not faster: parity subset under RCC_SYRK=v4"
  RCC_SYRK=v4 timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -3
fi
