#!/bin/bash
# Round-2 measurement set on one GPU (run under gpurun from the repo root); everything lands in gpurun_out/.
#   1. bench.py --impl reference, bench.py (the two arms, default workload)
#   2. ncu launch list of the same bench command (gpu__time_duration per launch)
#   3. (FULL=1) ncu --set full of the dominant kernels on the bench workload: K2 passes (the bench's roofline
#      kernel), K1, K3b SYRK
set -x
O=gpurun_out
python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err
python bench.py --steps 20 --warmup 5 > $O/r2_bench.json 2> $O/r2_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cfg2 --cpu-seconds 1 > $O/r2_bench_under_ncu.json 2> /dev/null
python tools/launch_summary.py $O/r2_launches.csv > $O/r2_launch_summary.txt
if [ "$FULL" = "1" ]; then
ncu --set full --import-source on --clock-control none -k regex:assemble_kernel -s 6 -c 2 -f -o $O/r2_ncu_k2_cfg4 \
    python tools/k2_time.py 4 1.0 3 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:materialise_kernel -s 1 -c 1 -f -o $O/r2_ncu_k1_cfg4 \
    python -c "
import sys; sys.path.insert(0, '.')
from bench import workload
from robot_camera_calibration_b200.problem import BAProblem
s, _, _ = workload(4, 0, 1, 0.5)
gp = BAProblem.from_scene(s, eliminate='views')
for _ in range(3): gp.evaluate_device(want_jacobians=True)
gp.synchronize()
" > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:schur_syrk -s 1 -c 1 -f -o $O/r2_ncu_syrk_cfg4 \
    python tools/schur_time.py 4 0.2 2 > /dev/null 2>&1
for r in r2_ncu_k2_cfg4 r2_ncu_k1_cfg4 r2_ncu_syrk_cfg4; do
  python tools/ncu_summary.py $O/$r.ncu-rep > $O/${r}_summary.txt 2>&1
done
fi
ls -la $O
