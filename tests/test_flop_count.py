"""CPU: bench.py's FP64 roofline numerators are an instrumented count, not an estimate.  The counting harness
compiles csrc/model.cuh (the functions the kernels run) with `double` replaced by a counting scalar."""
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_flop_constants_equal_the_instrumented_count(tmp_path):
    exe = str(tmp_path / "flop_count")
    subprocess.run(["g++", "-O1", "-std=c++17", os.path.join(ROOT, "tests", "host_harness", "flop_count.cpp"), "-o", exe],
                   check=True)
    got = json.loads(subprocess.run([exe], check=True, capture_output=True, text=True).stdout)
    import bench
    assert bench.FLOP_EVAL_PER_BLOCK == got["single_eval_per_block"]
    assert bench.FLOP_PRODUCTS_E == got["single_products_e_pass"]
    assert bench.FLOP_PRODUCTS_ALL == got["single_products_all"]
    with open(os.path.join(ROOT, "profiles", "r2_flop_count.json")) as f:
        assert json.load(f) == got                      # the committed copy is current


def test_strong_scaling_shards_tile_the_workload():
    """bench.py --gpus N: the ranks' keyframe ranges are disjoint, contiguous and cover the N=1 problem."""
    import bench
    for cfg in (2, 4):
        for world in (1, 2, 4, 8):
            views, ranges = bench.view_ranges(cfg, world)
            assert views == bench.CONFIGS[cfg][1]
            assert ranges[0][0] == 0 and ranges[-1][1] == views
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= bench.CONFIGS[cfg][4]
    # the union of the shards is the N=1 scene, bit for bit
    import numpy as np
    full, _, _ = bench.workload(2, 0, 1, scale=0.1)
    parts = [bench.workload(2, r, 4, scale=0.1)[0] for r in range(4)]
    assert np.array_equal(np.concatenate([p.pixels for p in parts]), full.pixels)
    assert np.array_equal(np.concatenate([p.views for p in parts]), full.views)
    assert sum(p.n_blocks for p in parts) == full.n_blocks
    assert all(np.array_equal(p.markers, full.markers) for p in parts)


def test_cpu_lm_iteration_figure_of_the_bench():
    """bench.py's lm_iter.cpu: measured whole for a one-slice workload; for a sliced one the stages proportional to the
    keyframes are scaled by the slice count and the dense solve (full-size reduced system) counts once."""
    import bench
    whole = bench.cpu_lm_iter(2, 0.05)                               # 250 views, one slice
    assert whole["slices"] == 1 and whole["reduced_system_n"] == 3009 and whole["kind"] == "port"
    assert abs(whole["s_per_iter"] - sum(whole["stages_s"].values())) < 1e-12
    assert 0 < whole["s_per_iter"] <= whole["measured_s"]
    sliced = bench.cpu_lm_iter(2, 0.05, max_views=125, min_free_gb=0.5)
    assert sliced["slices"] == 2 and sliced["reduced_system_n"] == 3009
    st = sliced["stages_s"]
    assert abs(sliced["s_per_iter"] - (st["linearize"] + st["schur"] + st["backsub"] + st["solve"])) < 1e-12
    assert "scaled by the 2 slices" in sliced["sample"]
