#!/usr/bin/env python
"""Benchmark of the BA hot path (BASELINE.json metric) -- one JSON line on stdout.

  python bench.py --gpus N --steps K --warmup W          (our CUDA path; N>1 under torchrun)
  python bench.py --impl reference ...                   (the Ceres-equivalent CPU restatement)

step      = one pass of the hot path over the batch: batched reprojection residual
            + analytic Jacobian + Gauss-Newton normal-equation assembly
            (rcc_ba_linearize: expand, assemble E pass, assemble F pass, finalize).
workload  = BASELINE.json configs[1]: 1 camera, 500 tags, 5 000 views, 20 % visibility
            (~0.47 M observation blocks = ~1.9 M corner observations) per GPU; N>1 is
            weak scaling: every rank owns another 5 000 views of the same tag cloud
            (observations shard by eliminated-block owner).
value     = corner observations / s, whole job, inputs resident in HBM, L2 flushed
            between timed iterations, device time (CUDA events), max over ranks.
e2e       = same metric through the C ABI with HOST buffers: per step H2D of the pixel
            batch + all parameter blocks, linearize, D2H of cost + gradient.
lm_iter   = seconds per full LM iteration (linearize + Schur + [all-reduce] +
            Cholesky + back-substitution + candidate cost).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "BA corner observations/s (residual+Jacobian+normal equations)"
UNIT = "observations/s"
# algorithmic work per observation block
BYTES_PER_BLOCK_E = 64 + 4 + 288          # pixels + other index in, cross block W out
BYTES_PER_BLOCK_F = 64 + 4
FLOP_PER_BLOCK = 2 * 8 * 253 + 1400       # J^T J products (253 unique entries x 8 rows) + Jacobian evaluation
FLOP_PER_BLOCK_E = 2 * 8 * 172 + 1400     # E pass alone: OO, OT, O x [S r], [S r] x [S r] (172 entries) + evaluation
NCU_DRAM_BYTES_E_PASS = 32313856 + 89391616   # measured once under ncu (cfg2, 454 996 blocks)


def workload(cfg, rank, scale):
    from robot_camera_calibration_b200.scenes import make_scene
    if cfg == 2:
        n_views = max(8, int(5000 * scale))
        s = make_scene(500, n_views, 0.20, seed=20242, view_seed=rank, name="cfg2")
        desc = f"cfg2: 1 camera, 500 tags, {n_views} views/GPU, 20% visibility"
    elif cfg == 1:
        s = make_scene(20, 200, 1.0, seed=20241, view_seed=rank, name="cfg1")
        desc = "cfg1: 1 camera, 20 tags, 200 views/GPU"
    elif cfg == 4:
        n_views = max(8, int(10000 * scale))
        s = make_scene(5000, n_views, 0.25, seed=20244, view_seed=rank, name="cfg4")
        desc = f"cfg4: 1 camera, 5000 tags, {n_views} views/GPU, 25% visibility"
    else:
        raise SystemExit(f"unsupported --config {cfg}")
    return s, desc


def bind_near_gpu(torch, local):
    """Best effort: run this rank on the CPUs NVML calls ideal for its GPU, so that first-touch puts the
    pinned host buffers of the end-to-end path on the GPU's NUMA node (a remote node halves H2D bandwidth)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return {"cpus_before": before, "cpus_after": len(os.sched_getaffinity(0))}
    except Exception as e:  # not permitted in this cgroup, NVML missing, ...
        return {"unchanged": str(e)[:100]}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(scene, seconds=12.0, max_views=400):
    """The Ceres-equivalent CPU restatement (oracle/) timed on a bounded sample."""
    if "cpu_baseline" not in sys.modules:
        ncpu = len(os.sched_getaffinity(0))
        os.environ["OMP_NUM_THREADS"] = str(ncpu)          # before libgomp is loaded
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from cpu_baseline import CpuBA
    keep = scene.view_idx < max_views
    import copy
    s = copy.copy(scene)
    s.views = scene.views[:max_views]
    s.const_views = scene.const_views[:max_views]
    s.view_idx, s.marker_idx, s.cam_idx, s.pixels = (scene.view_idx[keep], scene.marker_idx[keep],
                                                     scene.cam_idx[keep], scene.pixels[keep])
    cpu = CpuBA(s, eliminate="views")
    cpu.linearize()                                   # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        cpu.linearize()
        n += 1
        el = time.perf_counter() - t0
        if el > seconds or n >= 200:
            break
    obs = 4 * len(s.view_idx)
    tm = {}
    cpu.lm_iteration(1e4, timings=tm)
    return {"value": obs * n / el, "unit": UNIT, "cores": cpu.threads, "kind": "port",
            "sample": f"first {max_views} views of the workload ({obs} observations), {n} passes in {el:.1f} s; "
                      "C++/OpenMP Jet-autodiff restatement of the Ceres evaluation (Ceres itself is not available)",
            "lm_iter_s_sample": sum(tm.values()), "lm_iter_breakdown_s": tm}, s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1; the reference arm may use every host thread it can get
    ncpu = len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(ncpu)
    os.environ["OPENBLAS_NUM_THREADS"] = str(ncpu)
    scene, desc = workload(args.config, 0, args.scale)
    base, s = cpu_baseline(scene, seconds=1.0, max_views=args.ref_views)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from cpu_baseline import CpuBA
    cpu = CpuBA(s, eliminate="views")
    for _ in range(max(3, args.warmup)):      # same floor as our arm: a cold OpenMP pool must not be timed
        cpu.linearize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu.linearize()
    el = time.perf_counter() - t0
    obs = 4 * len(s.view_idx)
    v = obs * args.steps / el
    out = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": max(3, args.warmup),
           "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "impl": "reference",
           "config": {"workload": desc, "sample": f"each step = first {args.ref_views} views ({obs} observations)"},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cpu.threads, "kind": "port",
                            "sample": f"first {args.ref_views} views ({obs} observations) per step; Ceres-equivalent "
                                      "C++/OpenMP restatement (Ceres and the reference optimiser do not exist here)"},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(out)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from robot_camera_calibration_b200.problem import BAProblem, fp64_peak_tflops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    cpus_all = os.sched_getaffinity(0)
    placement = bind_near_gpu(torch, local)      # pinned host buffers must sit on the GPU's NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    scene, desc = workload(args.config, rank, args.scale)
    n_blocks, n_obs = scene.n_blocks, scene.n_observations
    gp = BAProblem.from_scene(scene, device=local, eliminate="views")
    stream = torch.cuda.Stream(device=local)
    gp._check(gp.lib.rcc_ba_set_stream(gp.h, stream.cuda_stream))
    if world > 1:
        ids = [BAProblem.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        gp.comm_init(ids[0], rank, world)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # ---------------- device-resident steps (value) -------------------------------------
    for _ in range(max(3, args.warmup)):
        gp.linearize(want_cost=False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    gp.profile_reset()
    gp.profile_enable(True)
    l0 = gp.launch_count()
    pairs = []
    barrier()
    for _ in range(args.steps):
        gp.flush_l2()                                     # evict L2 between timed iterations (untimed)
        a, b = ev(), ev()
        a.record(stream)
        gp.linearize(want_cost=False)
        b.record(stream)
        pairs.append((a, b))
    barrier()
    launches = gp.launch_count() - l0 - args.steps       # minus the flush kernels
    step_ms = [a.elapsed_time(b) for a, b in pairs]
    total_ms = float(sum(step_ms))
    prof = gp.profile()
    gp.profile_enable(False)

    # ---------------- materialised evaluation K1 (HBM-write-bound kernel) ----------------
    for _ in range(2):
        gp.evaluate_device(want_jacobians=True)
    barrier()
    gp.profile_reset()
    gp.profile_enable(True)
    mat_steps = max(3, min(args.steps, 10))
    for _ in range(mat_steps):
        gp.flush_l2()
        gp.evaluate_device(want_jacobians=True)
    barrier()
    mat_ms = gp.profile()["evaluate"][0] / mat_steps
    gp.profile_enable(False)

    # ---------------- full LM iterations (lm_iter) --------------------------------------
    def lm_iteration():
        gp.linearize(want_cost=False)
        gp.schur(1e4)
        gp.solve_step()
        gp.candidate_cost()

    views0, markers0 = scene.views.copy(), scene.markers.copy()
    lm_iteration()
    barrier()
    gp.profile_reset()
    gp.profile_enable(True)
    lm_n = max(2, min(args.steps, 5))
    a, b = ev(), ev()
    a.record(stream)
    for _ in range(lm_n):
        lm_iteration()
    b.record(stream)
    barrier()
    lm_ms = a.elapsed_time(b) / lm_n
    lm_prof = {k: v[0] / lm_n for k, v in gp.profile().items() if v[0] > 0}
    gp.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- end-to-end steps through the C ABI with host buffers --------------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    h_pix, h_views, h_markers = pin(scene.pixels), pin(views0), pin(markers0)
    h_intr, h_dist = pin(scene.intr), pin(scene.dist)
    d = gp.dims
    h_ge, h_gf, h_gs = pin(np.zeros((d.n_e, 6))), pin(np.zeros((d.n_f, 6))), pin(np.zeros(d.n_shared))
    from robot_camera_calibration_b200.problem import _dp

    def e2e_step():
        # parameters first: the pixel upload (99 % of the bytes) then overlaps with the E pass piece by piece
        gp.set_view_poses(h_views)
        gp.set_marker_poses(h_markers)
        gp.set_intrinsics(h_intr, h_dist)
        gp.update_pixels(h_pix)
        cost = gp.linearize(want_cost=True)
        gp._check(gp.lib.rcc_ba_get_normal_blocks(gp.h, None, _dp(h_ge), None, None, _dp(h_gf), None, None,
                                                  _dp(h_gs), None))
        return cost

    for _ in range(2):
        e2e_step()
    barrier()
    d_probe = torch.empty(h_pix.size, dtype=torch.float64, device="cuda")
    t_probe = torch.from_numpy(h_pix)
    h2d_ms = []
    for _ in range(3):
        a, b = ev(), ev()
        a.record(); d_probe.view(t_probe.shape).copy_(t_probe, non_blocking=True); b.record()
        torch.cuda.synchronize()
        h2d_ms.append(a.elapsed_time(b))
    h2d_gbs = h_pix.nbytes / (min(h2d_ms) * 1e-3) / 1e9
    del d_probe
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = h_pix.nbytes + h_views.nbytes + h_markers.nbytes + h_intr.nbytes + h_dist.nbytes
    d2h = 8 + h_ge.nbytes + h_gf.nbytes + h_gs.nbytes

    # ---------------- reduce over ranks --------------------------------------------------
    def rmax(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def rsum(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    total_ms_max = rmax(total_ms)
    obs_all = rsum(n_obs)
    lm_ms_max = rmax(lm_ms)
    e2e_s_max = rmax(e2e_s)
    launches_all = int(rsum(launches))

    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        fp64_peak = fp64_peak_tflops(local)
        k_ms = prof["assemble_e"][0] / max(1, prof["assemble_e"][1])
        kf_ms = prof["assemble_f"][0] / max(1, prof["assemble_f"][1])
        ach = BYTES_PER_BLOCK_E * n_blocks / (k_ms * 1e-3) / 1e9
        value = obs_all * args.steps / (total_ms_max * 1e-3)
        step_flops = FLOP_PER_BLOCK * n_blocks
        os.sched_setaffinity(0, cpus_all)        # the CPU baseline may use every host core again
        base = cpu_baseline(scene, seconds=args.cpu_seconds)[0] if world == 1 else None
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "observation_blocks_per_gpu": n_blocks, "observations_per_gpu": n_obs,
                       "eliminated": "views", "l2": "flushed (256 MiB write) between timed iterations, untimed",
                       "timing": "CUDA events per step on the library's stream, summed; max over ranks"},
            "roofline": {"bound": "hbm", "kernel": "assemble_kernel<E pass> (fused residual+Jacobian+J^T J tiles)",
                         "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                         "traffic": NCU_DRAM_BYTES_E_PASS if (args.config == 2 and args.scale == 1.0) else None,
                         "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one E-pass "
                                           "launch on this workload (profiles/r1_ncu_full_summary.txt); below the "
                                           "algorithmic bytes because part of W is still in L2 when the kernel ends",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": BYTES_PER_BLOCK_E * n_blocks,
                         "kernel_ms": k_ms, "f_pass_kernel_ms": kf_ms,
                         "note": "fused assembly is FP64-pipe-bound, not HBM-bound: see roofline_fp64"},
            "roofline_fp64_e_pass": {"bound": "fp64 (DFMA + DMMA.8x8x4 share one pipe; same peak either way)",
                                     "kernel": "assemble_kernel<E pass>",
                                     "achieved": FLOP_PER_BLOCK_E * n_blocks / (k_ms * 1e-3) / 1e12, "peak": fp64_peak,
                                     "unit": "TFLOP/s", "frac": FLOP_PER_BLOCK_E * n_blocks / (k_ms * 1e-3) / 1e12 / fp64_peak,
                                     "algorithmic_flop_per_launch": FLOP_PER_BLOCK_E * n_blocks,
                                     "fp64_pipe_busy_ncu": 0.64,
                                     "note": "executed work is larger: 8 DMMA x 512 flop + 4 corner evaluations per block; "
                                             "ncu sm__throughput (FP64 pipe) 64 % (profiles/r1_ncu_full_summary.txt)"},
            "roofline_fp64": {"bound": "fp64", "achieved": step_flops / (total_ms / args.steps * 1e-3) / 1e12,
                              "peak": fp64_peak, "unit": "TFLOP/s",
                              "frac": step_flops / (total_ms / args.steps * 1e-3) / 1e12 / fp64_peak,
                              "peak_source": "measured here: DFMA microbenchmark (rcc_fp64_peak_tflops)",
                              "algorithmic_flop_per_step": step_flops},
            "roofline_materialise": {"bound": "hbm", "kernel": "evaluate_kernel (K1: residuals + Ceres-layout Jacobians to HBM)",
                                     "achieved": 1480 * n_blocks / (mat_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": 1480 * n_blocks / (mat_ms * 1e-3) / 1e9 / hbm_peak,
                                     "algorithmic_bytes_per_launch": 1480 * n_blocks, "kernel_ms": mat_ms,
                                     "observations_per_s": n_obs / (mat_ms * 1e-3),
                                     "note": "includes the 1-CTA cost reduction launched behind it"},
            "cpu_baseline": base,
            "e2e": {"value": obs_all * args.steps / e2e_s_max, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s_max / args.steps,
                    "h2d_gb_per_s_this_box": h2d_gbs, "cpu_placement": placement,
                    "what": "set_view/marker_poses + set_intrinsics + update_pixels (pinned host -> device), "
                            "linearize, read back cost and gradient"},
            "gpu_launches": launches_all,
            "clocks": clocks,
            "lm_iter": {"s_per_iter": lm_ms_max * 1e-3, "stage_ms": lm_prof,
                        "reduced_system_n": int(d.n_reduced), "n_pairs": int(d.n_pairs)},
            "stage_ms_per_step": {k: v[0] / args.steps for k, v in prof.items() if v[0] > 0},
        }
        _emit(out)
    gp.close()
    if world > 1:
        dist.destroy_process_group()


def _emit(obj):
    """The one JSON line goes to the process's ORIGINAL stdout; everything else any
    library prints (NCCL banners, warnings) was re-routed to stderr in main()."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                      # fd 1 -> stderr for the rest of the run
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the config's views (debug)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-views", type=int, default=400)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
