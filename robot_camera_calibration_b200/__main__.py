"""Command line of the optimisation stage the reference pipeline is built around and does not ship ("Milestone 3",
real_preprocessing/README.md): read the `detections/` directory camera_pose_node leaves behind (targets.yaml +
detections_N.yaml with initial world_T_camera / world_T_target guesses, camera_pose.cpp:83-129, plus camera.yaml),
run the GPU bundle adjustment, and write the refined poses back in the same formats, where opt_vis_node
(opt_visualization.cpp:162-220) picks them up.

    python -m robot_camera_calibration_b200 <detections directory> [--fix-intrinsics] [--device 0] ...
"""
from __future__ import annotations

import argparse
import json
import sys

# rcc_lm_summary.termination (include/rcc_ba.h)
TERMINATION = {0: "no_convergence (iteration limit)", 1: "convergence (function tolerance)",
               2: "convergence (gradient tolerance)", 3: "convergence (parameter tolerance)", 4: "failure"}


def parser():
    ap = argparse.ArgumentParser(prog="python -m robot_camera_calibration_b200", description=__doc__.split("\n\n")[0])
    ap.add_argument("directory", help="the detections directory (camera_pose.cpp:41-51: <package>/detections)")
    ap.add_argument("--device", type=int, default=0, help="CUDA device")
    ap.add_argument("--fix-intrinsics", action="store_true",
                    help="keep camera.yaml's intrinsics and distortion constant (default: refined, written to "
                         "camera_refined.yaml)")
    ap.add_argument("--max-iterations", type=int, default=50)
    ap.add_argument("--function-tolerance", type=float, default=None)
    ap.add_argument("--gradient-tolerance", type=float, default=None)
    ap.add_argument("--parameter-tolerance", type=float, default=None)
    ap.add_argument("--precision", default=None,
                    help="'repr' writes poses with full round-trip precision instead of the reference's stream format")
    ap.add_argument("--verbose", action="store_true", help="one line per LM iteration on stderr")
    return ap


def main(argv=None):
    args = parser().parse_args(argv)
    from . import io_yaml
    opts = {"max_iterations": args.max_iterations, "verbose": int(args.verbose)}
    for k in ("function_tolerance", "gradient_tolerance", "parameter_tolerance"):
        if getattr(args, k) is not None:
            opts[k] = getattr(args, k)
    try:
        summary = io_yaml.optimise_directory(args.directory, device=args.device,
                                             refine_intrinsics=not args.fix_intrinsics, precision=args.precision, **opts)
    except (FileNotFoundError, ValueError) as e:          # missing camera.yaml / targets.yaml, empty dataset
        print(f"error: {e}", file=sys.stderr)
        return 2
    summary["termination_name"] = TERMINATION.get(summary["termination"], "?")
    print(json.dumps(summary))
    return 0 if summary["termination"] != 4 else 1


if __name__ == "__main__":
    sys.exit(main())
