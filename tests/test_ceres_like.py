"""The Ceres-shaped front end (Problem / AddResidualBlock / SetParameterBlockConstant /
Solve / CostFunction::Evaluate) over the GPU path -- reads like a Ceres user's test."""
import numpy as np
import pytest

import ba_oracle as O
from helpers import to_oracle
from robot_camera_calibration_b200.ceres_like import Problem, Solve, SolverOptions, TagReprojectionCost
from robot_camera_calibration_b200.scenes import make_scene


def _build(s):
    intr, dist = s.intr[0].copy(), s.dist[0].copy()
    views = [v.copy() for v in s.views]
    markers = [m.copy() for m in s.markers]
    problem = Problem()
    for b in range(s.n_blocks):
        cost = TagReprojectionCost(s.pixels[b], s.sizes[s.marker_idx[b]])
        problem.AddResidualBlock(cost, None, intr, dist, views[s.view_idx[b]], markers[s.marker_idx[b]])
    problem.SetParameterBlockConstant(markers[0])          # the world tag, camera_pose.cpp:71-80
    return problem, intr, dist, views, markers


def test_api_shape_and_argument_checks():
    s = make_scene(5, 6, 0.9, seed=2)
    problem, intr, dist, views, markers = _build(s)
    assert problem.NumResidualBlocks() == s.n_blocks and problem.NumResiduals() == 8 * s.n_blocks
    cost = TagReprojectionCost(s.pixels[0], 0.1)
    assert cost.num_residuals() == 8 and cost.parameter_block_sizes() == [4, 5, 6, 6]
    with pytest.raises(ValueError):
        problem.AddResidualBlock(cost, None, intr, dist, views[0])                 # missing block
    with pytest.raises(ValueError):
        problem.AddResidualBlock(cost, None, intr, dist, views[0], np.zeros(5))    # wrong size
    with pytest.raises(TypeError):
        problem.AddResidualBlock(cost, None, intr, dist, views[0], [0.0] * 6)      # not an array
    with pytest.raises(ValueError):
        problem.SetParameterBlockConstant(np.zeros(6))                             # unknown block


@pytest.mark.gpu
def test_cost_function_evaluate_layout():
    s = make_scene(5, 6, 0.9, seed=2)
    p = to_oracle(s)
    Jb = O.jacobian_blocks_cs(p)
    r = O.residuals(p)
    b = 3
    cost = TagReprojectionCost(s.pixels[b], s.sizes[s.marker_idx[b]])
    params = [s.intr[0].copy(), s.dist[0].copy(), s.views[s.view_idx[b]].copy(), s.markers[s.marker_idx[b]].copy()]
    residuals = np.zeros(8)
    jac = [np.zeros(8 * 4), np.zeros(8 * 5), None, np.zeros(8 * 6)]          # jacobians[2] == nullptr
    assert cost.Evaluate(params, residuals, jac)
    assert np.abs(residuals - r[b]).max() < 1e-9
    assert np.allclose(jac[0].reshape(8, 4), Jb["intr"][b], rtol=1e-9, atol=1e-9)   # row-major 8 x 4
    assert np.allclose(jac[1].reshape(8, 5), Jb["dist"][b], rtol=1e-9, atol=1e-9)
    assert np.allclose(jac[3].reshape(8, 6), Jb["marker"][b], rtol=1e-9, atol=1e-9)
    assert cost.Evaluate(params, residuals, None)                           # residual-only call
    params[2][0:3] = 0.0                                                    # camera axis = world +z ...
    params[2][3:6] = params[3][3:6] + [0, 0, 1.0]                           # ... and the tag is 1 m behind it
    assert cost.Evaluate(params, residuals, None) is False                  # Evaluate() == false


@pytest.mark.gpu
def test_solve_refines_the_callers_arrays_in_place():
    s = make_scene(8, 14, 0.9, seed=9, pixel_noise=0.0)
    problem, intr, dist, views, markers = _build(s)
    cost0, res, _ = problem.Evaluate()
    assert cost0 == pytest.approx(0.5 * float(res @ res))
    summary = Solve(SolverOptions(max_num_iterations=40, function_tolerance=1e-15, gradient_tolerance=1e-12,
                                  parameter_tolerance=1e-14), problem)
    assert summary.IsSolutionUsable() and "final cost" in summary.BriefReport()
    assert summary.final_cost < 1e-12 * summary.initial_cost
    assert np.all(markers[0] == 0.0)                                        # constant block untouched
    assert np.abs(np.stack(views) - s.truth["views"]).max() < 1e-6
    assert np.abs(np.stack(markers) - s.truth["markers"]).max() < 1e-6
    assert np.abs(dist - s.truth["dist"][0]).max() < 1e-6


def test_one_parameter_block_cannot_belong_to_two_gpu_cameras():
    """A Ceres parameter block is one array: an intrinsics array shared by residual blocks whose distortion arrays
    differ would become two independently optimised GPU cameras -- rejected instead of silently diverging; so is
    a mix of rig and single-camera cost functions.  (Raised before any GPU object is created.)"""
    s = make_scene(5, 6, 0.9, seed=3)
    intr, dist_a, dist_b = s.intr[0].copy(), s.dist[0].copy(), s.dist[0].copy()
    views = [v.copy() for v in s.views]
    markers = [m.copy() for m in s.markers]
    problem = Problem()
    for b in range(s.n_blocks):
        cost = TagReprojectionCost(s.pixels[b], s.sizes[s.marker_idx[b]])
        problem.AddResidualBlock(cost, None, intr, dist_a if b % 2 else dist_b, views[s.view_idx[b]],
                                 markers[s.marker_idx[b]])
    with pytest.raises(NotImplementedError):
        problem._build()
