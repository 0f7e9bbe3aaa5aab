"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle.

Tolerance: 1e-9 relative (per-block Frobenius) in FP64 for residuals,
Jacobians, normal equations and the reduced system -- BASELINE.json north_star.
Converged parameters: 1e-6 rad / 1e-6 m.
"""
import numpy as np
import pytest

import ba_oracle as O
from helpers import max_block_rel, oracle_blocks, oracle_reduced, rel_fro, to_oracle
from robot_camera_calibration_b200.problem import BAProblem
from robot_camera_calibration_b200.scenes import config_scene, make_scene

pytestmark = pytest.mark.gpu
TOL = 1e-9

SCENES = {
    "single": dict(n_markers=12, n_views=30, visibility=0.7, seed=11),
    "single_2cam": dict(n_markers=14, n_views=24, visibility=0.7, n_cam=2, seed=12),
    "rig": dict(n_markers=16, n_views=20, visibility=0.6, n_cam=3, model="rig", seed=13),
}


def _scene(name):
    kw = dict(SCENES[name])
    s = make_scene(kw.pop("n_markers"), kw.pop("n_views"), kw.pop("visibility"), **kw)
    if name == "single_2cam":
        # a "single" model with two cameras: every view belongs to one camera
        cam_of_view = np.arange(len(s.views)) % 2
        keep = s.cam_idx == cam_of_view[s.view_idx]
        s.view_idx, s.marker_idx, s.cam_idx, s.pixels = (s.view_idx[keep], s.marker_idx[keep],
                                                         s.cam_idx[keep], s.pixels[keep])
    return s


@pytest.mark.parametrize("name", list(SCENES))
def test_evaluate_matches_oracle(name):
    s = _scene(name)
    p = to_oracle(s)
    r = O.residuals(p)
    Jb = O.jacobian_blocks_cs(p)
    with BAProblem.from_scene(s) as gp:
        out = gp.evaluate()
    assert abs(out["cost"] - O.cost(p)) <= TOL * O.cost(p)
    assert max_block_rel(out["residuals"], r) < TOL
    for k in Jb:
        assert max_block_rel(out["jacobians"][k], Jb[k]) < TOL, k


@pytest.mark.parametrize("elim", ["views", "markers"])
@pytest.mark.parametrize("name", list(SCENES))
def test_normal_equations_match_oracle(name, elim):
    s = _scene(name)
    p = to_oracle(s)
    ob = oracle_blocks(p, elim == "views")
    with BAProblem.from_scene(s, eliminate=elim) as gp:
        cost = gp.linearize()
        nb = gp.normal_blocks()
    assert abs(cost - ob["cost"]) <= TOL * ob["cost"]
    scale_e = np.linalg.norm(ob["Hee"].reshape(len(ob["Hee"]), -1), axis=1).max()
    for k in ("Hee", "Hff", "W", "ge", "gf", "Hes", "Hfs"):
        floor = 1e-6 * np.abs(ob[k]).max()
        assert max_block_rel(nb[k], ob[k], floor=floor) < TOL, (k, scale_e)
    assert rel_fro(nb["Hss"], ob["Hss"]) < TOL
    assert rel_fro(nb["gs"], ob["gs"]) < TOL


@pytest.mark.parametrize("elim", ["views", "markers"])
@pytest.mark.parametrize("name", list(SCENES))
def test_schur_and_step_match_oracle(name, elim):
    s = _scene(name)
    p = to_oracle(s)
    radius = 1e4
    S, b, f_index, H, g, d2 = oracle_reduced(p, elim == "views", radius)
    with BAProblem.from_scene(s, eliminate=elim) as gp:
        gp.linearize()
        gp.schur(radius)
        Sg, bg = gp.reduced_system()
        mcc, step_norm, x_norm = gp.solve_step()
        st = gp.step()
        c_new = gp.candidate_cost()
    assert rel_fro(Sg, S) < TOL
    assert rel_fro(bg, b) < TOL
    delta, mcc_o, c, gmax = O.lm_step(p, radius)
    o_view, o_marker, o_shared, n = p.offsets()
    nv, nm = len(p.views), len(p.markers)
    dv, dm, ds = delta[o_view:o_marker].reshape(nv, 6), delta[o_marker:o_shared].reshape(nm, 6), delta[o_shared:]
    de, df = (dv, dm) if elim == "views" else (dm, dv)
    # the step solves a system with condition number >> 1: compare at 1e-6 relative
    assert rel_fro(st["d_e"], de) < 1e-6
    assert rel_fro(st["d_f"], df) < 1e-6
    assert rel_fro(st["d_shared"], ds) < 1e-6
    assert abs(mcc - mcc_o) <= 1e-6 * abs(mcc_o)
    assert abs(step_norm - np.linalg.norm(delta)) <= 1e-6 * np.linalg.norm(delta)
    cand = p.copy()
    cand.unpack(p.pack() + delta)
    assert abs(c_new - O.cost(cand)) <= 1e-6 * O.cost(cand)


@pytest.mark.parametrize("name,elim", [("single", "views"), ("single", "markers"), ("rig", "views")])
def test_converged_parameters_match_oracle(name, elim):
    s = _scene(name)
    p = to_oracle(s)
    O.lm_solve(p, max_iters=60)
    with BAProblem.from_scene(s, eliminate=elim) as gp:
        summ = gp.solve(max_iterations=60, function_tolerance=1e-15, gradient_tolerance=1e-12,
                        parameter_tolerance=1e-14)
        views, markers = gp.get_view_poses(), gp.get_marker_poses()
        intr, dist = gp.get_intrinsics()
        ext = gp.get_rig_extrinsics() if s.model == "rig" else None
    assert summ["final_cost"] <= summ["initial_cost"]
    assert abs(summ["final_cost"] - O.cost(p)) <= 1e-9 * O.cost(p)
    # 1e-6 rad / 1e-6 m (BASELINE.json north_star)
    assert np.abs(views - p.views).max() < 1e-6
    assert np.abs(markers - p.markers).max() < 1e-6
    assert np.abs(intr - p.intr).max() < 1e-4      # pixels
    assert np.abs(dist - p.dist).max() < 1e-6
    if ext is not None:
        assert np.abs(ext - p.ext).max() < 1e-6


def test_eval_failure_is_reported():
    s = _scene("single")
    from robot_camera_calibration_b200.scenes import compose
    # turn camera 3 around its own y axis by pi: every tag it saw is now behind it
    s.views[3] = compose(s.views[3], np.array([0.0, np.pi, 0.0, 0.0, 0.0, 0.0]))
    assert (O.depths(to_oracle(s)) <= 0).any()
    with BAProblem.from_scene(s) as gp:
        out = gp.evaluate(allow_failure=True)
    assert out["failed"]


def test_cfg1_scene_full_size():
    """BASELINE configs[0]: 20 tags x 200 views -- evaluate + linearise parity at full size."""
    s = config_scene(1)
    p = to_oracle(s)
    r = O.residuals(p)
    Jb = O.jacobian_blocks_cs(p)
    with BAProblem.from_scene(s) as gp:
        out = gp.evaluate()
        cost = gp.linearize()
        nb = gp.normal_blocks()
    assert max_block_rel(out["residuals"], r) < TOL
    for k in Jb:
        assert max_block_rel(out["jacobians"][k], Jb[k]) < TOL
    Je, Jf = Jb["view"], Jb["marker"]
    W = np.einsum('nri,nrj->nij', Je, Jf)
    assert max_block_rel(nb["W"], W, floor=1e-6 * np.abs(W).max()) < TOL
    Hee = np.zeros((len(s.views), 6, 6))
    np.add.at(Hee, s.view_idx, np.einsum('nri,nrj->nij', Je, Je))
    assert max_block_rel(nb["Hee"], Hee) < TOL
    assert abs(cost - O.cost(p)) <= TOL * O.cost(p)


from helpers import GOLDEN, load_golden  # noqa: E402


@pytest.mark.parametrize("name", GOLDEN)
def test_gpu_matches_committed_golden_vectors(name):
    """CUDA path vs tests/golden/*.npz (OpenCV analytic derivatives for the
    residuals / Jacobians, dense numpy for normal equations, Schur and step)."""
    s, g = load_golden(name)
    for elim in ("views", "markers"):
        with BAProblem.from_scene(s, eliminate=elim) as gp:
            out = gp.evaluate()
            assert np.abs(out["residuals"] - g["residuals"]).max() < 1e-9
            for k, J in out["jacobians"].items():
                assert max_block_rel(J, g[f"jac_{k}"]) < TOL, k
            cost = gp.linearize()
            assert abs(cost - float(g["cost"])) <= TOL * float(g["cost"])
            nb = gp.normal_blocks()
            for k in ("Hee", "ge", "Hes", "Hff", "gf", "Hfs", "Hss", "gs", "W"):
                assert rel_fro(nb[k], g[f"{elim}_{k}"]) < TOL, (elim, k)
            gp.schur(float(g["radius"]))
            S, b = gp.reduced_system()
            assert rel_fro(S, g[f"{elim}_S"]) < TOL
            assert rel_fro(b, g[f"{elim}_b"]) < TOL
            mcc, step_norm, _ = gp.solve_step()
            assert abs(step_norm - np.linalg.norm(g["delta"])) <= 1e-6 * np.linalg.norm(g["delta"])
            assert abs(mcc - float(g["mcc"])) <= 1e-6 * abs(float(g["mcc"]))


@pytest.mark.parametrize("loss,scale", [("huber", 2.0), ("cauchy", 3.0)])
@pytest.mark.parametrize("name", ["single", "rig"])
def test_robust_loss_matches_oracle(name, loss, scale):
    """Huber / Cauchy on each tag's 8-residual block (Ceres Corrector with alpha = 0)."""
    s = _scene(name)
    rng = np.random.default_rng(5)
    bad = rng.choice(s.n_blocks, 12, replace=False)
    s.pixels[bad] += rng.normal(0, 25.0, (12, 8))            # gross outliers
    p = to_oracle(s)
    p.loss, p.loss_scale = loss, scale
    ob = oracle_blocks(p, True)
    S, b, *_ = oracle_reduced(p, True, 1e4)
    with BAProblem.from_scene(s, eliminate="views") as gp:
        gp.set_loss(loss, scale)
        out = gp.evaluate(want_jacobians=False)
        cost = gp.linearize()
        nb = gp.normal_blocks()
        gp.schur(1e4)
        Sg, bg = gp.reduced_system()
    assert abs(out["cost"] - ob["cost"]) <= TOL * ob["cost"]          # 0.5 * sum rho(s)
    assert abs(cost - ob["cost"]) <= TOL * ob["cost"]
    for k in ("Hee", "Hff", "W", "ge", "gf", "Hes", "Hfs"):
        assert max_block_rel(nb[k], ob[k], floor=1e-6 * np.abs(ob[k]).max()) < TOL, k
    assert rel_fro(nb["Hss"], ob["Hss"]) < TOL and rel_fro(nb["gs"], ob["gs"]) < TOL
    assert rel_fro(Sg, S) < TOL and rel_fro(bg, b) < TOL


def test_huber_resists_outliers():
    s = make_scene(10, 30, 0.8, seed=23, pixel_noise=0.1)
    rng = np.random.default_rng(1)
    bad = rng.choice(s.n_blocks, s.n_blocks // 20, replace=False)
    s.pixels[bad] += rng.normal(0, 40.0, (len(bad), 8))
    err = {}
    for loss in ("trivial", "huber"):
        with BAProblem.from_scene(s) as gp:
            gp.set_loss(loss, 1.0)
            gp.solve(max_iterations=60)
            err[loss] = np.abs(gp.get_marker_poses()[:, 3:] - s.truth["markers"][:, 3:]).max()
    assert err["huber"] < 0.5 * err["trivial"]


def test_converged_parameters_match_an_independent_solver():
    """The GPU LM (Schur + trust region written here) and SciPy's trust-region-reflective least-squares
    solver (driven by the oracle's residuals and complex-step Jacobian) must find the same minimiser:
    1e-6 rad / 1e-6 m on poses (BASELINE.json north_star).  Pins the LM driver to a solver that shares
    no code with it."""
    from scipy.optimize import least_squares
    s = make_scene(8, 14, 0.9, seed=61)
    p = to_oracle(s)
    free = ~p.const_mask()
    x0 = p.pack()

    def with_x(xf):
        q = p.copy()
        x = x0.copy()
        x[free] = xf
        q.unpack(x)
        return q

    sol = least_squares(lambda xf: O.residuals(with_x(xf)).ravel(), x0[free],
                        jac=lambda xf: O.dense_jacobian(with_x(xf))[:, free], method="trf", x_scale="jac",
                        ftol=1e-15, xtol=1e-15, gtol=1e-13, max_nfev=200)
    ref = with_x(sol.x)
    with BAProblem.from_scene(s) as gp:
        summ = gp.solve(max_iterations=100, function_tolerance=1e-16, gradient_tolerance=1e-13,
                        parameter_tolerance=1e-15)
        views, markers = gp.get_view_poses(), gp.get_marker_poses()
        intr, dist = gp.get_intrinsics()
    assert abs(summ["final_cost"] - O.cost(ref)) <= 1e-9 * O.cost(ref)
    assert np.abs(views - ref.views).max() < 1e-6
    assert np.abs(markers - ref.markers).max() < 1e-6
    assert np.abs(dist - ref.dist).max() < 1e-6
    assert np.abs(intr - ref.intr).max() < 1e-4      # pixels


def test_jacobi_scaling_changes_only_the_clamped_diagonal():
    """rcc_lm_options.jacobi_scaling (Ceres's default): the LM diagonal clamp(s^2 diag J^T J) / (radius s^2),
    s = 1 / (1 + ||column||), equals diag(J^T J) / radius wherever the clamp is inactive, so the damped step and the
    minimiser must not move; where min_diagonal does bite (a parameter the data barely constrain) the scaled
    diagonal is the one the oracle formula gives."""
    s = make_scene(12, 30, 0.7, seed=33)
    p = to_oracle(s)
    H, g, _ = O.normal_equations(p)
    steps = {}
    for jac in (False, True):
        with BAProblem.from_scene(s, eliminate="views") as gp:
            gp.set_lm_diagonal(1e-6, 1e32, jac)
            gp.linearize(); gp.schur(1e3); gp.solve_step()
            st = gp.step()
            steps[jac] = np.concatenate([st["d_e"].ravel(), st["d_f"].ravel(), st["d_shared"]])
    hd = np.diag(H)
    sc = 1.0 / (1.0 + np.sqrt(hd))
    assert (hd * sc * sc).min() > 1e-6                      # no clamp on this scene: identical damping
    assert rel_fro(steps[True], steps[False]) < 1e-9
    # converged parameters with and without scaling (Ceres-like defaults on both sides)
    out = {}
    for jac in (0, 1):
        with BAProblem.from_scene(s) as gp:
            summ = gp.solve(max_iterations=40, function_tolerance=1e-14, gradient_tolerance=1e-12,
                            parameter_tolerance=1e-13, jacobi_scaling=jac)
            out[jac] = (gp.get_view_poses(), gp.get_marker_poses(), gp.get_intrinsics(), summ)
    assert np.abs(out[0][0] - out[1][0]).max() < 1e-8 and np.abs(out[0][1] - out[1][1]).max() < 1e-8
    assert abs(out[0][3]["final_cost"] - out[1][3]["final_cost"]) <= 1e-12 * out[0][3]["final_cost"]
    # a clamp that bites: min_diagonal = 10 exceeds s^2 h (< 1) everywhere, so with scaling every parameter is damped
    # by 10 / (radius s^2) = 10 (1 + sqrt(h))^2 / radius; the oracle's dense solve with that diagonal must agree
    radius = 1e3
    with BAProblem.from_scene(s, eliminate="views") as gp:
        gp.set_lm_diagonal(10.0, 1e32, True)
        gp.linearize(); gp.schur(radius); gp.solve_step()
        st = gp.step()
    cm = p.const_mask()
    d2 = 10.0 * (1.0 + np.sqrt(hd)) ** 2 / radius
    Hd, gm = O.masked_system(H + np.diag(d2), g, cm)
    want = -np.linalg.solve(Hd, gm)
    want[cm] = 0
    o_view, o_marker, o_shared, n = p.offsets()
    got = np.concatenate([st["d_e"].ravel(), st["d_f"].ravel(), st["d_shared"]])
    assert rel_fro(got, want) < 1e-7
