"""CPU oracle for the bundle-adjustment hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module; the product package
(``robot_camera_calibration_b200``) never does.

PARITY STATUS: **unpinned at the Ceres boundary** -- the mounted reference holds
no cost functor, no optimiser, no tests and no golden vectors (SURVEY.md section 0,
8c).  What *is* pinned: the projection model and its first derivatives are
checked against OpenCV's own ``cv2.projectPoints`` / ``cv2.composeRT`` /
``cv2.Rodrigues`` analytic derivatives (oracle B below; cv2 4.13.0 here, the
reference links OpenCV 3.4.4 -- real_preprocessing/README.md:40), the library the
reference calls at real_preprocessing/src/camera_pose.cpp:163.

Conventions restated from the reference (all citations relative to
/root/reference/real_preprocessing/src/):
  * intrinsics  K = [fx 0 cx; 0 fy cy; 0 0 1] row-major, dist = (k1,k2,p1,p2,k3)
    -- camera_pose.cpp:38-39, 55-68
  * poses are (Rodrigues rvec, translation) of world_T_camera / world_T_target
    -- camera_pose.cpp:88-98, 111-121;   world_T_cam = world_T_tag * tag_T_cam :184
  * object points of a tag of size s, order bl, br, tr, tl:
    (-s/2,-s/2,0) (s/2,-s/2,0) (s/2,s/2,0) (-s/2,s/2,0) -- camera_pose.cpp:123-126,158-161
  * gauge: first tag of frame 0 is the world tag, pose identity, never updated
    -- camera_pose.cpp:71-80
  * an observation block = one tag in one frame = 4 corners = 8 residuals
    -- corner_detections.cpp:34,51

Oracle A: closed-form projection in NumPy + complex-step derivative (h=1e-30).
Oracle B: OpenCV analytic derivative chain.
Normal equations / Schur / LM: dense NumPy.
"""
from __future__ import annotations

import numpy as np

CS_H = 1e-30  # complex-step size


# --------------------------------------------------------------------------
# problem container (plain numpy; no dependency on the product package)
# --------------------------------------------------------------------------
class OracleProblem:
    """Plain-numpy description of one BA problem.

    model   : "single" (views are world_T_camera) or "rig" (views are
              world_T_body and ext[c] is body_T_cam_c)
    intr    : (n_cam,4) fx fy cx cy      dist : (n_cam,5) k1 k2 p1 p2 k3
    ext     : (n_cam,6) body_T_cam (rig only, else zeros)
    views   : (n_views,6)  markers : (n_markers,6)  sizes : (n_markers,)
    view_idx, marker_idx, cam_idx : (N,) int     pixels : (N,8) u0 v0 .. u3 v3
    const_* : boolean masks of parameter blocks held constant
    """

    def __init__(self, model, intr, dist, ext, views, markers, sizes,
                 view_idx, marker_idx, cam_idx, pixels,
                 const_views=None, const_markers=None, const_intr=None,
                 const_dist=None, const_ext=None):
        self.model = model
        self.intr = np.array(intr, dtype=np.float64).reshape(-1, 4)
        self.dist = np.array(dist, dtype=np.float64).reshape(-1, 5)
        self.n_cam = self.intr.shape[0]
        self.ext = (np.zeros((self.n_cam, 6)) if ext is None
                    else np.array(ext, dtype=np.float64).reshape(-1, 6))
        self.views = np.array(views, dtype=np.float64).reshape(-1, 6)
        self.markers = np.array(markers, dtype=np.float64).reshape(-1, 6)
        self.sizes = np.array(sizes, dtype=np.float64).reshape(-1)
        self.view_idx = np.asarray(view_idx, dtype=np.int64)
        self.marker_idx = np.asarray(marker_idx, dtype=np.int64)
        self.cam_idx = (np.zeros_like(self.view_idx) if cam_idx is None
                        else np.asarray(cam_idx, dtype=np.int64))
        self.pixels = np.array(pixels, dtype=np.float64).reshape(-1, 8)
        nv, nm, nc = len(self.views), len(self.markers), self.n_cam
        z = lambda a, n: (np.zeros(n, bool) if a is None else np.asarray(a, bool).copy())
        self.const_views = z(const_views, nv)
        self.const_markers = z(const_markers, nm)
        self.const_intr = z(const_intr, nc)
        self.const_dist = z(const_dist, nc)
        self.const_ext = z(const_ext, nc) if model == "rig" else np.ones(nc, bool)
        self.loss, self.loss_scale = None, 1.0     # None | "huber" | "cauchy" (per tag residual block)

    # ---- global parameter vector layout: [views | markers | per-cam (intr4 dist5 [ext6])]
    @property
    def shared_per_cam(self):
        return 15 if self.model == "rig" else 9

    def offsets(self):
        nv, nm = len(self.views), len(self.markers)
        o_view = 0
        o_marker = 6 * nv
        o_shared = o_marker + 6 * nm
        n = o_shared + self.shared_per_cam * self.n_cam
        return o_view, o_marker, o_shared, n

    def pack(self):
        parts = [self.views.ravel(), self.markers.ravel()]
        for c in range(self.n_cam):
            parts += [self.intr[c], self.dist[c]]
            if self.model == "rig":
                parts.append(self.ext[c])
        return np.concatenate(parts)

    def unpack(self, x):
        o_view, o_marker, o_shared, n = self.offsets()
        assert x.shape[0] == n
        self.views = x[o_view:o_marker].reshape(-1, 6).copy()
        self.markers = x[o_marker:o_shared].reshape(-1, 6).copy()
        sp = self.shared_per_cam
        sh = x[o_shared:].reshape(self.n_cam, sp)
        self.intr = sh[:, 0:4].copy()
        self.dist = sh[:, 4:9].copy()
        if self.model == "rig":
            self.ext = sh[:, 9:15].copy()

    def const_mask(self):
        """boolean mask over the global parameter vector: True = held constant."""
        o_view, o_marker, o_shared, n = self.offsets()
        m = np.zeros(n, bool)
        m[o_view:o_marker] = np.repeat(self.const_views, 6)
        m[o_marker:o_shared] = np.repeat(self.const_markers, 6)
        sp = self.shared_per_cam
        for c in range(self.n_cam):
            b = o_shared + sp * c
            m[b:b + 4] = self.const_intr[c]
            m[b + 4:b + 9] = self.const_dist[c]
            if self.model == "rig":
                m[b + 9:b + 15] = self.const_ext[c]
        return m

    def copy(self):
        q = self._copy()
        q.loss, q.loss_scale = self.loss, self.loss_scale
        return q

    def _copy(self):
        return OracleProblem(self.model, self.intr, self.dist, self.ext, self.views,
                             self.markers, self.sizes, self.view_idx, self.marker_idx,
                             self.cam_idx, self.pixels, self.const_views,
                             self.const_markers, self.const_intr, self.const_dist,
                             self.const_ext if self.model == "rig" else None)


# --------------------------------------------------------------------------
# Oracle A: closed-form model (works for real and complex dtypes)
# --------------------------------------------------------------------------
def rodrigues(r):
    """Rotation matrices for Rodrigues vectors r (...,3) -> (...,3,3).

    Same parameterisation as cv::Rodrigues (camera_pose.cpp:93,116,164).
    R = I + A [r]x + B [r]x^2,  A = sin(t)/t,  B = (1-cos t)/t^2 computed via
    the half angle so it stays accurate near t = 0 (world tag sits at r = 0,
    camera_pose.cpp:75-77).  Analytic in r, so complex-step safe.
    """
    r = np.asarray(r)
    t2 = (r * r).sum(-1)
    small = np.abs(t2) < 1e-8
    t2s = np.where(small, 1.0, t2)
    t = np.sqrt(t2s)
    A = np.where(small, 1 - t2 / 6 + t2 * t2 / 120, np.sin(t) / t)
    sh = np.sin(t / 2) / (t / 2)
    B = np.where(small, 0.5 - t2 / 24 + t2 * t2 / 720, 0.5 * sh * sh)
    x, y, z = r[..., 0], r[..., 1], r[..., 2]
    zero = np.zeros_like(x)
    K = np.stack([np.stack([zero, -z, y], -1),
                  np.stack([z, zero, -x], -1),
                  np.stack([-y, x, zero], -1)], -2)
    K2 = K @ K
    eye = np.eye(3, dtype=r.dtype)
    return eye + A[..., None, None] * K + B[..., None, None] * K2


def obj_points(size):
    """(...,) tag sizes -> (...,4,3) corner coordinates in the tag frame,
    order bl, br, tr, tl (camera_pose.cpp:123-126, 158-161)."""
    s = np.asarray(size) / 2
    z = np.zeros_like(s)
    return np.stack([np.stack([-s, -s, z], -1), np.stack([s, -s, z], -1),
                     np.stack([s, s, z], -1), np.stack([-s, s, z], -1)], -2)


def project_blocks(model, intr, dist, ext, view, marker, size, dtype=np.float64):
    """Project the 4 corners of N tags.  All inputs are per-block (N,k) arrays
    (already gathered).  Returns (uv (N,4,2), z (N,4))."""
    intr = np.asarray(intr, dtype=dtype)
    dist = np.asarray(dist, dtype=dtype)
    view = np.asarray(view, dtype=dtype)
    marker = np.asarray(marker, dtype=dtype)
    o = obj_points(size).astype(dtype)                               # (N,4,3)
    Rm = rodrigues(marker[:, 0:3])
    Pw = np.einsum('nij,nkj->nki', Rm, o) + marker[:, None, 3:6]      # world_T_target
    Rv = rodrigues(view[:, 0:3])
    q = np.einsum('nji,nkj->nki', Rv, Pw - view[:, None, 3:6])        # R^T (Pw - t)
    if model == "rig":
        ext = np.asarray(ext, dtype=dtype)
        Rx = rodrigues(ext[:, 0:3])
        q = np.einsum('nji,nkj->nki', Rx, q - ext[:, None, 3:6])
    X, Y, Z = q[..., 0], q[..., 1], q[..., 2]
    x = X / Z
    y = Y / Z
    r2 = x * x + y * y
    k1, k2, p1, p2, k3 = (dist[:, i:i + 1] for i in range(5))
    rad = 1 + r2 * (k1 + r2 * (k2 + r2 * k3))
    xd = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    u = intr[:, 0:1] * xd + intr[:, 2:3]
    v = intr[:, 1:2] * yd + intr[:, 3:4]
    return np.stack([u, v], -1), Z


def _gather(p):
    return (p.intr[p.cam_idx], p.dist[p.cam_idx], p.ext[p.cam_idx],
            p.views[p.view_idx], p.markers[p.marker_idx], p.sizes[p.marker_idx])


def residuals(p):
    """(N,8) residuals  res = projected - observed,  rows u0 v0 u1 v1 ..."""
    uv, _ = project_blocks(p.model, *_gather(p))
    return uv.reshape(-1, 8) - p.pixels


def depths(p):
    _, z = project_blocks(p.model, *_gather(p))
    return z


def robust(p, r):
    """Per-block (sqrt(rho'(s)), rho(s)) with s = ||r_block||^2 -- Ceres LossFunction
    semantics (HuberLoss / CauchyLoss with scale a); rho'' <= 0 for both, so the
    corrected residuals / Jacobian rows are simply scaled by sqrt(rho')."""
    s = (r * r).sum(1)
    a2 = p.loss_scale ** 2
    if p.loss == "huber":
        rt = np.sqrt(np.maximum(s, 1e-300))
        rho = np.where(s <= a2, s, 2 * p.loss_scale * rt - a2)
        rho1 = np.where(s <= a2, 1.0, p.loss_scale / rt)
    elif p.loss == "cauchy":
        rho, rho1 = a2 * np.log1p(s / a2), 1.0 / (1.0 + s / a2)
    else:
        rho, rho1 = s, np.ones_like(s)
    return np.sqrt(rho1), rho


def cost(p):
    r = residuals(p)
    return 0.5 * float(robust(p, r)[1].sum())


def local_param_names(model):
    """Order of the per-block Jacobian column groups (the Ceres parameter-block
    order the C-ABI uses): intr(4) dist(5) view(6) marker(6) [ext(6)]."""
    names = [("intr", 4), ("dist", 5), ("view", 6), ("marker", 6)]
    if model == "rig":
        names.append(("ext", 6))
    return names


def jacobian_blocks_cs(p):
    """Per-block Jacobians by complex step: dict name -> (N,8,k).

    Each block's local parameters are perturbed independently (blocks do not
    interact), vectorised over all N blocks."""
    intr, dist, ext, view, marker, size = _gather(p)
    base = dict(intr=intr, dist=dist, ext=ext, view=view, marker=marker)
    out = {}
    for name, k in local_param_names(p.model):
        J = np.empty((len(size), 8, k))
        for c in range(k):
            args = {n: a.astype(np.complex128) for n, a in base.items()}
            args[name][:, c] += 1j * CS_H
            uv, _ = project_blocks(p.model, args["intr"], args["dist"], args["ext"],
                                   args["view"], args["marker"], size,
                                   dtype=np.complex128)
            J[:, :, c] = uv.reshape(-1, 8).imag / CS_H
        out[name] = J
    return out


# --------------------------------------------------------------------------
# Oracle B: OpenCV's own analytic derivatives (small cases; python loop)
# --------------------------------------------------------------------------
def jacobian_block_cv2(model, intr, dist, ext, view, marker, size):
    """One block (8 residual values + Jacobians) from cv2.projectPoints /
    cv2.composeRT / cv2.Rodrigues.  Returns (uv(8,), dict name->(8,k))."""
    import cv2
    K = np.array([[intr[0], 0, intr[2]], [0, intr[1], intr[3]], [0, 0, 1.0]])
    o = obj_points(size)

    def invert(rt):
        """(r,t) of T -> (r',t') of T^-1 and d(r',t')/d(r,t) (6x6)."""
        r = rt[0:3].reshape(3, 1)
        t = rt[3:6].reshape(3, 1)
        R, dRdr = cv2.Rodrigues(r)                 # dRdr: 3x9, d vec(R)/dr (row-major R)
        ti = -R.T @ t
        D = np.zeros((6, 6))
        D[0:3, 0:3] = -np.eye(3)                   # r' = -r
        # t' = -R^T t : d t'_i / d r_k = - sum_j dR[j,i]/dr_k t_j
        dR = dRdr.reshape(3, 3, 3)                 # [k, a, b] = dR[a,b]/dr_k
        for k in range(3):
            D[3:6, k] = -(dR[k].T @ t).ravel()
        D[3:6, 3:6] = -R.T
        return np.concatenate([-r.ravel(), ti.ravel()]), D

    def compose(rt1, rt2):
        """T = T2 * T1 (apply T1 first), with 6x6 partials wrt rt1 and rt2."""
        res = cv2.composeRT(rt1[0:3].reshape(3, 1), rt1[3:6].reshape(3, 1),
                            rt2[0:3].reshape(3, 1), rt2[3:6].reshape(3, 1))
        r3, t3, dr3dr1, dr3dt1, dr3dr2, dr3dt2, dt3dr1, dt3dt1, dt3dr2, dt3dt2 = res
        D1 = np.block([[dr3dr1, dr3dt1], [dt3dr1, dt3dt1]])
        D2 = np.block([[dr3dr2, dr3dt2], [dt3dr2, dt3dt2]])
        return np.concatenate([r3.ravel(), t3.ravel()]), D1, D2

    # cam_T_target = inv(world_T_camera) * world_T_target        (single)
    #              = inv(body_T_cam) * inv(world_T_body) * world_T_target   (rig)
    vinv, Dv = invert(np.asarray(view, float))
    ct, D_m, D_vi = compose(np.asarray(marker, float), vinv)      # T = vinv * marker
    dct = {"marker": D_m, "view": D_vi @ Dv}
    if model == "rig":
        xinv, Dx = invert(np.asarray(ext, float))
        ct2, D_prev, D_xi = compose(ct, xinv)
        dct = {k: D_prev @ v for k, v in dct.items()}
        dct["ext"] = D_xi @ Dx
        ct = ct2
    pts, J = cv2.projectPoints(o, ct[0:3].reshape(3, 1), ct[3:6].reshape(3, 1), K,
                               np.asarray(dist, float))
    # J columns: rvec 0:3 | tvec 3:6 | fx fy 6:8 | cx cy 8:10 | k1 k2 p1 p2 k3 10:15
    Jrt = J[:, 0:6]
    out = {k: Jrt @ v for k, v in dct.items()}
    out["intr"] = J[:, 6:10]
    out["dist"] = J[:, 10:15]
    return pts.reshape(8), out


# --------------------------------------------------------------------------
# dense normal equations, Schur complement, LM
# --------------------------------------------------------------------------
def dense_jacobian(p, Jb=None):
    """Dense (8N x n) Jacobian in the global layout of OracleProblem.offsets()."""
    if Jb is None:
        Jb = jacobian_blocks_cs(p)
    o_view, o_marker, o_shared, n = p.offsets()
    N = len(p.view_idx)
    J = np.zeros((8 * N, n))
    sp = p.shared_per_cam
    rows = np.arange(8 * N).reshape(N, 8)
    for b in range(N):
        rr = rows[b]
        v, m, c = p.view_idx[b], p.marker_idx[b], p.cam_idx[b]
        J[np.ix_(rr, range(o_view + 6 * v, o_view + 6 * v + 6))] = Jb["view"][b]
        J[np.ix_(rr, range(o_marker + 6 * m, o_marker + 6 * m + 6))] = Jb["marker"][b]
        s0 = o_shared + sp * c
        J[np.ix_(rr, range(s0, s0 + 4))] = Jb["intr"][b]
        J[np.ix_(rr, range(s0 + 4, s0 + 9))] = Jb["dist"][b]
        if p.model == "rig":
            J[np.ix_(rr, range(s0 + 9, s0 + 15))] = Jb["ext"][b]
    return J


def normal_equations(p):
    """H = J^T J, g = J^T r, cost -- dense, constant blocks NOT yet masked."""
    J = dense_jacobian(p)
    r = residuals(p)
    w, rho = robust(p, r)
    J = J * np.repeat(w, 8)[:, None]
    r = (r * w[:, None]).ravel()
    return J.T @ J, J.T @ r, 0.5 * float(rho.sum())


def lm_diagonal(H, radius, min_diag=1e-6, max_diag=1e32):
    """Ceres-style LM damping  D^2 = clamp(diag(J^T J)) / radius."""
    return np.clip(np.diag(H), min_diag, max_diag) / radius


def masked_system(H, g, const_mask):
    """Hold constant parameters: zero their rows/cols, unit diagonal, zero rhs."""
    H = H.copy()
    g = g.copy()
    idx = np.nonzero(const_mask)[0]
    H[idx, :] = 0
    H[:, idx] = 0
    H[idx, idx] = 1
    g[idx] = 0
    return H, g


def schur_reduce(H, g, e_slice, f_index):
    """Eliminate the parameters in e_slice (block diagonal 6x6 there) and return
    (S, b) over f_index:  S = Hff - Hfe Hee^-1 Hef,  b = gf - Hfe Hee^-1 ge."""
    e_index = np.arange(e_slice.start, e_slice.stop)
    Hee = H[np.ix_(e_index, e_index)]
    Hef = H[np.ix_(e_index, f_index)]
    Hff = H[np.ix_(f_index, f_index)]
    X = np.linalg.solve(Hee, np.concatenate([Hef, g[e_index, None]], 1))
    S = Hff - Hef.T @ X[:, :-1]
    b = g[f_index] - Hef.T @ X[:, -1]
    return S, b


def lm_step(p, radius):
    """One damped Gauss-Newton step.  Returns (delta, model_cost_change, cost, gmax)."""
    H, g, c = normal_equations(p)
    cm = p.const_mask()
    d2 = lm_diagonal(H, radius)
    Hd = H + np.diag(d2)
    Hd, gm = masked_system(Hd, g, cm)
    delta = -np.linalg.solve(Hd, gm)
    delta[cm] = 0
    d2m = np.where(cm, 0.0, d2)
    mcc = -0.5 * float(gm @ delta) + 0.5 * float(delta @ (d2m * delta))
    return delta, mcc, c, float(np.abs(gm).max())


def lm_solve(p, max_iters=50, initial_radius=1e4, min_relative_decrease=1e-3,
             function_tolerance=1e-15, gradient_tolerance=1e-12,
             parameter_tolerance=1e-14, verbose=False):
    """Ceres-style Levenberg-Marquardt trust region (semantics recalled from
    Ceres defaults; unverifiable here -- convergence parity only
    needs both sides to reach the same unique minimum).  Modifies p in place."""
    radius = initial_radius
    decrease = 2.0
    c = cost(p)
    hist = [c]
    for it in range(max_iters):
        delta, mcc, c, gmax = lm_step(p, radius)
        if gmax < gradient_tolerance:
            break
        x = p.pack()
        if np.linalg.norm(delta) <= parameter_tolerance * (np.linalg.norm(x) + parameter_tolerance):
            break
        cand = p.copy()
        cand.unpack(x + delta)
        z = depths(cand)
        c_new = cost(cand) if np.all(z > 0) else np.inf
        rho = (c - c_new) / mcc if mcc > 0 else -1.0
        if verbose:
            print(f"it {it:3d} cost {c:.6e} -> {c_new:.6e} rho {rho:.3f} radius {radius:.3e}")
        if np.isfinite(c_new) and rho > min_relative_decrease:
            p.unpack(x + delta)
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3))
            decrease = 2.0
            hist.append(c_new)
            if abs(c - c_new) <= function_tolerance * c:
                break
        else:
            radius /= decrease
            decrease *= 2.0
    return hist
