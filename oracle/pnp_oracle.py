"""cv2-based restatement of the reference's initialiser (camera_pose.cpp:132-285).
TEST INFRASTRUCTURE ONLY.  Per tag: cv2.solvePnP(..., SOLVEPNP_ITERATIVE), the call
at camera_pose.cpp:163 (OpenCV 4.13 here, 3.4.4 in the reference); then the same
pose chaining with cv2.Rodrigues and 4x4 matrices as the reference does with Eigen.
Parity: pinned to OpenCV's solver (the reference has no fixtures of its own)."""
import cv2
import numpy as np


POLISH = False   # True: run OpenCV's own LM (solvePnPRefineLM) to a tight TermCriteria after the stock call


def tag_T_cam_inverse(tag_size, pixels8, K, dist):
    """cam_T_tag as 4x4 from cv2.solvePnP (camera_pose.cpp:146-170).  The stock call stops at OpenCV's default
    tolerance (usually 1e-9 from the minimiser, occasionally 0.2 away along a flat depth direction); with
    POLISH the same cost is minimised to convergence so that a tight comparison is meaningful."""
    h = tag_size / 2
    obj = np.array([[-h, -h, 0], [h, -h, 0], [h, h, 0], [-h, h, 0]], float)
    img = np.asarray(pixels8, float).reshape(4, 2)
    ok, rvec, tvec = cv2.solvePnP(obj, img, K, dist, flags=cv2.SOLVEPNP_ITERATIVE)
    if POLISH:
        rvec, tvec = cv2.solvePnPRefineLM(obj, img, K, dist, rvec, tvec,
                                          criteria=(cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_COUNT, 1000, 1e-16))
    R, _ = cv2.Rodrigues(rvec)
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, tvec.ravel()
    return T


def to_rt(T):
    r, _ = cv2.Rodrigues(T[:3, :3])
    return np.concatenate([r.ravel(), T[:3, 3]])


def initialise(frames_pixels, intr, dist, polish=False):
    global POLISH
    POLISH = bool(polish)
    K = np.array([[intr[0], 0, intr[2]], [0, intr[1], intr[3]], [0, 0, 1.0]])
    dist = np.asarray(dist, float)
    ids, sizes, trans = [], [], []
    w_T_cam = [None] * len(frames_pixels)
    world = frames_pixels[0][0][0]
    ids.append(world); sizes.append(frames_pixels[0][0][1]); trans.append(np.eye(4))

    def status(n):
        if n == 0:
            return 0, 0
        st, known = 2, -1
        for k, (tid, _, _) in enumerate(frames_pixels[n]):
            if tid == world:
                return 0, k
            if tid in ids:
                known, st = k, 1
        return st, known

    def tag_calc(n, known):
        tid, sz, px = frames_pixels[n][known]
        wTc = trans[ids.index(tid)] @ np.linalg.inv(tag_T_cam_inverse(sz, px, K, dist))
        w_T_cam[n] = wTc
        for k, (t, s, p) in enumerate(frames_pixels[n]):
            if k != known and t not in ids:
                trans.append(wTc @ tag_T_cam_inverse(s, p, K, dist))
                ids.append(t); sizes.append(s)

    deferred = []
    for n in range(len(frames_pixels)):
        if not frames_pixels[n]:
            continue
        st, known = status(n)
        if st in (0, 1):
            tag_calc(n, known)
            for j in range(len(deferred) - 1, -1, -1):
                st2, k2 = status(deferred[j])
                if st2 == 1:
                    tag_calc(deferred[j], k2)
                    deferred.pop(j)
        else:
            deferred.append(n)
    return ids, sizes, np.array([to_rt(T) for T in trans]), [None if T is None else to_rt(T) for T in w_T_cam]
