"""The reference's per-frame CPU path restated call-for-call (test infrastructure / CPU baseline).

This is the "port" leg of bench.py's cpu_baseline and --impl reference arm: the same third-party
calls the reference makes (cv2.triangulatePoints, cv2.Rodrigues, cv2.projectPoints, np.linalg.svd),
with the same down-casts (quirk Q2), in the same one-frame-at-a-time loop, but written here from
the call-site description in SURVEY.md section 8a - /root/reference does not travel to the GPU box.
It is pinned against the reference's own outputs by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np

from . import geometry as G

try:  # cv2 ships in the image; keep the numpy fallbacks honest if it ever does not
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def triangulate_two_view(kp1, kp2, K, R, T) -> np.ndarray:
    """triangulation/triangulate.py:60-68: P1=K[I|0], P2=K[R|T], cv2.triangulatePoints,
    dehomogenise.  Output dtype follows the keypoint dtype (f32 in -> f32 out)."""
    kp1 = np.asarray(kp1)
    kp2 = np.asarray(kp2)
    if kp1.shape != kp2.shape or kp1.shape[1] != 2:
        raise ValueError(f"Keypoints shape mismatch: {kp1.shape} vs {kp2.shape}")
    P1 = G.make_P(K, np.eye(3), np.zeros(3))
    P2 = G.make_P(K, R, np.asarray(T, float).reshape(3))
    if cv2 is not None:
        h = cv2.triangulatePoints(P1, P2, kp1.T, kp2.T)
        return (h[:3] / h[3]).T
    X = G.dlt_triangulate(np.stack([P1, P2]), np.stack([kp1, kp2]))
    return X.astype(kp1.dtype if kp1.dtype.kind == "f" else np.float64)


def triangulate_point_svd(P1, P2, x1, x2) -> np.ndarray:
    """vggt/triangulate.py:19-34: 4x4 system, np.linalg.svd, Vt[-1], dehomogenise -> (3,) f64."""
    A = G.dlt_rows(np.stack([P1, P2]), np.asarray([x1, x2], float).reshape(2, 1, 2))[0]
    h = np.linalg.svd(A)[2][-1]
    return (h / h[3])[:3]


def triangulate_frame_svd(K, R, T, kptL, kptR) -> np.ndarray:
    """vggt/triangulate.py:64-71: per-joint loop into a float32 (J,3) buffer (quirk Q4)."""
    P1 = G.make_P(K[0], R[0], T[0])
    P2 = G.make_P(K[1], R[1], T[1])
    out = np.zeros((len(kptL), 3), np.float32)
    for j in range(len(kptL)):
        out[j] = triangulate_point_svd(P1, P2, kptL[j], kptR[j])
    return out


def relative_pose(R, T):
    """Extrinsics parsing of bundle_adjustment/reproject.py:99-132 (== vggt/reproject.py):
    mode A (2,3,3)/(2,3) world->cam pairs -> cam1->cam2 relative pose in float32;
    mode B (3,3)/(3,) or (3,1) relative pose as given.  ValueError like the reference."""
    R = np.asarray(R, np.float32)
    T = np.asarray(T, np.float32)
    if R.ndim == 3:
        if R.shape[0] != 2 or T.shape[0] != 2:
            raise ValueError(f"Expect R,T shape (2,3,3),(2,3), got {R.shape}, {T.shape}")
        R_rel = R[1] @ R[0].T
        t_rel = T[1].reshape(3, 1) - R_rel @ T[0].reshape(3, 1)
    elif R.ndim == 2:
        if R.shape != (3, 3) or T.shape not in [(3,), (3, 1)]:
            raise ValueError(f"Expect R(3,3), T(3,) for relative extrinsic, got {R.shape}, {T.shape}")
        R_rel, t_rel = R, T.reshape(3, 1)
    else:
        raise ValueError(f"Unsupported R shape: {R.shape}")
    return R_rel.astype(np.float32), t_rel.astype(np.float32)


def reproject_two_view(X3, K1, dist1, K2, dist2, R, T) -> dict:
    """triangulation/reproject.py:49-83 and bundle_adjustment/reproject.py:74-153: everything cast
    to float32, cam1 rvec=tvec=0, cam2 = Rodrigues round trip of the relative pose, projectPoints."""
    f32 = np.float32
    X3_ = np.asarray(X3, f32).reshape(-1, 1, 3)
    R_rel, t_rel = relative_pose(R, T)
    K1 = np.asarray(K1, f32).reshape(3, 3)
    K2 = np.asarray(K2, f32).reshape(3, 3)
    d1 = None if dist1 is None else np.asarray(dist1, f32).reshape(-1, 1)
    d2 = None if dist2 is None else np.asarray(dist2, f32).reshape(-1, 1)
    if cv2 is not None:
        z = np.zeros((3, 1), f32)
        rvec2, _ = cv2.Rodrigues(R_rel)
        p1, _ = cv2.projectPoints(X3_, z, z, K1, d1)
        p2, _ = cv2.projectPoints(X3_, rvec2, t_rel, K2, d2)
        return {"proj_L": p1.reshape(-1, 2), "proj_R": p2.reshape(-1, 2)}
    R2 = G.rodrigues(G.rodrigues_inv(R_rel).astype(f32))
    p1 = G.project_cv(X3_.reshape(-1, 3), np.eye(3), np.zeros(3), K1, d1)
    p2 = G.project_cv(X3_.reshape(-1, 3), R2, t_rel, K2, d2)
    return {"proj_L": p1.astype(f32), "proj_R": p2.astype(f32)}


def reprojection_report(proj: dict, kptL, kptR) -> dict:
    """Error statistics of reproject_and_visualize (triangulation/reproject.py:243-261): per-joint
    L2 pixel error in float64 (kpt cast with float, proj stays f32 - quirk Q6), nan-aware stats."""
    errL = np.linalg.norm(proj["proj_L"] - np.asarray(kptL, float), axis=1)
    errR = np.linalg.norm(proj["proj_R"] - np.asarray(kptR, float), axis=1)
    out = {"proj_L": proj["proj_L"], "proj_R": proj["proj_R"], "err_L": errL, "err_R": errR}
    for side, e in (("L", errL), ("R", errR)):
        s = G.error_stats(e)
        out[f"rmse_{side}"] = s["rmse"]
        out[f"mean_err_{side}"] = s["mean"]
        out[f"median_err_{side}"] = s["median"]
        out[f"max_err_{side}"] = s["max"]
    return out


def clip_two_view(left_kpts, right_kpts, K, R_list, T_list, dist=None, frames=None):
    """The hot loop of process_triangulate (triangulation/triangulate.py:76-116) without the
    image drawing: per frame triangulate -> reproject (dist1=dist2=dist) -> stats.
    Returns X (T,J,3), err (2,T,J) f64 and the per-frame stats list."""
    T = len(left_kpts) if frames is None else frames
    Xs, errs, stats = [], [], []
    for i in range(T):
        X = triangulate_two_view(left_kpts[i], right_kpts[i], K, R_list[i], T_list[i])
        rep = reprojection_report(
            reproject_two_view(X, K, dist, K, dist, R_list[i], T_list[i]), left_kpts[i], right_kpts[i]
        )
        Xs.append(X)
        errs.append(np.stack([rep["err_L"], rep["err_R"]]))
        stats.append({k: v for k, v in rep.items() if isinstance(v, float)})
    return np.stack(Xs), np.stack(errs, 1), stats
