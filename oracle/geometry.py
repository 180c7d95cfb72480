"""fp64 numpy restatement of the geometry on the hot path (test infrastructure only).

Every function cites the reference lines (relative to /root/reference) it restates.
Third-party arithmetic the reference calls and that is NOT under /root/reference
(opencv-python, unpinned in requirements.txt:28; 4.13.0 in this image) is restated from
its published algorithm and pinned against cv2 itself in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np

# OpenCV distortion coefficient order (k1,k2,p1,p2,k3,k4,k5,k6,s1,s2,s3,s4,taux,tauy)
N_DIST = 14

# camera_calibration/calibration_parameters.npz["camera_matrix"] == triangulation/main.py:51-63
K_CALIB = np.array(
    [
        [1116.9289548941917, 0.0, 955.77175993563799],
        [0.0, 1117.3341496962166, 538.91061167202145],
        [0.0, 0.0, 1.0],
    ]
)
# camera_calibration/calibration_parameters.npz["dist_coeffs"] == triangulation/triangulate.py:39-56
DIST_CALIB = np.array(
    [
        -1.1940477842823853,
        -15.440461757486913,
        0.00013163161053023783,
        0.00019082529328353381,
        98.843073622415901,
        -1.3588290520381034,
        -14.555841222727574,
        96.219667412855202,
        0.0,
        0.0,
        0.0,
        0.0,
        0.0,
        0.0,
    ]
)


# --------------------------------------------------------------------------- rotations
def hat(v: np.ndarray) -> np.ndarray:
    """[v]x for (...,3) -> (...,3,3)."""
    v = np.asarray(v, float)
    z = np.zeros_like(v[..., 0])
    return np.stack(
        [
            np.stack([z, -v[..., 2], v[..., 1]], -1),
            np.stack([v[..., 2], z, -v[..., 0]], -1),
            np.stack([-v[..., 1], v[..., 0], z], -1),
        ],
        -2,
    )


def rodrigues(rvec: np.ndarray) -> np.ndarray:
    """Rotation vector -> matrix; what cv2.Rodrigues computes for a 3-vector
    (used at triangulation/reproject.py:69, bundle_adjustment/reproject.py:139)."""
    r = np.asarray(rvec, float).reshape(3)
    th = np.linalg.norm(r)
    if th < 1e-12:
        return np.eye(3) + hat(r)
    k = r / th
    Kx = hat(k)
    return np.eye(3) + np.sin(th) * Kx + (1.0 - np.cos(th)) * (Kx @ Kx)


def rodrigues_inv(R: np.ndarray) -> np.ndarray:
    """Rotation matrix -> rotation vector (cv2.Rodrigues on a 3x3), theta=pi branch included
    (the reference's FIXED rig, two_view.py:209-221, sits exactly on it)."""
    R = np.asarray(R, float).reshape(3, 3)
    # project onto SO(3) like OpenCV does (SVD) so slightly non-orthogonal input behaves the same
    U, _, Vt = np.linalg.svd(R)
    R = U @ Vt
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = 0.5 * np.linalg.norm(v)
    c = np.clip(0.5 * (np.trace(R) - 1.0), -1.0, 1.0)
    th = np.arccos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        t = np.sqrt(np.maximum((np.diag(R) + 1.0) * 0.5, 0.0))
        if R[0, 1] < 0:
            t[1] = -t[1]
        if R[0, 2] < 0:
            t[2] = -t[2]
        if abs(t[0]) < abs(t[1]) and abs(t[0]) < abs(t[2]) and (R[1, 2] > 0) != (t[1] * t[2] > 0):
            t[2] = -t[2]
        return t * (th / np.linalg.norm(t))
    return v * (0.5 * th / s)


def so3_exp(w: np.ndarray) -> np.ndarray:
    """exp([w]x), series below 1e-4 rad."""
    w = np.asarray(w, float).reshape(3)
    th2 = float(w @ w)
    W = hat(w)
    if th2 < 1e-8:
        a, b = 1.0 - th2 / 6.0, 0.5 - th2 / 24.0
    else:
        th = np.sqrt(th2)
        a, b = np.sin(th) / th, (1.0 - np.cos(th)) / th2
    return np.eye(3) + a * W + b * (W @ W)


# --------------------------------------------------------------------------- triangulation
def make_P(K, R, t) -> np.ndarray:
    """P = K [R | t]  (vggt/triangulate.py:13-16; triangulation/triangulate.py:65-66)."""
    K = np.asarray(K, float)
    Rt = np.concatenate([np.asarray(R, float).reshape(3, 3), np.asarray(t, float).reshape(3, 1)], 1)
    return K @ Rt


def dlt_rows(P: np.ndarray, x: np.ndarray, w: np.ndarray | None = None, weight_power: float = 1.0):
    """Stack the 2V x 4 DLT system for N points.

    rows w_v*(u_v*P_v[2]-P_v[0]), w_v*(v_v*P_v[2]-P_v[1]) in un-normalised pixel units, exactly
    the rows of vggt/triangulate.py:23-31 (== cv2.triangulatePoints for V=2, w=1).
    P (V,3,4) or (N,V,3,4); x (V,N,2); w (V,N) or None -> A (N,2V,4)
    """
    P = np.asarray(P, float)
    x = np.asarray(x, float)
    V, N = x.shape[0], x.shape[1]
    if P.ndim == 3:
        P = np.broadcast_to(P[None], (N, V, 3, 4))
    A = np.empty((N, 2 * V, 4))
    for v in range(V):
        A[:, 2 * v] = x[v, :, 0:1] * P[:, v, 2] - P[:, v, 0]
        A[:, 2 * v + 1] = x[v, :, 1:2] * P[:, v, 2] - P[:, v, 1]
        if w is not None:
            wv = np.asarray(w[v], float) ** weight_power
            A[:, 2 * v] *= wv[:, None]
            A[:, 2 * v + 1] *= wv[:, None]
    return A


def dlt_triangulate(P, x, w=None, weight_power: float = 1.0) -> np.ndarray:
    """V-view (confidence-weighted) DLT: last right-singular vector of A, dehomogenised
    (vggt/triangulate.py:32-34; triangulation/triangulate.py:67-68).  No guards: w->0 or NaN
    propagate as inf/NaN like the reference.  Returns (N,3) fp64."""
    A = dlt_rows(P, x, w, weight_power)
    out = np.full((A.shape[0], 3), np.nan)
    ok = np.isfinite(A).all(axis=(1, 2))
    if ok.any():
        _, _, Vt = np.linalg.svd(A[ok])
        h = Vt[:, -1]
        with np.errstate(divide="ignore", invalid="ignore"):
            out[ok] = h[:, :3] / h[:, 3:4]
    return out


# --------------------------------------------------------------------------- projection
def distort_normalised(x, y, dist):
    """OpenCV rational + tangential + thin-prism model on normalised coords (what
    cv2.projectPoints applies).  Tilt (taux,tauy) must be zero."""
    d = np.zeros(N_DIST)
    if dist is not None:
        dd = np.asarray(dist, float).reshape(-1)
        d[: dd.size] = dd
    if d[12] != 0.0 or d[13] != 0.0:
        raise NotImplementedError("tilted sensor model (taux,tauy) not supported")
    k1, k2, p1, p2, k3, k4, k5, k6, s1, s2, s3, s4 = d[:12]
    r2 = x * x + y * y
    r4 = r2 * r2
    r6 = r4 * r2
    rad = (1.0 + k1 * r2 + k2 * r4 + k3 * r6) / (1.0 + k4 * r2 + k5 * r4 + k6 * r6)
    xd = x * rad + 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x) + s1 * r2 + s2 * r4
    yd = y * rad + p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y + s3 * r2 + s4 * r4
    return xd, yd


def project_cv(X, R, t, K, dist=None) -> np.ndarray:
    """cv2.projectPoints semantics in fp64: X_c = R X + t, perspective divide, distortion,
    u = fx*x'' + cx, v = fy*y'' + cy (skew ignored, like OpenCV).  X (...,3) -> (...,2).
    Restates the calls at triangulation/reproject.py:77-78 / bundle_adjustment/reproject.py:147-148."""
    X = np.asarray(X, float)
    R = np.asarray(R, float).reshape(3, 3)
    t = np.asarray(t, float).reshape(3)
    K = np.asarray(K, float).reshape(3, 3)
    Xc = X @ R.T + t
    with np.errstate(divide="ignore", invalid="ignore"):
        x = Xc[..., 0] / Xc[..., 2]
        y = Xc[..., 1] / Xc[..., 2]
        xd, yd = distort_normalised(x, y, dist)
    return np.stack([K[0, 0] * xd + K[0, 2], K[1, 1] * yd + K[1, 2]], -1)


def project_loss(X3d, R, t, K, zmin: float = 1e-6) -> np.ndarray:
    """bundle_adjustment/loss.py:17-84 in numpy fp64: X_cam = R X + t (:60-64),
    Z = clamp(z, min=1e-6) (:67), xy = X_cam.xy / Z (:68), proj = (K [x,y,1])[:2] with the FULL
    3x3 K, skew honoured (:74-82).  X3d (T,J,3)|(J,3); R (C,3,3)|(T,C,3,3); t (C,3)|(T,C,3);
    K (C,3,3)|(T,C,3,3) -> (T,C,J,2)."""
    X3d = np.asarray(X3d, float)
    R = np.asarray(R, float)
    t = np.asarray(t, float)
    K = np.asarray(K, float)
    if X3d.ndim == 2:
        X3d = X3d[None]
    T = X3d.shape[0]
    if R.ndim == 3:
        R = np.broadcast_to(R[None], (T,) + R.shape)
        t = np.broadcast_to(t[None], (T,) + t.shape)
    elif R.ndim == 4:
        assert R.shape[0] == T
        if t.ndim == 2:
            t = np.broadcast_to(t[None], (T,) + t.shape)
        else:
            assert t.shape[:2] == R.shape[:2]
    else:
        raise ValueError(f"Unsupported R shape: {R.shape}")
    if K.ndim == 3:
        K = K[None]
    elif K.ndim != 4:
        raise ValueError(f"Unsupported K shape: {K.shape}")
    Xc = np.einsum("tcab,tjb->tcja", R, X3d) + t[:, :, None, :]
    Z = np.maximum(Xc[..., 2], zmin)
    x = Xc[..., 0] / Z
    y = Xc[..., 1] / Z
    K = K[:, :, None]  # (T|1,C,1,3,3)
    u = K[..., 0, 0] * x + K[..., 0, 1] * y + K[..., 0, 2]
    v = K[..., 1, 0] * x + K[..., 1, 1] * y + K[..., 1, 2]
    return np.stack([u, v], -1)


def reprojection_loss(X3d, R, t, K, x2d, conf2d, w: float = 1.0) -> float:
    """bundle_adjustment/loss.py:90-94: w * sum(conf * |proj - x2d|^2) / (sum(conf) + 1e-6)."""
    pred = project_loss(X3d, R, t, K)
    diff = ((pred - np.asarray(x2d, float)) ** 2).sum(-1)
    conf2d = np.asarray(conf2d, float)
    return float(w * (conf2d * diff).sum() / (conf2d.sum() + 1e-6))


# --------------------------------------------------------------------------- regularisers (loss.py:97-155)
BONES = [(11, 13), (13, 15), (12, 14), (14, 16), (5, 7), (7, 9), (6, 8), (8, 10), (5, 6), (11, 12), (5, 11), (6, 12)]


def camera_center_from_Rt(R, t):
    """C = -R^T t  (loss.py:97-100)."""
    return -np.einsum("...ba,...b->...a", np.asarray(R, float), np.asarray(t, float))


def camera_smooth_loss(R, t, w=1e-2):
    """loss.py:103-106."""
    C = camera_center_from_Rt(R, t)
    return float(w * ((C[1:] - C[:-1]) ** 2).mean())


def baseline_reg_loss(R, t, w=1e-2):
    """loss.py:109-114 (first two cameras only; 0 for a single camera)."""
    C = camera_center_from_Rt(R, t)
    if C.shape[1] < 2:
        return 0.0
    b = np.linalg.norm(C[:, 0] - C[:, 1], axis=-1)
    return float(w * ((b - b.mean()) ** 2).mean())


def bone_length_loss(X3d, ref_bone_len=None, w=1e-2):
    """loss.py:134-150 (bones whose indices exceed J are skipped)."""
    X3d = np.asarray(X3d, float)
    J = X3d.shape[1]
    lens = [np.linalg.norm(X3d[:, i] - X3d[:, j], axis=-1) for i, j in BONES if i < J and j < J]
    if not lens:
        return 0.0
    L = np.stack(lens, -1)
    ref = L.mean(0, keepdims=True) if ref_bone_len is None else np.asarray(ref_bone_len, float)[None, :]
    return float(w * ((L - ref) ** 2).mean())


def pose_temporal_loss(X3d, w=1e-2):
    """loss.py:153-155."""
    X3d = np.asarray(X3d, float)
    return float(w * ((X3d[1:] - X3d[:-1]) ** 2).mean())


# --------------------------------------------------------------------------- statistics
def error_stats(err: np.ndarray) -> dict:
    """nan-aware rmse/mean/median/max of per-joint pixel errors
    (triangulation/reproject.py:249-261; bundle_adjustment/reproject.py:333-345)."""
    err = np.asarray(err, float)
    with np.errstate(invalid="ignore"):
        return {
            "rmse": float(np.sqrt(np.nanmean(err**2))),
            "mean": float(np.nanmean(err)),
            "median": float(np.nanmedian(err)),
            "max": float(np.nanmax(err)),
        }
