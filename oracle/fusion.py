"""fp64 numpy restatement of the reference's two-view 3D-3D fusion + adaptive EMA smoothing.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg may import this.

Array form (NaN rows = missing joints) of the reference's dict-per-frame code, function by function:
  kabsch_rigid_align        fuse/main_raw.py:48-68     (np.linalg.svd of the 3x3 cross-covariance, det fix)
  align_right_to_left       fuse/main_raw.py:71-95     (rigid right -> left over the joints finite in both views)
  fit_weakpersp_3d_to_2d    fuse/confidence.py:9-59    (u ~ s X M + t, M (3,2) orthonormal columns from the 3x2 SVD)
  weakpersp_reproj_conf     fuse/confidence.py:62-108  (conf = exp(-err^2 / (2 sigma_px^2)), 0 where undefined)
  canonicalize_pose_3d      fuse/confidence.py:118-181 (pelvis origin, hip / shoulder axes, hip-width or torso scale)
  crossview_consistency     fuse/confidence.py:183-224 (conf = exp(-dist^2 / (2 sigma_3d^2)) in canonical space)
  softmax2, fuse_frame_3d   fuse/fuse.py:87-94, 289-326
  temporal_smooth_ema       fuse/fuse.py:329-412       (per-joint base alpha, speed-adaptive alpha, NaN hold / reset)
  fuse_clip                 fuse/main_raw.py:199-250   (the per-frame pipeline of the `fuse` main)
PARITY: pinned by golden G8 (tests/golden/g8_fusion.npz), produced by oracle/make_golden.py from the reference's own
functions imported from /root/reference.
"""
from __future__ import annotations

import numpy as np

EPS = 1e-8  # fuse/fuse.py:19

# fuse/main_raw.py:18-22
IDX_PELVIS, IDX_LHIP, IDX_RHIP, IDX_LSHO, IDX_RSHO = 14, 11, 12, 5, 6
# fuse/fuse.py:364-366 (joint ids of the SAM-3D-Body skeleton)
CORE_IDS = (1, 2, 69)
LIMB_IDS = (5, 6, 7, 8, 9, 10, 11, 12)
ENDPOINT_IDS = (13, 14, 41, 62)


def _finite_rows(X):
    return np.isfinite(X).all(axis=1)


def kabsch_rigid_align(src, dst):
    """fuse/main_raw.py:48-68.  R, t with R src + t ~ dst."""
    sm, dm = src.mean(axis=0), dst.mean(axis=0)
    H = (src - sm).T @ (dst - dm)
    U, _, Vt = np.linalg.svd(H)
    R = Vt.T @ U.T
    if np.linalg.det(R) < 0:
        Vt[-1, :] *= -1
        R = Vt.T @ U.T
    return R, dm - R @ sm


def align_right_to_left(Xl, Xr):
    """fuse/main_raw.py:71-95 on (J,3) arrays: fewer than 3 common joints -> Xr unchanged."""
    valid = _finite_rows(Xl) & _finite_rows(Xr)
    if int(valid.sum()) < 3:
        return Xr.copy()
    R, t = kabsch_rigid_align(Xr[valid], Xl[valid])
    out = Xr.copy()
    out[valid] = (R @ Xr[valid].T).T + t
    return out


def fit_weakpersp_3d_to_2d(X3d, U2d, min_points=8):
    """fuse/confidence.py:9-59."""
    valid = _finite_rows(X3d) & _finite_rows(U2d)
    idx = np.where(valid)[0]
    if idx.size < min_points:
        raise ValueError(f"Not enough valid points to fit: {idx.size} < {min_points}")
    X, U = X3d[idx], U2d[idx]
    muX, muU = X.mean(axis=0, keepdims=True), U.mean(axis=0, keepdims=True)
    Xc, Uc = X - muX, U - muU
    C = Xc.T @ Uc
    Us, S, Vt = np.linalg.svd(C, full_matrices=True)
    M = Us[:, :2] @ Vt
    denom = (Xc**2).sum()
    if denom < 1e-12:
        raise ValueError("Degenerate 3D points (too small variance).")
    s = S.sum() / denom
    t = (muU - s * (muX @ M)).reshape(2)
    return float(s), M, t, valid


def weakpersp_reproj_confidence(X3d, U2d, sigma_px=12.0, min_points=8, eps=1e-12):
    """fuse/confidence.py:62-108 -> conf (J,), err (J,)."""
    s, M, t, _ = fit_weakpersp_3d_to_2d(X3d, U2d, min_points)
    Uhat = s * (X3d @ M) + t
    err = np.full(X3d.shape[0], np.nan)
    ok = _finite_rows(U2d) & _finite_rows(Uhat)
    err[ok] = np.sqrt(((Uhat[ok] - U2d[ok]) ** 2).sum(axis=1))
    sig2 = max(float(sigma_px), eps) ** 2
    conf = np.zeros_like(err)
    vv = np.isfinite(err)
    conf[vv] = np.exp(-(err[vv] ** 2) / (2.0 * sig2))
    return conf, err


def _normalize(v, eps=1e-9):
    n = np.linalg.norm(v)
    return v * 0.0 if n < eps else v / n


def canonicalize_pose_3d(X, root=IDX_PELVIS, lhip=IDX_LHIP, rhip=IDX_RHIP, lsho=IDX_LSHO, rsho=IDX_RSHO, scale_mode="hip", eps=1e-9):
    """fuse/confidence.py:118-181 -> canonical (J,3) or all-NaN."""
    if not np.isfinite(X[[root, lhip, rhip, lsho, rsho]]).all():
        return np.full_like(X, np.nan)
    X0 = X - X[root]
    Lh, Rh, Ls, Rs = X0[lhip], X0[rhip], X0[lsho], X0[rsho]
    mid_hip, mid_sh = 0.5 * (Lh + Rh), 0.5 * (Ls + Rs)
    x_axis = _normalize(Rh - Lh, eps)
    y_axis = _normalize(mid_sh - mid_hip, eps)
    z_axis = _normalize(np.cross(x_axis, y_axis), eps)
    y_axis = _normalize(np.cross(z_axis, x_axis), eps)
    R = np.stack([x_axis, y_axis, z_axis], axis=0)
    Xr = (R @ X0.T).T
    if scale_mode == "hip":
        s = np.linalg.norm(Rh - Lh)
    elif scale_mode == "torso":
        s = np.linalg.norm(mid_sh - mid_hip)
    else:
        raise ValueError("scale_mode must be 'hip' or 'torso'")
    if not np.isfinite(s) or s < eps:
        return np.full_like(X, np.nan)
    return Xr / s


def crossview_consistency_confidence(Xa, Xb, sigma_3d=0.08, scale_mode="hip", eps=1e-12, **idx):
    """fuse/confidence.py:183-224 -> conf (J,), dist (J,)."""
    Xa_c = canonicalize_pose_3d(Xa, scale_mode=scale_mode, **idx)
    Xb_c = canonicalize_pose_3d(Xb, scale_mode=scale_mode, **idx)
    dist = np.full(Xa.shape[0], np.nan)
    ok = _finite_rows(Xa_c) & _finite_rows(Xb_c)
    dist[ok] = np.sqrt(((Xa_c[ok] - Xb_c[ok]) ** 2).sum(axis=1))
    sig2 = max(float(sigma_3d), eps) ** 2
    conf = np.zeros_like(dist)
    vv = np.isfinite(dist)
    conf[vv] = np.exp(-(dist[vv] ** 2) / (2.0 * sig2))
    return conf, dist


def softmax2(a, b):
    """fuse/fuse.py:87-94."""
    m = np.maximum(a, b)
    ea, eb = np.exp(a - m), np.exp(b - m)
    s = ea + eb + EPS
    return ea / s, eb / s


def fuse_frame_3d(Xl, Xr, q_l, q_r):
    """fuse/fuse.py:289-326 on (J,3) arrays in the same coordinate system."""
    ok_l, ok_r = _finite_rows(Xl), _finite_rows(Xr)
    wl, wr = softmax2(q_l, q_r)
    fused = np.full_like(Xl, np.nan)
    both = ok_l & ok_r
    fused[both] = (wl[both, None] * Xl[both] + wr[both, None] * Xr[both]) / (wl[both, None] + wr[both, None] + EPS)
    only_l, only_r = ok_l & ~ok_r, ok_r & ~ok_l
    fused[only_l] = Xl[only_l]
    fused[only_r] = Xr[only_r]
    return fused


def alpha_per_joint(target_ids, alpha=0.7, adaptive=True, alpha_min=0.45, alpha_max=0.92):
    """fuse/fuse.py:362-376."""
    a = np.full(len(target_ids), float(alpha))
    if adaptive:
        for j, jid in enumerate(target_ids):
            if jid in CORE_IDS:
                a[j] = alpha * 0.85
            elif jid in LIMB_IDS:
                a[j] = alpha * 1.00
            elif jid in ENDPOINT_IDS:
                a[j] = alpha * 1.15
        a = np.clip(a, alpha_min, alpha_max)
    return a


def temporal_smooth_ema(X, target_ids=None, alpha=0.7, adaptive=True, alpha_min=0.45, alpha_max=0.92, speed_gain=0.25):
    """fuse/fuse.py:329-412 on a (T,J,3) array (NaN rows = missing)."""
    T, J, _ = X.shape
    if T == 0:
        return X.copy()
    ids = list(range(J)) if target_ids is None else list(target_ids)
    aj = alpha_per_joint(ids, alpha, adaptive, alpha_min, alpha_max)
    Y = np.full_like(X, np.nan)
    Y[0] = np.where(_finite_rows(X[0])[:, None], X[0], np.nan)  # array_to_dict drops non-finite rows (fuse.py:76-82)
    for t in range(1, T):
        xt, yp = X[t], Y[t - 1]
        ok_x, ok_p = _finite_rows(xt), _finite_rows(yp)
        both = ok_x & ok_p
        if np.any(both):
            if adaptive:
                speed = np.linalg.norm(xt[both] - yp[both], axis=1)
                ad = np.clip(aj[both] + speed_gain * speed, alpha_min, alpha_max)
            else:
                ad = np.full(np.count_nonzero(both), float(alpha))
            Y[t, both] = ad[:, None] * xt[both] + (1.0 - ad)[:, None] * yp[both]
        Y[t, ~ok_x & ok_p] = yp[~ok_x & ok_p]
        Y[t, ok_x & ~ok_p] = xt[ok_x & ~ok_p]
    return Y


def fuse_clip(Xl, Xr, Ul, Ur, sigma_px=12.0, sigma_3d=0.08, scale_mode="hip", align=True):
    """The per-frame pipeline of fuse/main_raw.py:199-250 on (T,J,.) arrays; align=False is fuse/main_unity.py:96-132
    (both views already share a coordinate system: no rigid alignment).
    Returns fused (T,J,3), q_l (T,J), q_r (T,J), Xr_aligned (T,J,3)."""
    T, J, _ = Xl.shape
    fused, ql, qr, Xa = (np.full((T, J, 3), np.nan), np.zeros((T, J)), np.zeros((T, J)), np.full((T, J, 3), np.nan))
    for t in range(T):
        Xa[t] = align_right_to_left(Xl[t], Xr[t]) if align else Xr[t]
        c1l, _ = weakpersp_reproj_confidence(Xl[t], Ul[t], sigma_px)
        c1r, _ = weakpersp_reproj_confidence(Xr[t], Ur[t], sigma_px)
        c2, _ = crossview_consistency_confidence(Xl[t], Xr[t], sigma_3d, scale_mode)
        ql[t], qr[t] = np.sqrt(c1l * c2), np.sqrt(c1r * c2)
        fused[t] = fuse_frame_3d(Xl[t], Xa[t], ql[t], qr[t])
    return fused, ql, qr, Xa


# ---------------------------------------------------------------------------------------------------------------
# the second fusion path of the reference: bundle_adjustment/fuse/fuse.py (== fuse/side/fuse/fuse.py ==
# front_side/side/fuse/fuse.py), called per frame by bundle_adjustment/run.py:225, fuse/side/run.py:81,
# front_side/side/run.py:81
TORSO_IDX = (69, 9, 10, 5, 6)  # NECK, L_HIP, R_HIP, L_SHO, R_SHO (bundle_adjustment/fuse/fuse.py:27-31)


def estimate_rigid_umeyama(target, source, allow_scale=False):
    """bundle_adjustment/fuse/fuse_check.py:26-78 -> R, t, s with s R source + t ~ target."""
    mask = _finite_rows(target) & _finite_rows(source)
    target, source = target[mask], source[mask]
    N = target.shape[0]
    if N < 3:
        raise ValueError("need at least 3 corresponding points")
    tm, sm = target.mean(0), source.mean(0)
    tc, sc = target - tm, source - sm
    H = (sc.T @ tc) / N
    U, S, Vt = np.linalg.svd(H)
    R = Vt.T @ U.T
    if np.linalg.det(R) < 0:
        Vt[-1, :] *= -1
        R = Vt.T @ U.T
    s = S.sum() / ((sc**2).sum() / N + 1e-12) if allow_scale else 1.0
    return R, tm - s * (R @ sm), float(s)


def fuse_two(L, R_align, tau, wL, wR):
    """bundle_adjustment/fuse/fuse.py:55-93: per joint - one view missing: the other; both present and farther apart than
    tau: the more confident view (left on ties); else the confidence-weighted mean."""
    out = np.full_like(L, np.nan)
    okl, okr = _finite_rows(L), _finite_rows(R_align)
    out[okl & ~okr] = L[okl & ~okr]
    out[okr & ~okl] = R_align[okr & ~okl]
    both = okl & okr
    d = np.linalg.norm(L - R_align, axis=1)
    far = both & (d > tau)
    pickL = far & (wL >= wR)
    out[pickL] = L[pickL]
    out[far & ~pickL] = R_align[far & ~pickL]
    near = both & ~far
    out[near] = (wL[near, None] * L[near] + wR[near, None] * R_align[near]) / (wL[near, None] + wR[near, None] + 1e-9)
    return out


def rigid_transform_3D(target, source, tau=0.08, allow_scale=False, wL=None, wR=None):
    """bundle_adjustment/fuse/fuse.py:96-232 on (T,J,3) arrays -> fused (T,J,3), Rhat (T,3,3), that (T,3), s (T,),
    diag (T,4) = [LR_before, Fused_vs_L, Fused_vs_R, gain] (plain means: NaN if any joint is missing, like the reference)."""
    L, R = np.asarray(target, float), np.asarray(source, float)
    T, J, _ = L.shape
    ex = lambda w: np.ones((T, J)) if w is None else np.broadcast_to(np.asarray(w, float), (T, J))
    wLs, wRs = ex(wL), ex(wR)
    tauv = np.broadcast_to(np.asarray(tau, float), (J,))
    fused, Rh, th, sh, diag = np.empty_like(L), np.empty((T, 3, 3)), np.empty((T, 3)), np.empty(T), np.empty((T, 4))
    idx = list(TORSO_IDX)
    for t in range(T):
        Rh[t], th[t], sh[t] = estimate_rigid_umeyama(L[t][idx], R[t][idx], allow_scale)
        Ra = sh[t] * (Rh[t] @ R[t].T).T + th[t]
        fused[t] = fuse_two(L[t], Ra, tauv, wLs[t], wRs[t])
        lr = np.linalg.norm(L[t] - R[t], axis=-1).mean()
        fl = np.linalg.norm(fused[t] - L[t], axis=-1).mean()
        fr = np.linalg.norm(fused[t] - R[t], axis=-1).mean()
        diag[t] = [lr, fl, fr, lr - 0.5 * (fl + fr)]
    return fused, Rh, th, sh, diag
