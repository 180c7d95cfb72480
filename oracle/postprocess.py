"""fp64 restatement of the reference's post-triangulation triage (row N2 of SURVEY.md section 8f).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows triangulation/postprocess.py of the
reference (file:line cited per function) and restates the two third-party algorithms it delegates to:
  * cv2.undistortPoints(x, K, d, P=K)   (opencv-python, unpinned in requirements.txt:28; 4.13.0 here):
    normalise with K, five fixed-point iterations of the inverse distortion (the default termination
    criterion of the 5-argument overload is MAX_ITER = 5, no epsilon test), re-project with P.  Output is
    float32 when the input pixels are float32 - which they are on the reference's path.
  * scipy.signal.savgol_filter(v, window_length, polyorder, mode="interp")   (scipy, not even listed in
    requirements.txt; 1.18.1 here): least-squares polynomial smoothing, the first / last window//2
    samples evaluated from a polynomial fitted to the first / last window samples.
Pinned by tests/test_oracle_postprocess.py against cv2 / scipy directly and against the reference's own
post_triage_sequence outputs frozen in tests/golden/g7_post_triage.npz.
"""
from __future__ import annotations

import numpy as np


def undistort_points(x, K, dist, iters: int = 5):
    """cv2.undistortPoints(x, K, dist, P=K) -> (N,2) float32 for float32 input (postprocess.py:93-98)."""
    x = np.asarray(x)
    out_dtype = np.float32 if x.dtype == np.float32 else np.float64
    K = np.asarray(K, np.float64)
    k = np.zeros(14)
    d = np.asarray(dist, np.float64).reshape(-1)
    k[: d.size] = d
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    u = x.astype(np.float64)
    x0 = (u[:, 0] - cx) / fx
    y0 = (u[:, 1] - cy) / fy
    xx, yy = x0.copy(), y0.copy()
    alive = np.ones(len(xx), bool)  # cv2 stops a point (and resets it) when the inverse radial factor turns negative
    for _ in range(iters):
        r2 = xx * xx + yy * yy
        icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2)
        bad = alive & (icdist < 0)
        xx = np.where(bad, x0, xx)
        yy = np.where(bad, y0, yy)
        alive &= ~bad
        dx = 2 * k[2] * xx * yy + k[3] * (r2 + 2 * xx * xx) + k[8] * r2 + k[9] * r2 * r2
        dy = k[2] * (r2 + 2 * yy * yy) + 2 * k[3] * xx * yy + k[10] * r2 + k[11] * r2 * r2
        xn = (x0 - dx) * icdist
        yn = (y0 - dy) * icdist
        xx = np.where(alive, xn, xx)
        yy = np.where(alive, yn, yy)
    return np.stack([fx * xx + cx, fy * yy + cy], -1).astype(out_dtype)


def project(P, X3):
    """postprocess.py:32-35: pinhole projection with the 1e-12 guard in the denominator."""
    Xh = np.hstack([X3, np.ones((X3.shape[0], 1))])
    Y = Xh @ np.asarray(P).T
    return Y[:, :2] / (Y[:, 2:3] + 1e-12)


def post_triage_single(X3, kptL, kptR, K1, K2, R, T, dist1=None, dist2=None, confL=None, confR=None, conf_thr=0.3,
                       err_thresh_px=2.0):
    """postprocess.py:71-125 -> (X_clean (J,3), report dict, keep (J,) bool, em (J,))."""
    x1 = kptL if dist1 is None else undistort_points(kptL, K1, dist1)
    x2 = kptR if dist2 is None else undistort_points(kptR, K2, dist2)
    P1 = np.asarray(K1) @ np.hstack([np.eye(3), np.zeros((3, 1))])
    P2 = np.asarray(K2) @ np.hstack([np.asarray(R), np.asarray(T).reshape(3, 1)])
    e1 = np.linalg.norm(project(P1, X3) - x1, axis=1)
    e2 = np.linalg.norm(project(P2, X3) - x2, axis=1)
    em = 0.5 * (e1 + e2)
    pos = (X3[:, 2] > 0) & (((X3 @ np.asarray(R).T) + np.asarray(T).reshape(1, 3))[:, 2] > 0)
    cm = np.ones(len(X3), bool)
    if confL is not None and confR is not None:
        cm = (confL >= conf_thr) & (confR >= conf_thr)
    keep = pos & np.isfinite(em) & (em <= err_thresh_px) & cm
    Xc = X3.copy()
    Xc[~keep] = np.nan
    with np.errstate(all="ignore"):
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rep = {"rmse_px": float(np.sqrt(np.nanmean(em**2))), "median_err_px": float(np.nanmedian(em)),
                   "pos_depth_ratio": float(np.mean(pos)), "kept_ratio": float(np.mean(keep)), "kept_count": int(keep.sum())}
    return Xc, rep, keep, em


def savgol_coeffs(win: int, poly: int):
    """Interior weights (win,) and the edge matrices of mode='interp': E_first (h, win) maps the first `win`
    samples to the first h = win//2 outputs, E_last (h, win) the last `win` samples to the last h outputs."""
    h = win // 2
    z = np.arange(win, dtype=np.float64)
    A = np.vander(z, poly + 1, increasing=True)          # (win, poly+1): fit a polynomial in the sample index
    pinv = np.linalg.pinv(A)                              # (poly+1, win)
    ev = lambda pos: np.vander(np.atleast_1d(np.asarray(pos, np.float64)), poly + 1, increasing=True) @ pinv
    centre = ev(h)[0]
    return centre, ev(np.arange(h)), ev(np.arange(win - h, win))


def savgol_interp(v, win: int, poly: int):
    """scipy.signal.savgol_filter(v, win, poly) with the default mode='interp' for a 1-D array, len(v) >= win."""
    v = np.asarray(v, np.float64)
    n, h = len(v), win // 2
    c, Ef, El = savgol_coeffs(win, poly)
    out = np.empty(n)
    out[h: n - h] = np.convolve(v, c[::-1], mode="valid")
    out[:h] = Ef @ v[:win]
    out[n - h:] = El @ v[n - win:]
    return out


def effective_window(T: int, win: int) -> int:
    """postprocess.py:58: odd window, capped by the clip length."""
    return min(win if win % 2 == 1 else win + 1, max(1 if T % 2 == 1 else T - 1, 3))


def smooth_skeleton(X, win=9, poly=2):
    """postprocess.py:54-68: per (joint, coordinate) series, the FINITE samples are compacted, smoothed and
    scattered back; series with fewer finite samples than the window stay as they are."""
    Xs = X.copy()
    T, J, C = X.shape
    win = effective_window(T, win)
    for j in range(J):
        for c in range(C):
            vec = X[:, j, c]
            m = np.isfinite(vec)
            if m.sum() >= win:
                v = vec.copy()
                v[m] = savgol_interp(vec[m], win, poly)
                Xs[:, j, c] = v
    return Xs


def post_triage_sequence(X3_seq, kptL_seq, kptR_seq, K1, K2, R, T, dist1=None, dist2=None, confL=None, confR=None,
                         conf_thr=0.3, err_thresh_px=2.0, smooth=False, sg_win=9, sg_poly=2):
    """postprocess.py:129-170 -> (X_clean (T,J,3) float32, list of per-frame reports)."""
    Tn = X3_seq.shape[0]
    Xc = np.full_like(X3_seq, np.nan, dtype=np.float32)
    stats = []
    for t in range(Tn):
        cl, rep, _, _ = post_triage_single(X3_seq[t], kptL_seq[t], kptR_seq[t], K1, K2, R, T, dist1, dist2,
                                           None if confL is None else confL[t], None if confR is None else confR[t],
                                           conf_thr, err_thresh_px)
        Xc[t] = cl
        stats.append(rep)
    if smooth:
        Xc = smooth_skeleton(Xc, win=sg_win, poly=sg_poly)
    return Xc, stats
