"""Stand-in for ``triangulation/postprocess.py`` (post-triangulation triage + Savitzky-Golay smoothing).

Same public names and behaviour; ``post_triage_sequence`` - which receives the whole clip - runs ONE triage
launch (+ the smoothing passes) on the GPU instead of a Python loop over frames.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _common


def build_P(K, R=np.eye(3), t=np.zeros(3)):
    """postprocess.py:28-29 (host arithmetic: 36 flops)."""
    return np.asarray(K) @ np.hstack([np.asarray(R), np.asarray(t).reshape(3, 1)])


def project(P, X3):
    """postprocess.py:32-35: (N,3) -> (N,2) with the 1e-12 guard; GPU reprojection (ska_reproject_points_f32
    takes K [R|t]; an arbitrary P is passed as K = I, [R|t] = P)."""
    from .. import api

    dev = _common.device()
    P = np.asarray(P, np.float64)
    X = torch.from_numpy(np.ascontiguousarray(np.asarray(X3, np.float32).reshape(1, -1, 3))).to(dev)
    proj, _ = api.reproject_points(X, np.eye(3), P[None, :, :3], P[None, :, 3], None)
    return proj[0, 0].cpu().numpy().astype(np.float64)


def reproj_errors(P1, P2, X3, x1_pix, x2_pix):
    """postprocess.py:38-44 -> (e1, e2, 0.5 (e1 + e2))."""
    p1, p2 = project(P1, X3), project(P2, X3)
    e1 = np.linalg.norm(p1 - x1_pix, axis=1)
    e2 = np.linalg.norm(p2 - x2_pix, axis=1)
    return e1, e2, 0.5 * (e1 + e2)


def positive_depth_mask(R, T, X3):
    """postprocess.py:47-52: depth > 0 in the left camera (= world) and in the right camera."""
    X3 = np.asarray(X3)
    z2 = (X3 @ np.asarray(R).T + np.asarray(T).reshape(1, 3))[:, 2]
    return (X3[:, 2] > 0) & (z2 > 0)


def smooth_skeleton(X, win=9, poly=2):
    """postprocess.py:54-68: X (T,J,3) -> smoothed array of the same shape and dtype."""
    from .. import post

    X = np.asarray(X)
    out = post.smooth_skeleton(torch.from_numpy(np.ascontiguousarray(X, np.float32)).to(_common.device()), win=win, poly=poly)
    return out.cpu().numpy().astype(X.dtype if X.dtype.kind == "f" else np.float32)


def _run(X3, kL, kR, K1, K2, R, T, dist1, dist2, confL, confR, conf_thr, err_thresh_px):
    from .. import post

    dev = _common.device()
    X = torch.from_numpy(np.ascontiguousarray(X3, np.float32)).to(dev)
    k = torch.from_numpy(np.ascontiguousarray(np.stack([np.asarray(kL, np.float32), np.asarray(kR, np.float32)]))).to(dev)
    c = None
    if confL is not None and confR is not None:
        c = torch.from_numpy(np.ascontiguousarray(np.stack([np.asarray(confL, np.float32), np.asarray(confR, np.float32)]))).to(dev)
    return post.post_triage(X, k, K1, K2, R, T, dist1, dist2, conf=c, conf_thr=conf_thr, err_thresh_px=err_thresh_px)


def _report(row) -> dict:
    return {"rmse_px": float(row[0]), "median_err_px": float(row[1]), "pos_depth_ratio": float(row[2]),
            "kept_ratio": float(row[3]), "kept_count": int(row[4])}


def post_triage_single(X3_frame, kptL_frame, kptR_frame, K1, K2, R, T, dist1=None, dist2=None, confL=None, confR=None,
                       conf_thr=0.3, err_thresh_px=2.0, return_masks=False):
    """postprocess.py:71-125: one frame (J,3) -> (X3_clean, report[, keep])."""
    X3_frame = np.asarray(X3_frame)
    res = _run(X3_frame[None], np.asarray(kptL_frame)[None], np.asarray(kptR_frame)[None], K1, K2, R, T, dist1, dist2,
               None if confL is None else np.asarray(confL)[None], None if confR is None else np.asarray(confR)[None],
               conf_thr, err_thresh_px)
    keep = (res.flags[0].cpu().numpy() & 8) != 0
    Xc = X3_frame.copy()
    Xc[~keep] = np.nan
    rep = _report(res.report[0].cpu().numpy())
    return (Xc, rep, keep) if return_masks else (Xc, rep)


def post_triage_sequence(X3_seq, kptL_seq, kptR_seq, K1, K2, R, T, dist1=None, dist2=None, confL=None, confR=None,
                         conf_thr=0.3, err_thresh_px=2.0, smooth=False, sg_win=9, sg_poly=2):
    """postprocess.py:129-170: (T,J,3) -> (X_clean (T,J,3) float32, list of per-frame reports)."""
    from .. import post

    res = _run(np.asarray(X3_seq), kptL_seq, kptR_seq, K1, K2, R, T, dist1, dist2, confL, confR, conf_thr, err_thresh_px)
    Xc = res.X_clean
    if smooth:
        Xc = post.smooth_skeleton(Xc, win=sg_win, poly=sg_poly)
    rep = res.report.cpu().numpy()
    return Xc.cpu().numpy().astype(np.float32), [_report(r) for r in rep]
