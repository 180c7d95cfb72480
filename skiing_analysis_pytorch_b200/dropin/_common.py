"""Helpers shared by the drop-in shims: device choice, host<->device marshalling for the per-frame
signatures (J <= 70 points per call: a GPU round trip, kept for API parity - throughput comes from
the clip-level entry points), the reprojection panel and its statistics."""
from __future__ import annotations

from pathlib import Path
from typing import Optional, Sequence

import warnings

import numpy as np
import torch


def device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: skiing_analysis_pytorch_b200 has no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def f32(a) -> np.ndarray:
    """The reference's _as_float32 (triangulation/reproject.py:63-75): every input is down-cast."""
    return np.asarray(a, dtype=np.float32)


def rotation_as_projectpoints_sees_it(R) -> np.ndarray:
    """The rotation cv2.projectPoints actually applies when the reference hands it R: reproject.py:69 converts the float32
    matrix to a rotation VECTOR (cv2.Rodrigues: nearest rotation U V^T by SVD, then axis * angle, stored as float32) and
    projectPoints turns that vector back into a matrix in double.  For an exactly orthonormal R this is R to ~1e-7; for
    R2 @ R1.T formed in float32 or a network-predicted matrix it is the projection onto SO(3), not R itself."""
    R32 = np.asarray(R, np.float32).reshape(3, 3).astype(np.float64)
    U, _, Vt = np.linalg.svd(R32)
    Ro = U @ Vt
    r = np.array([Ro[2, 1] - Ro[1, 2], Ro[0, 2] - Ro[2, 0], Ro[1, 0] - Ro[0, 1]])
    s = np.sqrt((r @ r) * 0.25)
    c = np.clip((np.trace(Ro) - 1.0) * 0.5, -1.0, 1.0)
    theta = np.arccos(c)
    if s < 1e-5:
        if c > 0:
            rvec = np.zeros(3)
        else:  # rotation by pi: axis from the diagonal, signs from the off-diagonal entries (OpenCV's branch)
            t = (Ro[0, 0] + 1) * 0.5
            rx = np.sqrt(max(t, 0.0))
            t = (Ro[1, 1] + 1) * 0.5
            ry = np.sqrt(max(t, 0.0)) * (-1.0 if Ro[0, 1] < 0 else 1.0)
            t = (Ro[2, 2] + 1) * 0.5
            rz = np.sqrt(max(t, 0.0)) * (-1.0 if Ro[0, 2] < 0 else 1.0)
            if abs(rx) < abs(ry) and abs(rx) < abs(rz) and (Ro[1, 2] > 0) != (ry * rz > 0):
                rz = -rz
            v = np.array([rx, ry, rz])
            rvec = v * (theta / max(np.linalg.norm(v), 1e-300))
    else:
        rvec = r * (theta / (2.0 * s))
    rvec = rvec.astype(np.float32).astype(np.float64)
    th = np.linalg.norm(rvec)
    if th < 2.220446049250313e-16:
        return np.eye(3)
    k = rvec / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(k, k) + np.sin(th) * Kx


def nan_stats(err: np.ndarray) -> dict:
    """rmse / mean / median / max exactly as reproject.py:254-261 (nan-aware, float64)."""
    return {
        "rmse": float(np.sqrt(np.nanmean(err**2))),
        "mean": float(np.nanmean(err)),
        "median": float(np.nanmedian(err)),
        "max": float(np.nanmax(err)),
    }


def _bgr_u8(img) -> np.ndarray:
    im = np.asarray(img.cpu() if torch.is_tensor(img) else img)
    if im.ndim == 2:
        im = np.stack([im] * 3, axis=-1)
    elif im.ndim == 3 and im.shape[2] == 1:
        im = np.repeat(im, 3, axis=2)
    elif not (im.ndim == 3 and im.shape[2] == 3):
        raise ValueError(f"Unsupported image shape: {im.shape}")
    if im.dtype != np.uint8:
        im = np.clip(im, 0, 255)
        im = (im * 255.0).astype(np.uint8) if im.max() <= 1.0 else im.astype(np.uint8)
    return np.ascontiguousarray(im)


def render_reprojection_panel(
    img1,
    img2,
    kptL,
    kptR,
    proj_L,
    proj_R,
    joint_names: Optional[Sequence[str]] = None,
    circle_r: int = 5,
    thickness: int = 2,
    align_height: bool = True,
    title_left: str = "Left/Cam1 (Green=Observed, Red=Reprojected)",
    title_right: str = "Right/Cam2",
):
    """Host-side visualisation (NOT accelerated): observed keypoints in green,
    reprojections in red, error vectors in cyan, both views side by side with an RMSE banner.
    Same signature and return value (vis_left, vis_right, panel) as
    triangulation/reproject.py:86-200 / bundle_adjustment/reproject.py:156-278."""
    import cv2

    views = [_bgr_u8(img1), _bgr_u8(img2)]
    if align_height and views[0].shape[0] != views[1].shape[0]:
        h = max(v.shape[0] for v in views)
        views = [v if v.shape[0] == h else cv2.resize(v, (int(v.shape[1] * h / v.shape[0]), h), interpolation=cv2.INTER_LINEAR)
                 for v in views]
    green, red, cyan, white = (0, 255, 0), (0, 0, 255), (255, 255, 0), (255, 255, 255)
    font = cv2.FONT_HERSHEY_SIMPLEX
    drawn, banners = [], []
    for vis, obs, rep, title in zip(views, (kptL, kptR), (proj_L, proj_R), (title_left, title_right)):
        vis = vis.copy()
        h, w = vis.shape[:2]
        obs = np.asarray(obs, dtype=float).reshape(-1, 2)
        rep = np.asarray(rep, dtype=float).reshape(-1, 2)
        rep_c = np.stack([np.clip(rep[:, 0], 0, w - 1), np.clip(rep[:, 1], 0, h - 1)], axis=1)
        for j in range(len(obs)):
            if not (np.isfinite(obs[j]).all() and np.isfinite(rep_c[j]).all()):
                continue
            o = (int(round(obs[j, 0])), int(round(obs[j, 1])))
            r = (int(round(rep_c[j, 0])), int(round(rep_c[j, 1])))
            cv2.circle(vis, o, circle_r, green, thickness, cv2.LINE_AA)
            cv2.circle(vis, r, circle_r, red, thickness, cv2.LINE_AA)
            cv2.line(vis, o, r, cyan, 1, cv2.LINE_AA)
            name = str(joint_names[j]) if joint_names is not None and j < len(joint_names) else str(j)
            cv2.putText(vis, name, (o[0] + 6, o[1] - 6), font, 0.45, green, 1, cv2.LINE_AA)
        e = np.linalg.norm(rep - obs, axis=1)
        s = nan_stats(e)
        banners.append(f"{title} | RMSE={s['rmse']:.2f}px  (mean={s['mean']:.2f}, med={s['median']:.2f}, max={s['max']:.2f})")
        drawn.append(vis)
    visL, visR = drawn
    panel = np.zeros((max(visL.shape[0], visR.shape[0]), visL.shape[1] + visR.shape[1], 3), np.uint8)
    panel[: visL.shape[0], : visL.shape[1]] = visL
    panel[: visR.shape[0], visL.shape[1]:] = visR
    cv2.putText(panel, banners[0], (20, 30), font, 0.7, white, 2, cv2.LINE_AA)
    cv2.putText(panel, banners[1], (visL.shape[1] + 20, 30), font, 0.7, white, 2, cv2.LINE_AA)
    return visL, visR, panel


def visualize(img1, img2, proj, kptL, kptR, joint_names, circle_r, thickness, out_path) -> dict:
    """Panel + statistics + imwrite; returns the dict of reproject.py:249-266 (same keys).
    `proj` comes from reproject_pair(..., kptL, kptR): projections, per-joint errors and the nan-aware
    scalars were all computed on the GPU; the host only draws."""
    import cv2

    visL, visR, panel = render_reprojection_panel(img1, img2, kptL, kptR, proj["proj_L"], proj["proj_R"],
                                                  joint_names=joint_names, circle_r=circle_r, thickness=thickness)
    res = dict(proj)
    Path(out_path).parent.mkdir(parents=True, exist_ok=True)
    cv2.imwrite(str(out_path), panel)
    res.update(out_path=str(out_path), vis_left=visL, vis_right=visR, panel=panel)
    return res


def reproject_pair(X3, K1, dist1, K2, dist2, R_rel, t_rel, kptL=None, kptR=None) -> dict:
    """cam1 = identity with (K1, dist1), cam2 = (R_rel, t_rel) with (K2, dist2); float32 in,
    float32 (J,2) out - the arithmetic of reproject_points on the GPU (ska_reproject_points_f32).
    With the observed keypoints the per-joint errors and the nan-aware rmse / mean / median / max come back too, formed
    like reproject.py:242-261 (float64, from the float32 projections: quirk Q6)."""
    from .. import api

    dev = device()
    X = torch.from_numpy(np.ascontiguousarray(f32(X3).reshape(1, -1, 3))).to(dev)
    K = np.stack([f32(K1).reshape(3, 3), f32(K2).reshape(3, 3)]).astype(np.float64)
    R = np.stack([np.eye(3), rotation_as_projectpoints_sees_it(R_rel)])
    t = np.stack([np.zeros(3), f32(t_rel).reshape(3).astype(np.float64)])
    dists = [None if d is None else f32(d).reshape(-1).astype(np.float64) for d in (dist1, dist2)]
    if kptL is None:
        proj, _ = api.reproject_points(X, K, R, t, dists, want=("proj",))
        p = proj.cpu().numpy()
        return {"proj_L": p[0, 0], "proj_R": p[1, 0]}
    # The projections are the GPU's (float32, as cv2.projectPoints returns them for float32 input).  The per-joint errors and
    # their nan-aware statistics are J numbers per view: formed here exactly as reproject.py:242-261 forms them - float64
    # arithmetic on float32 projections minus the keypoints AT THEIR OWN PRECISION (np.asarray(kpt, float)) - so float64
    # keypoints keep their last bits (quirk Q6).  Whole clips use api.reproject_points / api.frame_stats (float32 on the GPU).
    proj, _ = api.reproject_points(X, K, R, t, dists, want=("proj",))
    p = proj.cpu().numpy()
    out = {"proj_L": p[0, 0], "proj_R": p[1, 0]}
    with np.errstate(invalid="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)  # all-NaN frames: the reference's nanmean warns and returns nan too
        for v, (side, kpt) in enumerate((("L", kptL), ("R", kptR))):
            e = np.linalg.norm(p[v, 0] - np.asarray(kpt, float).reshape(-1, 2), axis=1)
            out[f"err_{side}"] = e
            out[f"rmse_{side}"] = float(np.sqrt(np.nanmean(e**2)))
            out[f"mean_err_{side}"] = float(np.nanmean(e))
            out[f"median_err_{side}"] = float(np.nanmedian(e))
            out[f"max_err_{side}"] = float(np.nanmax(e))
    return out
