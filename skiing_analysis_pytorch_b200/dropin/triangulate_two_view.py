"""Stand-in for ``triangulation/triangulate.py``.

``triangulate_joints`` keeps the per-frame signature; ``process_triangulate`` - which already
receives the whole clip - runs ONE fused triangulate+reproject launch with per-frame extrinsics
(ska_triangulate_reproject_frames_f32) and then only loops for the host-side drawing.
"""
from __future__ import annotations

import logging
import os

import numpy as np
import torch

from . import _common
from .reproject_stereo import reproject_and_visualize  # noqa: F401  (triangulate.py:31 imports it here)

logger = logging.getLogger(__name__)

# triangulation/triangulate.py:39-56 == camera_calibration/calibration_parameters.npz["dist_coeffs"] (quirk Q5)
K_dist = np.array(
    [-1.1940477842823853, -15.440461757486913, 0.00013163161053023783, 0.00019082529328353381, 98.843073622415901,
     -1.3588290520381034, -14.555841222727574, 96.219667412855202, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]
)


def triangulate_joints(keypoints1, keypoints2, K, R, T):
    """triangulation/triangulate.py:60-68: P1 = K [I|0], P2 = K [R|T], DLT, dehomogenise.
    (J,2) + (J,2) -> (J,3); output dtype follows the keypoint dtype; ValueError on shape mismatch.
    float64 keypoints take the kernel's fp64 Jacobi solver so the result keeps fp64 accuracy."""
    keypoints1 = np.asarray(keypoints1)
    keypoints2 = np.asarray(keypoints2)
    if keypoints1.shape != keypoints2.shape or keypoints1.shape[1] != 2:
        raise ValueError(f"Keypoints shape mismatch: {keypoints1.shape} vs {keypoints2.shape}")
    from .. import api

    dev = _common.device()
    out_dtype = keypoints1.dtype if keypoints1.dtype.kind == "f" else np.float64
    k = np.stack([keypoints1, keypoints2]).astype(np.float32)[:, None]  # (2,1,J,2)
    Rv = np.stack([np.eye(3), np.asarray(R, np.float64).reshape(3, 3)])
    tv = np.stack([np.zeros(3), np.asarray(T, np.float64).reshape(3)])
    solver = "jacobi64" if out_dtype == np.float64 else "secular"
    res = api.triangulate_reproject(torch.from_numpy(np.ascontiguousarray(k)).to(dev), np.asarray(K, np.float64), Rv, tv,
                                    solver=solver, pinhole_reproj=True, want=("X",))
    return res.X[0].cpu().numpy().astype(out_dtype)


def triangulate_clip(left_kpts, right_kpts, K, R, T, dist=K_dist):
    """The arithmetic of process_triangulate for a whole clip in one launch: (T,J,2) x2 keypoints,
    per-frame R[i] (3,3) / T[i] (3,)|(3,1) -> X (T,J,3) f32, err (2,T,J) f32, proj (2,T,J,2) f32, all
    as CUDA tensors (reprojection WITH `dist`, triangulation of raw pixels: quirk Q1)."""
    from .. import api

    dev = _common.device()
    kL = torch.as_tensor(np.asarray(left_kpts, np.float32))
    kR = torch.as_tensor(np.asarray(right_kpts, np.float32))
    n = min(len(kL), len(kR), len(R), len(T))  # zip() semantics of triangulate.py:76-78
    k = torch.stack([kL[:n], kR[:n]]).to(dev)
    Rf = np.zeros((n, 2, 3, 3))
    tf = np.zeros((n, 2, 3))
    Rf[:, 0] = np.eye(3)
    Rf[:, 1] = np.asarray([np.asarray(r, np.float64).reshape(3, 3) for r in R[:n]])
    tf[:, 1] = np.asarray([np.asarray(t, np.float64).reshape(3) for t in T[:n]])
    return api.triangulate_reproject(k, np.asarray(K, np.float64), Rf, tf, dist=dist, want=("X", "err", "proj"))


def process_triangulate(left_kpts, right_kpts, left_vframes, right_vframes, K, R, T, output_path):
    """triangulation/triangulate.py:71-118: returns the list of per-frame (J,3) joints; per frame
    it also writes the 3-D plot (reference's own matplotlib helper, when importable) and the
    reprojection panel, and logs the mean errors.  The numbers come from one fused GPU launch."""
    from .. import api

    res = triangulate_clip(left_kpts, right_kpts, K, R, T)
    X = res.X.cpu().numpy()
    proj = res.proj.cpu().numpy()
    err = res.err.cpu().numpy().astype(np.float64)
    stats = api.frame_stats(res.err).cpu().numpy().astype(np.float64)  # (T,2,4): rmse, mean, median, max per frame and view
    try:  # the reference's own viz module (matplotlib); optional
        from triangulation.vis.pose_visualization import visualize_3d_joints
    except Exception:  # pragma: no cover - not installed outside a reference checkout
        visualize_3d_joints = None
    joints_3d_all = []
    n = min(len(X), len(left_vframes), len(right_vframes))
    for i in range(n):
        l_frame, r_frame = left_vframes[i], right_vframes[i]
        W, H = l_frame.shape[1], l_frame.shape[0]
        joints_3d = X[i].astype(np.asarray(left_kpts[i]).dtype if np.asarray(left_kpts[i]).dtype.kind == "f" else np.float32)
        if visualize_3d_joints is not None:
            visualize_3d_joints(joints_3d=joints_3d, R=R[i], T=T[i], K=K, image_size=(W, H),
                                save_path=os.path.join(output_path, f"frame_{i:04d}.png"), title=f"Frame {i} - 3D Joints",
                                y_up=True)
        pr = {"proj_L": proj[0, i], "proj_R": proj[1, i], "err_L": err[0, i], "err_R": err[1, i]}
        for v, side in enumerate("LR"):
            pr.update({f"rmse_{side}": float(stats[i, v, 0]), f"mean_err_{side}": float(stats[i, v, 1]),
                       f"median_err_{side}": float(stats[i, v, 2]), f"max_err_{side}": float(stats[i, v, 3])})
        out = _common.visualize(l_frame, r_frame, pr, left_kpts[i], right_kpts[i], None, 5, 2,
                                os.path.join(output_path, "reproj", f"{i:04d}.jpg"))
        logger.info(f"Saved to: {out['out_path']}")
        logger.info(f"Reprojection error - Frame {i}: Left {out['mean_err_L']:.2f}px, Right {out['mean_err_R']:.2f}px")
        joints_3d_all.append(joints_3d)
    return joints_3d_all
