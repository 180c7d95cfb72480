"""Stand-in for ``vggt/triangulate.py`` (two world->camera cameras, per-camera K)."""
from __future__ import annotations

import logging
import os
from pathlib import Path

import numpy as np
import torch

from . import _common
from .reproject_world import reproject_and_visualize  # noqa: F401  (vggt/triangulate.py:8)

logger = logging.getLogger(__name__)


def make_P(K, R, t):
    """vggt/triangulate.py:13-16: K (3,3), R (3,3), t (3,) -> P (3,4).  (Host arithmetic: 36 flops.)"""
    return np.asarray(K) @ np.concatenate([np.asarray(R), np.asarray(t).reshape(3, 1)], axis=1)


def triangulate_point(P1, P2, x1, x2):
    """vggt/triangulate.py:19-34: two-view DLT of one point from arbitrary 3x4 P -> (3,) float64.
    Runs the kernel's fp64 Jacobi solver (the P matrices are passed through as K = I, [R|t] = P)."""
    from .. import api

    dev = _common.device()
    P = np.stack([np.asarray(P1, np.float64), np.asarray(P2, np.float64)])
    k = np.asarray([x1, x2], np.float32).reshape(2, 1, 1, 2)
    res = api.triangulate_reproject(torch.from_numpy(k).to(dev), np.eye(3), P[:, :, :3], P[:, :, 3], solver="jacobi64",
                                    centre=(0.0, 0.0, 0.0), pinhole_reproj=True, want=("X",))
    return res.X[0, 0].cpu().numpy().astype(np.float64)


def triangulate_frame(K, R, T, kptL, kptR) -> np.ndarray:
    """The joint loop of vggt/triangulate.py:64-71 as one launch -> (J,3) float32 (quirk Q4)."""
    from .. import api

    dev = _common.device()
    k = np.stack([np.asarray(kptL, np.float32), np.asarray(kptR, np.float32)])[:, None]
    res = api.triangulate_reproject(torch.from_numpy(np.ascontiguousarray(k)).to(dev), np.asarray(K, np.float64),
                                    np.asarray(R, np.float64), np.asarray(T, np.float64), pinhole_reproj=True, want=("X",))
    return res.X[0].cpu().numpy()


def triangulate_one_frame(K, R, T, kptL, kptR, frame_L: np.ndarray = None, frame_R: np.ndarray = None,
                          save_dir: Path = Path("./output/triangulation"), dist=None, visualize_3d=False,
                          frame_num: int = None):
    """vggt/triangulate.py:38-132.  K (2,3,3), R (2,3,3), T (2,3) world->camera; kptL/kptR (J,2).
    Returns (X3d (J,3) float32, res) with res the reprojection dict (mode A: quirk Q3 reproduced)."""
    assert K.shape == (2, 3, 3)
    assert R.shape == (2, 3, 3)
    assert T.shape == (2, 3)
    H, W, C = frame_L.shape  # the reference dereferences frame_L unconditionally (:62)

    X3d = triangulate_frame(K, R, T, kptL, kptR)

    if save_dir:
        os.makedirs(save_dir, exist_ok=True)
        np.save(os.path.join(save_dir, "triangulated_3d.npy"), X3d)
        logger.info(f"[3D Saved] triangulated_3d.npy | shape={X3d.shape}")

    if frame_L is not None and frame_R is not None:
        res = reproject_and_visualize(img1=frame_L, img2=frame_R, X3=X3d, kptL=kptL, kptR=kptR, K1=K[0], K2=K[1], dist1=dist,
                                      dist2=dist, R=R, T=T, out_path=save_dir / "reprojection.jpg")
        if save_dir:
            with open(os.path.join(save_dir, "reprojection_error.txt"), "a") as f:
                f.write("Reprojection Error (in pixels):\n")
                f.write(f"Mean Reprojection Error Left: {res['mean_err_L']:.4f} px\n")
                f.write(f"Mean Reprojection Error Right: {res['mean_err_R']:.4f} px\n")
        logger.info(f"[Reproj] L={res['mean_err_L']:.2f}px  R={res['mean_err_R']:.2f}px")

    if visualize_3d and save_dir:
        try:  # the reference's own matplotlib helpers; optional outside a reference checkout
            from vggt.vis.pose_visualization import save_stereo_pose_frame, visualize_3d_joints
        except Exception:  # pragma: no cover
            visualize_3d_joints = save_stereo_pose_frame = None
        if visualize_3d_joints is not None:
            visualize_3d_joints(R=R, T=T, K=K, joints_3d=X3d, save_path=save_dir / "3d_joints.png",
                                title="3D Triangulated Result", image_size=(W, H))
            save_stereo_pose_frame(R=R, T=T, K=K, img_left=frame_L, img_right=frame_R, kpt_left=kptL, kpt_right=kptR,
                                   pose_3d=X3d, output_path=save_dir / "stereo_pose_frame.jpg", repoj_error=res,
                                   frame_num=frame_num)
    return X3d, res
