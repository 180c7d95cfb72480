"""Stand-in for ``triangulation/reproject.py`` (stereo convention: R, T = cam1->cam2).

Same public names and behaviour as the reference module; the projection arithmetic
(cv2.Rodrigues + cv2.projectPoints, triangulation/reproject.py:63-83) runs in libska.so.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

from . import _common
from ._common import render_reprojection_panel  # noqa: F401  (re-exported: reproject.py:86-200)


def reproject_points(X3, K1, dist1: Optional[np.ndarray], K2, dist2: Optional[np.ndarray], R, T) -> Dict[str, np.ndarray]:
    """triangulation/reproject.py:49-83.  X3 (J,3) in cam1 coordinates; cam1 is the identity
    (rvec = tvec = 0, :64-65), cam2 = (R (3,3), T (3,)|(3,1)).  Every input is cast to float32
    first (:63-75, quirk Q2); returns {"proj_L": (J,2) f32, "proj_R": (J,2) f32}."""
    R = _common.f32(R).reshape(3, 3)
    T = _common.f32(T).reshape(3, 1)
    return _common.reproject_pair(X3, K1, dist1, K2, dist2, R, T)


def reproject_and_visualize(img1, img2, X3, kptL, kptR, K1, dist1, K2, dist2, R, T,
                            joint_names: Optional[Sequence[str]] = None, circle_r: int = 5, thickness: int = 2,
                            out_path: str = "/mnt/data/reprojection_compare.jpg") -> Dict[str, object]:
    """triangulation/reproject.py:203-266: reproject, draw, save, return the same dict
    (proj_L/R, err_L/R, rmse_*, mean_err_*, median_err_*, max_err_*, out_path, vis_left/right, panel)."""
    R_ = _common.f32(R).reshape(3, 3)
    T_ = _common.f32(T).reshape(3, 1)
    proj = _common.reproject_pair(X3, K1, dist1, K2, dist2, R_, T_, kptL, kptR)
    return _common.visualize(img1, img2, proj, kptL, kptR, joint_names, circle_r, thickness, out_path)
