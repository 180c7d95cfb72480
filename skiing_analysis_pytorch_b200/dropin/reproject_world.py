"""Stand-in for ``bundle_adjustment/reproject.py`` and its copies ``vggt/reproject.py``,
``front_side/side/reproject.py``, ``fuse/side/reproject.py`` (world->camera or relative extrinsics).

Same public names and behaviour; projection arithmetic in libska.so.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Optional, Sequence

import numpy as np

from . import _common
from ._common import render_reprojection_panel  # noqa: F401  (bundle_adjustment/reproject.py:156-278)


def _relative_pose(R, T):
    """bundle_adjustment/reproject.py:99-132.  Mode A: (2,3,3)/(2,3) world->camera pairs ->
    R_rel = R2 R1^T, t_rel = t2 - R_rel t1 (float32, like the reference).  Mode B: (3,3) + (3,)|(3,1)
    relative pose.  ValueError on anything else."""
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
    return R_rel, t_rel.astype(np.float32)


def reproject_points(X3, K1, dist1: Optional[np.ndarray], K2, dist2: Optional[np.ndarray], R, T) -> Dict[str, np.ndarray]:
    """bundle_adjustment/reproject.py:74-153.  cam1 is ALWAYS the identity (:135-136): X3 is taken
    as cam1 coordinates even in mode A, where it is only correct if R[0] = I, t[0] = 0 (quirk Q3,
    reproduced).  Returns {"proj_L", "proj_R"} (J,2) float32."""
    R_rel, t_rel = _relative_pose(R, T)
    return _common.reproject_pair(X3, K1, dist1, K2, dist2, R_rel, t_rel)


def reproject_and_visualize(img1, img2, X3, kptL, kptR, K1, dist1, K2, dist2, R, T,
                            joint_names: Optional[Sequence[str]] = None, circle_r: int = 5, thickness: int = 2,
                            out_path: Path = Path("reprojection_panel.jpg")) -> Dict[str, object]:
    """bundle_adjustment/reproject.py:281-350."""
    R_rel, t_rel = _relative_pose(R, T)
    proj = _common.reproject_pair(X3, K1, dist1, K2, dist2, R_rel, t_rel, kptL, kptR)
    return _common.visualize(img1, img2, proj, kptL, kptR, joint_names, circle_r, thickness, out_path)
