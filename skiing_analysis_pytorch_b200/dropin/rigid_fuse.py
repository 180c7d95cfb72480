"""Stand-in for ``bundle_adjustment/fuse/fuse.py`` (== ``fuse/side/fuse/fuse.py``, ``front_side/side/fuse/fuse.py``).

The pipelines import one name from it - ``rigid_transform_3D`` (bundle_adjustment/run.py:28, fuse/side/run.py:25,
front_side/side/run.py:25); it runs on the GPU here (fusion.rigid_transform_3D -> ska_rigid_fuse_f64) with the reference's
signature, return structure and errors.  The module-level constants keep their names.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple, Union

import numpy as np

from .. import fusion
from ..fusion import L_HIP, L_SHO, NECK, R_HIP, R_SHO, TORSO_IDX  # noqa: F401  (bundle_adjustment/fuse/fuse.py:27-31)
from . import _common

__all__ = ["NECK", "L_HIP", "R_HIP", "L_SHO", "R_SHO", "TORSO_IDX", "rigid_transform_3D"]


def rigid_transform_3D(
    target: np.ndarray,
    source: np.ndarray,
    tau: float = 0.08,
    allow_scale: bool = False,
    wL: Optional[Union[np.ndarray, float]] = None,
    wR: Optional[Union[np.ndarray, float]] = None,
    return_diagnostics: bool = True,
    verbose: bool = False,
) -> Tuple[np.ndarray, Optional[Dict]]:
    """bundle_adjustment/fuse/fuse.py:96-232: (J,3) or (T,J,3) left / right sam-3d-body joints -> (fused, diag)."""
    return fusion.rigid_transform_3D(target, source, tau=tau, allow_scale=allow_scale, wL=wL, wR=wR,
                                     return_diagnostics=return_diagnostics, verbose=verbose, device=_common.device())
