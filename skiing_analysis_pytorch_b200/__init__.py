"""B200-native batched multi-view 3D keypoint reconstruction (DLT triangulation + reprojection
scoring + LM bundle adjustment) behind the call signatures of the reference's hot-path modules.

Compute lives in lib/libska.so (hand-written sm_100a CUDA, C ABI in include/ska.h); this package
is the thin Python/PyTorch host layer.  There is no CPU fallback anywhere in the package.
"""
from .api import TriangulationResult, triangulate_reproject, triangulate_reproject_host  # noqa: F401

__all__ = ["triangulate_reproject", "triangulate_reproject_host", "TriangulationResult"]
