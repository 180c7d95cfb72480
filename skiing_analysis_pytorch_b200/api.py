"""Batched GPU API: whole clips in one launch.  PyTorch is only the allocator / stream provider;
all arithmetic happens in libska.so (hand-written sm_100a CUDA) through the C ABI of include/ska.h.

The drop-in shims in ``dropin/`` call these with the reference's signatures; pipelines that hold a
whole clip should call them directly.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _cabi, _lib


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} must be a CUDA tensor: this package has no CPU path (got device {t.device})"
        )


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


@dataclass
class TriangulationResult:
    X: torch.Tensor  # (T,J,3) f32
    err: Optional[torch.Tensor]  # (V,T,J) or (T,V,J) f32 pixel error per view
    proj: Optional[torch.Tensor]  # same layout as kpts
    status: Optional[torch.Tensor]  # (T,J) uint8: 0 fast path, 1 Jacobi fallback, 2 non-finite


def triangulate_reproject(
    kpts: torch.Tensor,
    K,
    R,
    t,
    conf: Optional[torch.Tensor] = None,
    dist=None,
    *,
    layout: str = "VTJ2",
    solver: str = "secular",
    weight_power: float = 1.0,
    pinhole_reproj: bool = False,
    centre: Optional[Sequence[float]] = None,
    want: Sequence[str] = ("X", "err"),
    out: Optional[dict] = None,
) -> TriangulationResult:
    """Confidence-weighted V-view DLT triangulation fused with reprojection scoring.

    kpts  (V,T,J,2) [layout "VTJ2"] or (T,V,J,2) [layout "TVJ2"], float32 CUDA, pixels.
    conf  matching (V,T,J) / (T,V,J) or None (unit weights = the reference's behaviour,
          triangulation/triangulate.py:65-67).
    K     (3,3) or (V,3,3); R (V,3,3), t (V,3): world->camera, host arrays (fp64 kept).
    dist  None | (n,) OpenCV coefficients shared by all views | list per view; used for SCORING only
          (the reference triangulates raw pixels and reprojects with distortion - quirk Q1,
          triangulate.py:82 vs :103-105).
    Returns X (T,J,3) f32, err = |proj - kpt| per view, optionally proj and status.
    V=2, conf=None reproduces triangulate_joints / triangulate_point of the reference to fp32
    rounding (tests/test_triangulate_gpu.py).
    """
    _require_cuda(kpts, "kpts")
    if kpts.dtype != torch.float32:
        raise TypeError(f"kpts must be float32, got {kpts.dtype}")
    if kpts.dim() != 4 or kpts.shape[-1] != 2:
        raise ValueError(f"kpts must be 4-D with last dim 2, got {tuple(kpts.shape)}")
    if layout == "VTJ2":
        V, T, J, _ = kpts.shape
        lay = _cabi.LAYOUT_VIEW_MAJOR
        eshape = (V, T, J)
    elif layout == "TVJ2":
        T, V, J, _ = kpts.shape
        lay = _cabi.LAYOUT_FRAME_MAJOR
        eshape = (T, V, J)
    else:
        raise ValueError(f"layout must be 'VTJ2' or 'TVJ2', got {layout!r}")
    kpts = kpts.contiguous()
    if conf is not None:
        _require_cuda(conf, "conf")
        if tuple(conf.shape) != eshape:
            raise ValueError(f"conf shape {tuple(conf.shape)} does not match {eshape}")
        conf = conf.to(torch.float32).contiguous()
    if solver not in _cabi.SOLVERS:
        raise ValueError(f"solver must be one of {sorted(_cabi.SOLVERS)}")
    flags = _cabi.SOLVERS[solver]
    if weight_power == 0.5:
        flags |= _cabi.WEIGHT_SQRT
    elif weight_power != 1.0:
        raise ValueError("weight_power must be 1.0 (row weight = conf) or 0.5 (row weight = sqrt(conf))")
    if pinhole_reproj:
        flags |= _cabi.PINHOLE_REPROJ
    R = np.asarray(R, np.float64)
    if R.shape != (V, 3, 3):
        raise ValueError(f"R must be ({V},3,3), got {R.shape}")
    cams = _cabi.make_cameras(K, R, t, dist)
    dev = kpts.device
    out = out or {}
    X = out.get("X")
    if X is None:
        X = torch.empty((T, J, 3), dtype=torch.float32, device=dev)
    err = out.get("err") if "err" in want else None
    if "err" in want and err is None:
        err = torch.empty(eshape, dtype=torch.float32, device=dev)
    proj = out.get("proj") if "proj" in want else None
    if "proj" in want and proj is None:
        proj = torch.empty_like(kpts)
    status = out.get("status") if "status" in want else None
    if "status" in want and status is None:
        status = torch.empty((T, J), dtype=torch.uint8, device=dev)
    cptr = None
    if centre is not None:
        carr = (C.c_double * 3)(*[float(x) for x in centre])
        cptr = C.cast(carr, C.c_void_p)
    lib = _lib.load()
    with torch.cuda.device(dev):
        rc = lib.ska_triangulate_reproject_f32(
            cams, V, cptr, None, _ptr(kpts), _ptr(conf), T, J, lay, flags, _ptr(X), _ptr(err), _ptr(proj),
            _ptr(status), _stream_ptr(dev),
        )
    _lib.check(rc)
    return TriangulationResult(X=X, err=err, proj=proj, status=status)
