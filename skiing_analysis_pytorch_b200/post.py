"""Post-triangulation triage and temporal smoothing for whole clips: the reference's
triangulation/postprocess.py (:54-170) as three streaming GPU passes (libska.so, csrc/ska_post.cu)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _cabi, _lib

UNDISTORT0, UNDISTORT1 = 1, 2
FLAG_POS, FLAG_ERR, FLAG_CONF, FLAG_KEEP = 1, 2, 4, 8


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


@dataclass
class TriageResult:
    X_clean: torch.Tensor   # (T,J,3) f32, rejected joints NaN
    em: torch.Tensor        # (T,J) f32 mean reprojection error of the two views (px)
    flags: torch.Tensor     # (T,J) uint8: FLAG_POS | FLAG_ERR | FLAG_CONF | FLAG_KEEP
    report: torch.Tensor    # (T,5) f64: rmse_px, median_err_px, pos_depth_ratio, kept_ratio, kept_count


def post_triage(X: torch.Tensor, kpts: torch.Tensor, K1, K2, R, T, dist1=None, dist2=None,
                conf: Optional[torch.Tensor] = None, conf_thr: float = 0.3, err_thresh_px: float = 2.0) -> TriageResult:
    """X (T,J,3) and kpts (2,T,J,2) [view-major: left, right] float32 CUDA; cam 1 = K1 [I|0], cam 2 = K2 [R|T];
    dist1 / dist2: undistort that view's pixels first (None = pixels already undistorted);
    conf (2,T,J) optional.  One launch for the clip + the per-frame report."""
    from . import api

    if not (X.is_cuda and kpts.is_cuda):
        raise RuntimeError("X and kpts must be CUDA tensors: this package has no CPU path")
    X = X.to(torch.float32).contiguous()
    kpts = kpts.to(torch.float32).contiguous()
    Tn, J, _ = X.shape
    if tuple(kpts.shape) != (2, Tn, J, 2):
        raise ValueError(f"kpts must be (2,{Tn},{J},2), got {tuple(kpts.shape)}")
    if conf is not None:
        conf = conf.to(device=X.device, dtype=torch.float32).contiguous()
        if tuple(conf.shape) != (2, Tn, J):
            raise ValueError(f"conf must be (2,{Tn},{J}), got {tuple(conf.shape)}")
    Kk = np.stack([np.asarray(K1, np.float64).reshape(3, 3), np.asarray(K2, np.float64).reshape(3, 3)])
    Rr = np.stack([np.eye(3), np.asarray(R, np.float64).reshape(3, 3)])
    tt = np.stack([np.zeros(3), np.asarray(T, np.float64).reshape(3)])
    dists = [None if d is None else np.asarray(d, np.float64).reshape(-1) for d in (dist1, dist2)]
    cams = _cabi.make_cameras(Kk, Rr, tt, dists if any(d is not None for d in dists) else None)
    flags = (UNDISTORT0 if dist1 is not None else 0) | (UNDISTORT1 if dist2 is not None else 0)
    dev = X.device
    Xc = torch.empty_like(X)
    em = torch.empty((Tn, J), dtype=torch.float32, device=dev)
    fl = torch.empty((Tn, J), dtype=torch.uint8, device=dev)
    counts = torch.zeros((Tn, 4), dtype=torch.int32, device=dev)
    lib = _lib.load()
    if Tn > 0:
        with torch.cuda.device(dev):
            _lib.check(lib.ska_post_triage_f32(cams, _p(X), _p(kpts), _p(conf), Tn, J, flags, float(conf_thr), float(err_thresh_px),
                                               _p(Xc), _p(em), _p(fl), _stream(dev)))
            _lib.check(lib.ska_frame_flag_counts_u8(_p(fl), Tn, J, _p(counts), _stream(dev)))
    st = api.frame_stats(em.reshape(1, Tn, J))  # (T,1,4): rmse, mean, median, max (nan-aware)
    rep = torch.stack([st[:, 0, 0].double(), st[:, 0, 2].double(), counts[:, 0].double() / J, counts[:, 3].double() / J,
                       counts[:, 3].double()], dim=1)
    return TriageResult(X_clean=Xc, em=em, flags=fl, report=rep)


def effective_window(T: int, win: int) -> int:
    """Window actually used by smooth_skeleton (postprocess.py:58): odd, capped by the clip length."""
    return min(win if win % 2 == 1 else win + 1, max(1 if T % 2 == 1 else T - 1, 3))


def smooth_skeleton(X: torch.Tensor, win: int = 9, poly: int = 2) -> torch.Tensor:
    """Savitzky-Golay smoothing along time of every (joint, coordinate) series over its finite samples
    (postprocess.py:54-68; scipy.signal.savgol_filter mode='interp').  X (T,J,C) float32 CUDA -> same shape."""
    if not X.is_cuda:
        raise RuntimeError("X must be a CUDA tensor: this package has no CPU path")
    Xc = X.to(torch.float32).contiguous()
    Tn = Xc.shape[0]
    S = int(np.prod(Xc.shape[1:]))
    w = effective_window(Tn, int(win))
    if int(poly) >= w:
        raise ValueError("polyorder must be less than window_length.")  # what scipy.signal.savgol_filter raises
    out = torch.empty_like(Xc)
    if Tn == 0:
        return out
    lib = _lib.load()
    dev = Xc.device
    nbytes = int(lib.ska_savgol_workspace_bytes(Tn, S))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ska_savgol_f32(_p(Xc), Tn, S, w, int(poly), _p(out), _p(ws), nbytes, _stream(dev)))
    ws.record_stream(torch.cuda.current_stream(dev))
    return out
