"""Two-view 3D-3D fusion of monocular pose estimates + adaptive EMA smoothing for a whole clip.

Host side only: argument checking, dict <-> array conversion and the choice of the EMA chunking; all arithmetic runs in
libska.so (ska_fuse_frames_f64, ska_ema_f64 - csrc/ska_fuse.cu).  There is no CPU path.

Mirrors the reference's `fuse` pipeline:
  fuse/main_raw.py:199-250  per frame: _align_right_to_left -> weakpersp_reproj_confidence (x2) ->
                            crossview_consistency_confidence -> q = sqrt(conf1 * conf2) -> fuse_frame_3d
  fuse/fuse.py:329-412      temporal_smooth_ema over the fused sequence
`fuse_clip` / `temporal_smooth_ema` take (T,J,.) float64 CUDA tensors (NaN rows = missing joints);
`fuse_person` / `temporal_smooth_ema_dicts` take the reference's dict-per-frame structures and give back what
main_raw.py builds (lists of {joint id: xyz}), so the main's frame loop collapses into two launches.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import NamedTuple, Optional

import numpy as np
import torch

from . import _cabi, _lib

# fuse/main_raw.py:18-22
IDX_PELVIS, IDX_LHIP, IDX_RHIP, IDX_LSHO, IDX_RSHO = 14, 11, 12, 5, 6
# fuse/fuse.py:364-366
CORE_IDS = frozenset({1, 2, 69})
LIMB_IDS = frozenset({5, 6, 7, 8, 9, 10, 11, 12})
ENDPOINT_IDS = frozenset({13, 14, 41, 62})


class FusedClip(NamedTuple):
    fused: torch.Tensor                 # (T,J,3) float64, NaN rows where neither view has the joint
    q_l: Optional[torch.Tensor]         # (T,J) left-view quality sqrt(conf_weakpersp * conf_crossview)
    q_r: Optional[torch.Tensor]
    aligned: Optional[torch.Tensor]     # (T,J,3) right view in the left view's frame
    status: torch.Tensor                # (T,) uint8 bit set (_cabi.FUSE_*)


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _f64_cuda(x, name, last):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: this package has no CPU path")
    if x.dim() != 3 or x.shape[-1] != last:
        raise ValueError(f"{name} must be (T,J,{last}), got {tuple(x.shape)}")
    return x.to(torch.float64).contiguous()


def fuse_clip(Xl, Xr, Ul, Ur, *, sigma_px: float = 12.0, sigma_3d: float = 0.08, scale_mode: str = "hip", min_points: int = 8,
              key_joints=(IDX_PELVIS, IDX_LHIP, IDX_RHIP, IDX_LSHO, IDX_RSHO), want=("q", "aligned"), strict: bool = True,
              force_jacobi: bool = False, align: bool = True, per_frame_kernel: bool = False) -> FusedClip:
    """Fuse two per-view 3D skeleton sequences.  Xl, Xr (T,J,3) in each view's own frame, Ul, Ur (T,J,2) pixels.
    strict=True raises ValueError if the weak-perspective fit of any frame is impossible (fewer than `min_points`
    joints with finite 3D and 2D, or degenerate 3D), as the reference does (fuse/confidence.py:31-32, 52-53) - this
    reads the status back (one synchronisation); strict=False leaves those frames NaN and reports them in `status`.
    force_jacobi=True runs the one-sided Jacobi SVD for every frame's rigid alignment (the kernel otherwise takes a Newton
    polar-decomposition fast path when the cross-covariance is well conditioned and not reflected) - a test hook.
    per_frame_kernel=True runs the single warp-per-frame kernel instead of the three-stage path (A/B testing).
    align=False skips the rigid alignment: the two views already share a coordinate system, as in the reference's Unity
    pipeline (fuse/main_unity.py:96-132, 15 target joints whose key indices are positions in that array)."""
    Xl = _f64_cuda(Xl, "Xl", 3)
    T, J, _ = Xl.shape
    Xr, Ul, Ur = _f64_cuda(Xr, "Xr", 3), _f64_cuda(Ul, "Ul", 2), _f64_cuda(Ur, "Ur", 2)
    for n, a, last in (("Xr", Xr, 3), ("Ul", Ul, 2), ("Ur", Ur, 2)):
        if tuple(a.shape) != (T, J, last):
            raise ValueError(f"{n} must be ({T},{J},{last}), got {tuple(a.shape)}")
        if a.device != Xl.device:
            raise ValueError("all inputs must live on the same device")
    if J > _cabi.FUSE_MAX_JOINTS:
        raise ValueError(f"at most {_cabi.FUSE_MAX_JOINTS} joints, got {J}")
    if scale_mode not in ("hip", "torso"):
        raise ValueError("scale_mode must be 'hip' or 'torso'")  # fuse/confidence.py:175
    dev = Xl.device
    prm = _cabi.SkaFuseParams(sigma_px=float(sigma_px), sigma_3d=float(sigma_3d), scale_mode=0 if scale_mode == "hip" else 1,
                              min_points=int(min_points), root=int(key_joints[0]), lhip=int(key_joints[1]), rhip=int(key_joints[2]),
                              lsho=int(key_joints[3]), rsho=int(key_joints[4]), pad_=(1 if force_jacobi else 0) | (0 if align else 2) | (4 if per_frame_kernel else 0))
    f64 = dict(dtype=torch.float64, device=dev)
    fused = torch.empty((T, J, 3), **f64)
    ql = torch.empty((T, J), **f64) if "q" in want else None
    qr = torch.empty((T, J), **f64) if "q" in want else None
    al = torch.empty((T, J, 3), **f64) if "aligned" in want else None
    status = torch.zeros((T,), dtype=torch.uint8, device=dev)
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    lib = _lib.load()
    ws_bytes = 0 if per_frame_kernel else int(lib.ska_fuse_workspace_bytes(T))
    ws = torch.empty(max(ws_bytes, 8) // 8, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ska_fuse_frames_f64(p(Xl), p(Xr), p(Ul), p(Ur), T, J, C.byref(prm), p(fused), p(ql), p(qr), p(al), p(status),
                                           p(ws), ws_bytes, _stream(dev)))
    if strict and T:
        bad = (status & (_cabi.FUSE_FIT_LEFT_FAILED | _cabi.FUSE_FIT_RIGHT_FAILED)) != 0
        if bool(bad.any()):
            t = int(torch.nonzero(bad)[0])
            raise ValueError(f"Not enough valid points to fit (or degenerate 3D points) in frame {t}")
    return FusedClip(fused, ql, qr, al, status)


def alpha_per_joint(target_ids, alpha: float = 0.7, adaptive: bool = True, alpha_min: float = 0.45, alpha_max: float = 0.92) -> np.ndarray:
    """Per-joint base alpha of fuse/fuse.py:362-376 (core joints smoother, extremities more responsive)."""
    a = np.full(len(target_ids), float(alpha))
    if adaptive:
        for j, jid in enumerate(target_ids):
            if jid in CORE_IDS:
                a[j] = alpha * 0.85
            elif jid in LIMB_IDS:
                a[j] = alpha * 1.00
            elif jid in ENDPOINT_IDS:
                a[j] = alpha * 1.15
        a = np.clip(a, alpha_min, alpha_max)
    return a


def ema_halo(alpha: float, adaptive: bool, alpha_min: float, alpha_max: float, tol: float = 1e-18):
    """Finite samples a chunk must replay so that the truncated recurrence equals the sequential scan to below fp64
    rounding; None if the recurrence does not contract fast enough (run sequentially).  One step maps a state error d
    to ((1 - a) I - g |x - y| dd^T) d with a = clip(alpha_j + g |x - y|): the factors are 1 - a and 1 - alpha_j - 2 g |x - y|,
    both inside [1 + alpha_min - 2 alpha_max, 1 - alpha_min] when adaptive, and exactly 1 - alpha otherwise."""
    rho = max(abs(1.0 - alpha_min), abs(1.0 + alpha_min - 2.0 * alpha_max)) if adaptive else abs(1.0 - alpha)
    if not (rho < 0.9):
        return None
    if rho <= 0.0:
        return 1
    return int(math.ceil(math.log(tol) / math.log(rho)))


def temporal_smooth_ema(X, target_ids=None, alpha: float = 0.7, adaptive: bool = True, alpha_min: float = 0.45, alpha_max: float = 0.92,
                        speed_gain: float = 0.25, *, chunk: int = 512, exact: bool = False) -> torch.Tensor:
    """fuse/fuse.py:329-412 on a (T,J,3) CUDA tensor (NaN rows = missing).  exact=True runs the sequential scan
    (bit-identical to the reference's numpy loop); otherwise frames are processed in parallel chunks with a replay halo
    chosen by `ema_halo` (equal to the sequential result to below fp64 rounding)."""
    X = _f64_cuda(X, "X", 3)
    T, J, _ = X.shape
    ids = list(range(J)) if target_ids is None else list(target_ids)
    if len(ids) != J:
        raise ValueError(f"target_ids has {len(ids)} entries for {J} joints")
    dev = X.device
    aj = torch.from_numpy(alpha_per_joint(ids, alpha, adaptive, alpha_min, alpha_max)).to(dev)
    halo = None if exact else ema_halo(alpha, adaptive, alpha_min, alpha_max)
    Y = torch.empty_like(X)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.ska_ema_f64(C.c_void_p(X.data_ptr()), T, J, C.c_void_p(aj.data_ptr()), 1 if adaptive else 0, float(alpha),
                                   float(alpha_min), float(alpha_max), float(speed_gain), int(chunk), -1 if halo is None else int(halo),
                                   C.c_void_p(Y.data_ptr()), _stream(dev)))
    return Y


# ---------------------------------------------------------------------------------------------------------------
# dict-per-frame structures of the reference (fuse/fuse.py:67-82, fuse/load/load_raw.py)
def dicts_to_array(seq, joint_ids, dim: int) -> np.ndarray:
    """[{joint id: values}, ...] -> (T,J,dim) float64 with NaN rows for missing joints (fuse/fuse.py:67-74)."""
    out = np.full((len(seq), len(joint_ids), dim), np.nan)
    for t, d in enumerate(seq):
        for k, jid in enumerate(joint_ids):
            if jid in d:
                out[t, k] = np.asarray(d[jid], np.float64)
    return out


def array_to_dicts(arr: np.ndarray, joint_ids):
    """(T,J,dim) -> [{joint id: row}] keeping finite rows only (fuse/fuse.py:76-82)."""
    ok = np.isfinite(arr).all(-1)
    return [{jid: arr[t, k].copy() for k, jid in enumerate(joint_ids) if ok[t, k]} for t in range(arr.shape[0])]


def temporal_smooth_ema_dicts(fused_seq_dicts, target_ids, alpha: float = 0.7, adaptive: bool = True, alpha_min: float = 0.45,
                              alpha_max: float = 0.92, speed_gain: float = 0.25, device="cuda"):
    """Signature of the reference's temporal_smooth_ema (fuse/fuse.py:329-337): list of dicts in, list of dicts out."""
    if len(fused_seq_dicts) == 0:
        return []
    X = torch.from_numpy(dicts_to_array(fused_seq_dicts, target_ids, 3)).to(device)
    Y = temporal_smooth_ema(X, target_ids, alpha, adaptive, alpha_min, alpha_max, speed_gain)
    return array_to_dicts(Y.cpu().numpy(), target_ids)


def fuse_person(all_frame_results, *, sigma_px: float = 12.0, sigma_3d: float = 0.08, alpha: float = 0.7, adaptive_smooth: bool = True,
                smooth_alpha_min: float = 0.45, smooth_alpha_max: float = 0.92, smooth_speed_gain: float = 0.25, device="cuda"):
    """The body of fuse/main_raw.py's person loop (:185-250) on the structure `load_raw` returns
    ({frame: {"L_2D": {"pred": {id: uv}}, "L_3D": {...}, "R_2D": {...}, "R_3D": {...}}}): returns
    (fused_seq, smooth_seq, joint_ids) - the two lists of {joint id: xyz} the main saves."""
    frames = list(all_frame_results.values())
    if not frames:
        return [], [], []
    ids = list(range(len(frames[0]["L_3D"]["pred"])))  # main_raw.py:186-189
    get = lambda key, dim: torch.from_numpy(dicts_to_array([f[key]["pred"] for f in frames], ids, dim)).to(device)
    res = fuse_clip(get("L_3D", 3), get("R_3D", 3), get("L_2D", 2), get("R_2D", 2), sigma_px=sigma_px, sigma_3d=sigma_3d, want=())
    smooth = temporal_smooth_ema(res.fused, ids, alpha, adaptive_smooth, smooth_alpha_min, smooth_alpha_max, smooth_speed_gain)
    return array_to_dicts(res.fused.cpu().numpy(), ids), array_to_dicts(smooth.cpu().numpy(), ids), ids
