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


# ---------------------------------------------------------------------------------------------------------------
# rigid_transform_3D: the fusion of bundle_adjustment/fuse/fuse.py (== fuse/side/fuse/fuse.py, front_side/side/fuse/fuse.py)
NECK, L_HIP, R_HIP, L_SHO, R_SHO = 69, 9, 10, 5, 6          # bundle_adjustment/fuse/fuse.py:27-29
TORSO_IDX = [NECK, L_HIP, R_HIP, L_SHO, R_SHO]             # :31


class RigidFusedClip(NamedTuple):
    fused: torch.Tensor   # (T,J,3) float64
    R: torch.Tensor       # (T,3,3) right -> left rotation
    t: torch.Tensor       # (T,3)
    s: torch.Tensor       # (T,) scale (1 unless allow_scale)
    diag: torch.Tensor    # (T,4): LR_before, Fused_vs_L, Fused_vs_R, gain
    status: torch.Tensor  # (T,) uint8: 1 = fewer than 3 usable torso joints


def rigid_fuse_clip(target, source, *, tau=0.08, allow_scale: bool = False, wL=None, wR=None, torso_idx=TORSO_IDX,
                    strict: bool = True) -> RigidFusedClip:
    """Clip-level form of rigid_transform_3D (bundle_adjustment/fuse/fuse.py:96-232): target / source (T,J,3) CUDA
    tensors (left / right view), tau scalar or (J,), wL / wR None, scalar, (J,) or (T,J).  One launch pair for the clip."""
    L = _f64_cuda(target, "target", 3)
    T, J, _ = L.shape
    R = _f64_cuda(source, "source", 3)
    if tuple(R.shape) != (T, J, 3):
        raise ValueError(f"source must be ({T},{J},3), got {tuple(R.shape)}")   # the reference asserts equal shapes (:128)
    if max(torso_idx) >= J:
        raise ValueError(f"torso index {max(torso_idx)} outside a {J}-joint skeleton")  # assertion at fuse.py:131-133
    dev = L.device
    f64 = dict(dtype=torch.float64, device=dev)

    def weights(w):
        if w is None:
            return None, 0
        w = torch.as_tensor(w, **f64)
        if w.dim() == 0:
            return w.expand(J).contiguous(), 0
        if w.dim() == 1:
            if w.shape[0] != J:
                raise ValueError("wL/wR shape must be (J,), (T,J) or a scalar")
            return w.contiguous(), 0
        if tuple(w.shape) != (T, J):
            raise ValueError("wL/wR shape must be (J,), (T,J) or a scalar")
        return w.contiguous(), J

    (wl, sl), (wr, sr) = weights(wL), weights(wR)
    if sl != sr:  # one per-frame, one per-joint: expand the per-joint one
        if sl == 0 and wl is not None:
            wl, sl = wl[None].expand(T, J).contiguous(), J
        if sr == 0 and wr is not None:
            wr, sr = wr[None].expand(T, J).contiguous(), J
    stride = max(sl, sr)
    tau_t = torch.as_tensor(tau, **f64)
    tau_j = tau_t.contiguous() if tau_t.dim() == 1 else None
    if tau_j is not None and tau_j.shape[0] != J:
        raise ValueError("tau must be a scalar or (J,)")
    fused = torch.empty((T, J, 3), **f64)
    Rts = torch.empty((T, 13), **f64)
    diag = torch.empty((T, 4), **f64)
    status = torch.zeros((T,), dtype=torch.uint8, device=dev)
    torso = (C.c_int32 * 5)(*[int(k) for k in torso_idx])
    p = lambda x: None if x is None else C.c_void_p(x.data_ptr())
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.ska_rigid_fuse_f64(p(L), p(R), T, J, torso, float(tau_t) if tau_j is None else 0.0, p(tau_j), 1 if allow_scale else 0,
                                          p(wl), p(wr), stride, p(fused), p(Rts), p(diag), p(status), _stream(dev)))
    if strict and T and bool((status != 0).any()):
        raise ValueError(f"frame {int(torch.nonzero(status)[0])}: at least 3 non-collinear corresponding torso points are needed")
    return RigidFusedClip(fused, Rts[:, :9].reshape(T, 3, 3), Rts[:, 9:12], Rts[:, 12], diag, status)


def rigid_transform_3D(target, source, tau=0.08, allow_scale=False, wL=None, wR=None, return_diagnostics=True, verbose=False,
                       device="cuda"):
    """Signature and return structure of the reference's rigid_transform_3D (bundle_adjustment/fuse/fuse.py:96-105):
    numpy (J,3) or (T,J,3) in, (fused like the input, diag dict or None) out - with the whole sequence in one launch."""
    Ln, Rn = np.asarray(target, dtype=float), np.asarray(source, dtype=float)
    single = Ln.ndim == 2
    if single:
        Ln, Rn = Ln[None], Rn[None]
    assert Ln.shape == Rn.shape and Ln.shape[-1] == 3
    T, J, _ = Ln.shape
    assert max(TORSO_IDX) < J, f"max(TORSO_IDX)={max(TORSO_IDX)}, J={J}"
    r = rigid_fuse_clip(torch.from_numpy(Ln).to(device), torch.from_numpy(Rn).to(device), tau=tau, allow_scale=allow_scale, wL=wL, wR=wR)
    fused = r.fused.cpu().numpy()
    out = fused[0] if single else fused
    if not return_diagnostics:
        return out, None
    d, Rh, th, sh = r.diag.cpu().numpy(), r.R.cpu().numpy(), r.t.cpu().numpy(), r.s.cpu().numpy()
    per = [{"frame": t, "LR_before": float(d[t, 0]), "Fused_vs_L": float(d[t, 1]), "Fused_vs_R": float(d[t, 2]), "gain": float(d[t, 3]),
            "s": float(sh[t]), "R": Rh[t].copy(), "t": th[t].copy()} for t in range(T)]
    gains = d[:, 3]
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mean_gain = float(np.nanmean(gains)) if T else float("nan")
    return out, {"per_frame": per, "mean_gain": mean_gain, "bad_frames": [int(t) for t in np.nonzero(gains < 0)[0]]}
