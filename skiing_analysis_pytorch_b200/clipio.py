"""Clip-level loaders / writers for the data formats either side of the hot path.

The reference's pipelines read whole clips from disk and then walk them frame by frame; the batch API of this package
wants the clip as ONE array.  These functions read the reference's on-disk schemas straight into clip arrays (and write
the reference's output files from clip arrays), so a pipeline can go  file -> pinned host buffer -> GPU -> file  without
the per-frame Python objects in between.  Pure host code: no arithmetic of the path happens here.

  load_keypoints_pt        `.pt` dict of prepare_dataset/process/preprocess.py:157-181 (`YOLO` / `detectron2` blocks:
                           keypoints (T,K,>=2), keypoints_score (T,K)) with the semantics of
                           triangulation/load.py:74-190 (`_to_numpy_xy`, `_maybe_denorm_xy`, `_extract_scores`)
  stereo_clip_to_device    two such clips -> view-major (2,T,K,2) / (2,T,K) CUDA tensors through pinned memory:
                           the inputs of api.triangulate_reproject (what triangulation/main.py:99-109 feeds the frame loop)
  load_sam3d_sequence      SAM-3D-Body outputs, one npz (`outputs` / `arr_0` object array) or a directory of
                           `frame_*_sam_3d_body_outputs.npz` (fuse/load/load_raw.py:28-97) -> (T,J,2), (T,J,3) float64
  load_sam3d_pair          fuse/load/load_raw.py:106-147 as arrays: Xl, Xr, Ul, Ur truncated to the shorter view: the
                           inputs of fusion.fuse_clip
  save_sequence_npy        fuse/save.py:52-68 (`save_smoothed_results`) from a (T,J,3) array
  save_camera_npz / load_camera_npz / mean_intrinsics
                           vggt/save.py:84-110 (`camera_intrinsics`, `R`, `t`, `C`, stacked over frames) and the
                           per-camera mean K the BA call site forms (vggt/multi_view_process.py:543)
  save_3d_joints_clip      triangulation/save.py:31-76 (`save_3d_joints`: frame_%04d_3dpose.{npy,csv,json}) for every
                           frame of a clip
"""
from __future__ import annotations

import glob
import json
import os
from pathlib import Path
from typing import Optional

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------ .pt keypoints
def _frame_shape(data: dict):
    """(H, W) as triangulation/load.py:23-75 finds them, without decoding any video."""
    fr = data.get("frames")
    if isinstance(fr, torch.Tensor) and fr.dim() == 4:
        return int(fr.shape[1]), int(fr.shape[2])
    shp = data.get("img_shape")
    if isinstance(shp, (tuple, list)):
        try:
            return int(shp[0]), int(shp[1])
        except Exception:
            pass
    return None, None


def load_keypoints_pt(file_path, source: str = "detectron2", assume_normalized: Optional[bool] = None):
    """-> (keypoints_xy (T,K,2), keypoints_score (T,K)) numpy, like load_keypoints_from_yolo_pt /
    load_kpt_and_bbox_from_d2_pt (triangulation/load.py:147-190, 193-245): xy = keypoints[..., :2]; normalised
    coordinates (max <= 1.5, or assume_normalized=True) are scaled to pixels when the frame size is known; scores come
    from `keypoints_score`, else keypoints[..., 2], else ones."""
    data = file_path if isinstance(file_path, dict) else torch.load(file_path, map_location="cpu")
    if source not in data or "keypoints" not in data[source]:
        raise KeyError(f"pt file missing {source}.keypoints: {'<dict>' if isinstance(file_path, dict) else file_path}")
    blk = data[source]
    k = blk["keypoints"]
    if not isinstance(k, torch.Tensor):
        k = torch.as_tensor(k)
    if k.dim() < 3 or k.shape[-1] < 2:
        raise ValueError(f"Invalid keypoints shape: {tuple(k.shape)} (expect (T,K,>=2))")
    xy = k[..., :2].cpu().numpy()
    H, W = _frame_shape(data)
    if H is not None and W is not None:
        norm = (float(np.nanmax(xy)) if xy.size else 0.0) <= 1.5 if assume_normalized is None else assume_normalized
        if norm:
            xy = xy.copy()
            xy[..., 0] *= W
            xy[..., 1] *= H
    ks = blk.get("keypoints_score")
    if isinstance(ks, torch.Tensor):
        sc = ks.cpu().numpy()
    elif isinstance(ks, np.ndarray):
        sc = ks
    elif k.shape[-1] >= 3:
        sc = k[..., 2].cpu().numpy()
    else:
        sc = np.ones(xy.shape[:2], dtype=np.float32)
    if xy.ndim != 3 or xy.shape[2] != 2:
        raise ValueError(f"Invalid {source} keypoints shape after processing: {xy.shape}")
    if sc.size and sc.shape != xy.shape[:2]:
        raise ValueError(f"{source} keypoints_score shape {sc.shape} mismatches keypoints {xy.shape}")
    if not sc.size:
        sc = np.ones(xy.shape[:2], dtype=np.float32)
    return xy, sc


def stereo_clip_to_device(left, right, device="cuda"):
    """left / right = (xy (T,K,2), score (T,K)) of the two views -> (kpts (2,T,K,2) f32, conf (2,T,K) f32) on `device`,
    staged through pinned host memory (one async copy each).  Clips are truncated to the shorter view like the frame
    loop of triangulation/triangulate.py:76-78 (`zip`)."""
    T = min(len(left[0]), len(right[0]))
    K = left[0].shape[1]
    if right[0].shape[1] != K:
        raise ValueError(f"left and right have different joint counts: {K} vs {right[0].shape[1]}")
    hk = torch.empty((2, T, K, 2), dtype=torch.float32).pin_memory() if torch.cuda.is_available() else torch.empty((2, T, K, 2))
    hc = torch.empty((2, T, K), dtype=torch.float32).pin_memory() if torch.cuda.is_available() else torch.empty((2, T, K))
    for v, (xy, sc) in enumerate((left, right)):
        hk[v].copy_(torch.from_numpy(np.ascontiguousarray(xy[:T], dtype=np.float32)))
        hc[v].copy_(torch.from_numpy(np.ascontiguousarray(sc[:T], dtype=np.float32)))
    return hk.to(device, non_blocking=True), hc.to(device, non_blocking=True)


# ------------------------------------------------------------------------------------------------ SAM-3D-Body npz
def _sam_frames(path):
    """fuse/load/load_raw.py:28-97: a single npz holding the list of per-frame dicts, or per-frame npz files."""
    path = Path(path)
    if path.is_file() and path.suffix == ".npz":
        data = np.load(str(path), allow_pickle=True)
        if "arr_0" in data:
            out = data["arr_0"]
        elif "outputs" in data:
            out = data["outputs"]
        else:
            raise KeyError(f"Cannot find 'arr_0' or 'outputs' in {path}")
        if isinstance(out, np.ndarray) and out.ndim == 0:
            return [out.item()]
        return list(out)
    pattern = str((path if path.is_dir() else path.parent) / "frame_*_sam_3d_body_outputs.npz")
    files = sorted(glob.glob(pattern))
    if not files:
        raise FileNotFoundError(f"Cannot load SAM data from {path}")
    frames = []
    for f in files:
        data = np.load(f, allow_pickle=True)
        for key in ("outputs", "arr_0"):
            if key in data:
                fo = data[key]
                frames.append(fo[0] if isinstance(fo, (list, np.ndarray)) and len(fo) > 0 else fo)
                break
    return frames


def load_sam3d_sequence(path):
    """-> (p2d (T,J,2), p3d (T,J,3)) float64: `pred_keypoints_2d` / `pred_keypoints_3d` of every frame stacked."""
    frames = _sam_frames(path)
    if not frames:
        return np.zeros((0, 0, 2)), np.zeros((0, 0, 3))
    p2 = np.stack([np.asarray(f["pred_keypoints_2d"], dtype=np.float64) for f in frames])
    p3 = np.stack([np.asarray(f["pred_keypoints_3d"], dtype=np.float64) for f in frames])
    return p2, p3


def load_sam3d_pair(paths: dict):
    """fuse/load/load_raw.py:106-147 as arrays: {'sam_l': ..., 'sam_r': ...} -> Xl, Xr (T,J,3), Ul, Ur (T,J,2) with
    T = min(len(left), len(right)) - the arguments of fusion.fuse_clip."""
    Ul, Xl = load_sam3d_sequence(paths["sam_l"])
    Ur, Xr = load_sam3d_sequence(paths["sam_r"])
    T = min(len(Xl), len(Xr))
    return Xl[:T], Xr[:T], Ul[:T], Ur[:T]


def save_sequence_npy(seq, out_path) -> str:
    """fuse/save.py:52-68 (`save_smoothed_results`) from a (T,J,3) array or CUDA tensor: float64 .npy, NaN = missing."""
    arr = seq.detach().cpu().numpy() if isinstance(seq, torch.Tensor) else np.asarray(seq)
    if arr.shape[0] == 0:
        raise ValueError("smooth_seq is empty; nothing to save")
    arr = np.array(arr, dtype=np.float64)
    arr[~np.isfinite(arr).all(-1)] = np.nan  # a row with a non-finite entry is a missing joint (fuse/fuse.py:76-82)
    out_path = str(out_path)
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    np.save(out_path, arr)
    return out_path


# ------------------------------------------------------------------------------------------------ VGGT cameras
def save_camera_npz(out_path, camera_intrinsics, R, t, C) -> Path:
    """vggt/save.py:84-110: per-frame lists (or stacked arrays) -> `<out_path>.npz` with keys camera_intrinsics (N,C,3,3),
    R (N,C,3,3), t (N,C,3), C (N,C,3)."""
    p = Path(out_path).with_suffix(".npz")
    np.savez_compressed(p, camera_intrinsics=np.stack(list(camera_intrinsics), axis=0), R=np.stack(list(R), axis=0),
                        t=np.stack(list(t), axis=0), C=np.stack(list(C), axis=0))
    return p


def load_camera_npz(path) -> dict:
    d = np.load(Path(path), allow_pickle=False)
    out = {k: d[k] for k in ("camera_intrinsics", "R", "t", "C")}
    N, Cn = out["R"].shape[:2]
    if out["camera_intrinsics"].shape != (N, Cn, 3, 3) or out["t"].shape != (N, Cn, 3) or out["C"].shape != (N, Cn, 3):
        raise ValueError(f"inconsistent camera npz shapes: { {k: v.shape for k, v in out.items()} }")
    return out


def mean_intrinsics(camera_intrinsics) -> np.ndarray:
    """(N,C,3,3) -> (C,3,3): the per-camera mean K the BA call site passes (vggt/multi_view_process.py:543)."""
    return np.mean(np.asarray(camera_intrinsics), axis=0)


# ------------------------------------------------------------------------------------------------ 3D joints per frame
def save_3d_joints_clip(joints_3d, save_dir, R, T, video_path: dict, fmt: str = "npy", first_frame: int = 0):
    """triangulation/save.py:31-76 for every frame of a clip: joints_3d (T,J,3) array / CUDA tensor; R, T either one
    pose for the clip or per-frame sequences; writes frame_%04d_3dpose.<fmt> exactly as `save_3d_joints` does.
    Returns the list of written paths."""
    X = joints_3d.detach().cpu().numpy() if isinstance(joints_3d, torch.Tensor) else np.asarray(joints_3d)
    if fmt not in ("npy", "csv", "json"):
        raise ValueError(f"Unsupported format: {fmt}")
    os.makedirs(save_dir, exist_ok=True)
    R, T = np.asarray(R), np.asarray(T)
    per_frame_R, per_frame_T = R.ndim == 3, T.ndim >= 2 and T.shape[0] == X.shape[0] and T.size != 3
    out = []
    for i in range(X.shape[0]):
        idx = first_frame + i
        base = os.path.join(save_dir, f"frame_{idx:04d}_3dpose")
        if fmt == "npy":
            np.save(base + ".npy", X[i])
        elif fmt == "csv":
            import pandas as pd

            df = pd.DataFrame(X[i], columns=["X", "Y", "Z"])
            df.index.name = "joint_id"
            df.to_csv(base + ".csv", float_format="%.6f")
        else:
            r = R[i] if per_frame_R else R
            t = T[i] if per_frame_T else T
            data = {"frame": idx, "num_joints": len(X[i]), "joints_3d": X[i].tolist(), "R": r.tolist(), "T": t.tolist(),
                    "video_path": {k: str(v) for k, v in video_path.items()}}
            with open(base + ".json", "w", encoding="utf-8") as f:
                json.dump(data, f, ensure_ascii=False, indent=2)
        out.append(base + "." + fmt)
    return out
