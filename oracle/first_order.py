"""First-order (Adam) bundle adjustment over the reference's FULL configured objective - the oracle of the
regularised `run_local_ba` variant (first-order form).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

What the reference configures for its (undefined) optimiser - configs/vggt.yaml:43-52: `lr`, `num_iters` and one weight
per term of bundle_adjustment/loss.py - is a first-order minimisation of
    total = w_reproj * reprojection_loss(X, R, t, K, x2d, conf)       loss.py:90-94
          + w_smooth * camera_smooth_loss(R, t)                        loss.py:103-106   (per-frame cameras)
          + w_baseline * baseline_reg_loss(R, t)                       loss.py:109-114
          + w_bone_length * bone_length_loss(X)                        loss.py:134-150
          + w_pose_temporal * pose_temporal_loss(X)                    loss.py:153-155
over X (mode "pose_only"), X and t ("pose_cam_t") or X, t and R ("full"; vggt/multi_view_process.py:338) with per-frame
cameras R (T,C,3,3), t (T,C,3) as the call site passes them (vggt/multi_view_process.py:546-564).
PARITY: the loss VALUES are the reference's own functions (pass the imported module as `L`; golden G9 is generated that
way); the optimiser is "parity unpinned" - run_local_ba is an undefined symbol - so this file is its specification:
  * Adam exactly as torch.optim.Adam computes it (m <- m + (1 - b1)(g - m); v <- b2 v + (1 - b2) g^2;
    p <- p - (lr / (1 - b1^k)) m / (sqrt(v) / sqrt(1 - b2^k) + eps)), betas (0.9, 0.999), eps 1e-8;
  * rotations move on SO(3): the gradient w.r.t. the left tangent d_omega of R = exp([d_omega]x) R_cur is
    g = (B21 - B12, B02 - B20, B10 - B01), B = (dL/dR) R_cur^T; Adam runs on g (moments kept in the moving tangent
    frame) and the step s retracts R <- exp([-s]x) R_cur.
"""
from __future__ import annotations

import torch

from . import geometry as G

MODES = ("pose_only", "pose_cam_t", "full")
DEFAULT_WEIGHTS = dict(reproj=1.0, smooth=0.1, baseline=0.01, bone_length=0.1, pose_temporal=0.1)  # configs/vggt.yaml:46-50
TERMS = ("reproj", "smooth", "baseline", "bone_length", "pose_temporal")


def tangent_grad(gR: torch.Tensor, R: torch.Tensor) -> torch.Tensor:
    B = gR @ R.transpose(-1, -2)
    return torch.stack([B[..., 2, 1] - B[..., 1, 2], B[..., 0, 2] - B[..., 2, 0], B[..., 1, 0] - B[..., 0, 1]], -1)


def so3_exp(w: torch.Tensor) -> torch.Tensor:
    """(...,3) -> (...,3,3), Rodrigues (series below 1e-8 like oracle/geometry.so3_exp)."""
    out = torch.empty(w.shape[:-1] + (3, 3), dtype=w.dtype)
    flat = w.reshape(-1, 3)
    o = out.reshape(-1, 3, 3)
    for i in range(flat.shape[0]):
        o[i] = torch.from_numpy(G.so3_exp(flat[i].numpy()))
    return out


class Adam:
    def __init__(self, lr, betas=(0.9, 0.999), eps=1e-8):
        self.lr, self.b1, self.b2, self.eps, self.k, self.state = lr, betas[0], betas[1], eps, 0, {}

    def begin(self):
        self.k += 1

    def step_of(self, name, g):
        m, v = self.state.setdefault(name, (torch.zeros_like(g), torch.zeros_like(g)))
        m += (1.0 - self.b1) * (g - m)
        v.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
        bc1, bc2 = 1.0 - self.b1**self.k, 1.0 - self.b2**self.k
        return (self.lr / bc1) * m / (v.sqrt() / bc2**0.5 + self.eps)


def total_loss(L, X, R, t, K, x2d, conf, weights):
    terms = {}
    w = weights
    if w["reproj"]:
        terms["reproj"] = L.reprojection_loss(X, R, t, K, x2d, conf, w=w["reproj"])
    if w["smooth"]:
        terms["smooth"] = L.camera_smooth_loss(R, t, w=w["smooth"])
    if w["baseline"]:
        terms["baseline"] = L.baseline_reg_loss(R, t, w=w["baseline"])
    if w["bone_length"]:
        terms["bone_length"] = L.bone_length_loss(X, None, w=w["bone_length"])
    if w["pose_temporal"]:
        terms["pose_temporal"] = L.pose_temporal_loss(X, w=w["pose_temporal"])
    return sum(terms.values()), terms


def run_adam(L, K, R0, t0, X0, x2d, conf, num_iters=50, lr=1e-2, mode="pose_only", weights=None):
    """L: module / object with bundle_adjustment/loss.py's functions.  All tensors CPU; computed in float64.
    Returns R (T,C,3,3), t (T,C,3), X (T,J,3), history (list of dict: iter, loss, one entry per term)."""
    if mode not in MODES:
        raise ValueError(f"unknown mode {mode!r}")
    w = dict(DEFAULT_WEIGHTS, **(weights or {}))
    f64 = lambda a: torch.as_tensor(a).to(torch.float64).clone()
    K, R, t, X, x2d, conf = f64(K), f64(R0), f64(t0), f64(X0), f64(x2d), f64(conf)
    opt = Adam(lr)
    hist = []
    for it in range(num_iters):
        Xv, Rv, tv = X.clone().requires_grad_(True), R.clone().requires_grad_(mode == "full"), t.clone().requires_grad_(mode != "pose_only")
        tot, terms = total_loss(L, Xv, Rv, tv, K, x2d, conf, w)
        tot.backward()
        row = dict(iter=it, loss=float(tot.detach()))
        row.update({k: float(terms[k].detach()) if k in terms else 0.0 for k in TERMS})
        hist.append(row)
        opt.begin()
        X = X - opt.step_of("X", Xv.grad)
        if mode != "pose_only":
            t = t - opt.step_of("t", tv.grad if tv.grad is not None else torch.zeros_like(t))
        if mode == "full":
            s = opt.step_of("w", tangent_grad(Rv.grad, R))
            R = so3_exp(-s) @ R
    return R, t, X, hist
