"""fp64 Schur-complement Levenberg-Marquardt bundle adjustment - the oracle of the BA hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY STATUS: the *cost* is pinned by the reference's own ``reprojection_loss``
(bundle_adjustment/loss.py:90-94; golden G3 in tests/golden/g3_g4_loss.npz).  The LM *trajectory*
is PARITY UNPINNED: the reference calls ``run_local_ba`` (vggt/multi_view_process.py:553-564) but
defines it nowhere, so there is no reference LM to compare with.  This file IS the specification
(DESIGN.md section 5); the CUDA path implements the same algorithm in fp32
with fp64 reductions and is tested against the history produced here.

Algorithm
  cost      F = sum_tcj w_tcj |pi_c(X_tj) - x_tcj|^2,  w = conf / (sum conf + 1e-6)   (== loss.py, w=1)
  pi        X_c = R X + t;  Z = max(z, 1e-6);  (x, y) = X_c.xy / Z;  (u, v) = (K [x, y, 1])[:2]
            d/dz is zero while the clamp is active (what autograd of loss.py:67 gives)
  params    points X (T,J,3); per camera c >= 1 (camera 0 is the gauge): [d_omega(3), d_t(3)],
            update R <- exp([d_omega]x) R, t <- t + d_t; `free` masks individual camera parameters
  step      H = J^T W J, g = J^T W r; damping H + lam*diag(H) on both blocks; Schur complement onto the
            cameras; Cholesky; back-substitution of the points
  control   accept iff F_trial < F; gain ratio rho = (F - F_trial) / (delta^T (lam D delta - g));
            accept: lam <- lam * max(1/3, 1 - (2 rho - 1)^3), nu <- 2;  reject: lam <- lam * nu, nu <- 2 nu;
            lam0 = 1e-3 (Nielsen 1999)
  history   one row per trial: iter, cost (before), trial_cost, lambda (used), rho, accepted, n_clamped, pred
Multi-shard: every sum over points is a plain sum, so shards add (tests run 1/2/4/8 shards).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import geometry as G

ZMIN = 1e-6
LAMBDA0 = 1e-3


# golden set G6 (tests/golden/g6_lm_history.npz): name -> (rig, T, J, mode); BASELINE configs 3 and 5 at reduced T
G6_CASES = {"c3": ("2b", 200, 17, "full"), "c5": ("8", 24, 70, "full"), "c3_t": ("2b", 120, 17, "pose_cam_t"),
            "c4cam": ("4", 60, 17, "full")}


def make_problem(rig: str, T: int, J: int, seed: int = 0):
    """Deterministic BA test problem: GT rig perturbed by N(0,0.01) rad /
    N(0,0.05) m (camera 0 untouched), points = DLT of the noisy observations under the perturbed rig.
    Returns (clip, R0, t0, X0)."""
    from skiing_analysis_pytorch_b200 import synth

    clip = synth.make_clip(rig, T, J, seed=seed)
    R0, t0 = synth.perturb_cameras(clip.R, clip.t, seed=seed + 1)
    V = len(R0)
    P = np.stack([G.make_P(clip.K[v], R0[v], t0[v]) for v in range(V)])
    X0 = G.dlt_triangulate(P, clip.x_vm.reshape(V, -1, 2)).reshape(T, J, 3)
    return clip, R0, t0, X0


def free_mask(C: int, mode: str = "full") -> np.ndarray:
    """(C,6) bool: which of [d_omega(3), d_t(3)] are optimised. Camera 0 is always fixed.
    Modes follow the names at vggt/multi_view_process.py:338 / configs/vggt.yaml:52."""
    m = np.zeros((C, 6), bool)
    if mode == "full":
        m[1:] = True
    elif mode == "pose_cam_t":
        m[1:, 3:] = True
    elif mode != "pose_only":
        raise ValueError(f"unknown mode {mode!r}")
    return m


def residual_blocks(X, R, t, K, x2d, w):
    """Per observation: weighted squared error, point Jacobian A (2x3), camera Jacobian B (2x6).
    X (N,3) [N = T*J flattened], x2d (N,C,2), w (N,C).  Returns e (N,C,2), A (N,C,2,3), B (N,C,2,6), clamped (N,C)."""
    Xc = np.einsum("cab,nb->nca", R, X) + t[None]
    z = Xc[..., 2]
    clamped = z < ZMIN
    Z = np.maximum(z, ZMIN)
    iz = 1.0 / Z
    x = Xc[..., 0] * iz
    y = Xc[..., 1] * iz
    K = K[None]
    u = K[..., 0, 0] * x + K[..., 0, 1] * y + K[..., 0, 2]
    v = K[..., 1, 0] * x + K[..., 1, 1] * y + K[..., 1, 2]
    e = np.stack([u, v], -1) - x2d
    live = (~clamped).astype(float)
    ju = np.stack([K[..., 0, 0] * iz, K[..., 0, 1] * iz, -(K[..., 0, 0] * x + K[..., 0, 1] * y) * iz * live], -1)
    jv = np.stack([K[..., 1, 0] * iz, K[..., 1, 1] * iz, -(K[..., 1, 0] * x + K[..., 1, 1] * y) * iz * live], -1)
    Jx = np.stack([ju, jv], -2)  # (N,C,2,3) d(u,v)/dX_c
    A = np.einsum("ncij,cjk->ncik", Jx, R)
    p = Xc - t[None]  # R X
    Bw = np.cross(p[..., None, :], Jx)  # rows p x ju, p x jv
    B = np.concatenate([Bw, Jx], -1)
    del w
    return e, A, B, clamped


def cost_only(X, R, t, K, x2d, w):
    Xc = np.einsum("cab,nb->nca", R, X) + t[None]
    Z = np.maximum(Xc[..., 2], ZMIN)
    x = Xc[..., 0] / Z
    y = Xc[..., 1] / Z
    K = K[None]
    u = K[..., 0, 0] * x + K[..., 0, 1] * y + K[..., 0, 2]
    v = K[..., 1, 0] * x + K[..., 1, 1] * y + K[..., 1, 2]
    d = (u - x2d[..., 0]) ** 2 + (v - x2d[..., 1]) ** 2
    return float((w * d).sum()), int((Xc[..., 2] < ZMIN).sum())


def _safe_inverse(Hd):
    """Inverse of the damped 3x3 point blocks; a block that is not positive definite (a point nobody observes:
    all confidences 0, or a rank-deficient block) gets the ZERO matrix - no step and no Schur term for that
    point, the rule the CUDA kernels implement with their Cholesky pivot test (ska_ba.cuh chol3)."""
    ok = np.all(np.isfinite(Hd), axis=(1, 2))
    ev = np.linalg.eigvalsh(np.where(ok[:, None, None], Hd, np.eye(3)))
    ok &= ev[:, 0] > 0
    inv = np.linalg.inv(np.where(ok[:, None, None], Hd, np.eye(3)))
    return inv * ok[:, None, None]


@dataclass
class Linearisation:
    Hcc: np.ndarray  # (C,6,6) undamped
    gc: np.ndarray  # (C,6)
    Sw: np.ndarray  # (6C,6C) sum W^T Hpp_d^-1 W
    bw: np.ndarray  # (6C,)   sum W^T Hpp_d^-1 gp
    cost: float
    n_clamped: int
    Hpp: np.ndarray = field(repr=False, default=None)  # (N,3,3) undamped
    gp: np.ndarray = field(repr=False, default=None)  # (N,3)
    W: np.ndarray = field(repr=False, default=None)  # (N,3,6C)

    def __add__(self, o):
        return Linearisation(self.Hcc + o.Hcc, self.gc + o.gc, self.Sw + o.Sw, self.bw + o.bw, self.cost + o.cost,
                             self.n_clamped + o.n_clamped)


def linearise(X, R, t, K, x2d, w, lam, keep_points=False) -> Linearisation:
    N, C = w.shape
    e, A, B, clamped = residual_blocks(X, R, t, K, x2d, w)
    wA = A * w[..., None, None]
    Hpp = np.einsum("ncij,ncik->njk", wA, A)
    gp = np.einsum("ncij,nci->nj", wA, e)
    Hcc = np.einsum("ncij,ncik->cjk", B * w[..., None, None], B)
    gc = np.einsum("ncij,nci->cj", B * w[..., None, None], e)
    W = np.einsum("ncij,ncik->njck", wA, B).reshape(N, 3, 6 * C)
    Hd = Hpp + lam * np.einsum("nii,ij->nij", Hpp, np.eye(3))
    Hinv = _safe_inverse(Hd)
    HW = np.einsum("nij,njk->nik", Hinv, W)
    Sw = np.einsum("nik,nil->kl", W, HW)
    bw = np.einsum("nik,nij,nj->k", W, Hinv, gp)
    cost = float((w[..., None] * e**2).sum())
    lin = Linearisation(Hcc, gc, Sw, bw, cost, int(clamped.sum()))
    if keep_points:
        lin.Hpp, lin.gp, lin.W = Hpp, gp, W
    return lin


def solve_reduced(lin: Linearisation, lam: float, free: np.ndarray):
    """S = Hcc + lam*diag(Hcc) - Sw on the free parameters (fixed ones masked to identity), Cholesky.
    Returns delta_c (C,6), pred_cam, ok."""
    C = lin.Hcc.shape[0]
    n = 6 * C
    S = -lin.Sw.copy()
    for c in range(C):
        blk = lin.Hcc[c] + lam * np.diag(np.diag(lin.Hcc[c]))
        S[6 * c : 6 * c + 6, 6 * c : 6 * c + 6] += blk
    b = -lin.gc.reshape(n) + lin.bw
    f = free.reshape(n)
    S[~f, :] = 0.0
    S[:, ~f] = 0.0
    S[~f, ~f] = 1.0
    b = np.where(f, b, 0.0)
    try:
        L = np.linalg.cholesky(S)
    except np.linalg.LinAlgError:
        return np.zeros((C, 6)), 0.0, False
    d = np.linalg.solve(L.T, np.linalg.solve(L, b))
    Hd = np.concatenate([np.diag(lin.Hcc[c]) for c in range(C)])
    pred_cam = float(np.sum(d * (lam * Hd * d - lin.gc.reshape(n)) * f))
    return d.reshape(C, 6), pred_cam, True


def back_substitute(X, R, t, K, x2d, w, lam, delta_c):
    """delta_p = -Hpp_d^-1 (gp + W delta_c); also the point part of the predicted decrease."""
    e, A, B, _ = residual_blocks(X, R, t, K, x2d, w)
    wA = A * w[..., None, None]
    Hpp = np.einsum("ncij,ncik->njk", wA, A)
    gp = np.einsum("ncij,nci->nj", wA, e)
    lin_e = e + np.einsum("ncik,ck->nci", B, delta_c)  # residual after the (linearised) camera move
    rhs = np.einsum("ncij,nci->nj", wA, lin_e)  # gp + W delta_c
    diag = np.einsum("nii->ni", Hpp)
    Hd = Hpp + lam * np.einsum("ni,ij->nij", diag, np.eye(3))
    dp = -np.einsum("nij,nj->ni", _safe_inverse(Hd), rhs)
    pred_pts = float(np.sum(dp * (lam * diag * dp - gp)))
    return dp, pred_pts


def apply_camera_step(R, t, delta_c):
    R2 = np.stack([G.so3_exp(delta_c[c, :3]) @ R[c] for c in range(len(R))])
    return R2, t + delta_c[:, 3:]


def nielsen_update(lam, nu, rho, accepted):
    if accepted:
        return lam * max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3), 2.0
    return lam * nu, 2.0 * nu


def run_lm(X0, R0, t0, K, x2d, conf, num_iters=20, mode="full", lam0=LAMBDA0, shards=1):
    """X0 (T,J,3); R0 (C,3,3); t0 (C,3); K (C,3,3); x2d (T,C,J,2); conf (T,C,J).
    Returns R, t, X (T,J,3), history (list of dict).  `shards` > 1 computes every reduction as a sum
    of per-frame-range partials (what the multi-GPU path does) - results must not depend on it."""
    T, J, _ = X0.shape
    C = R0.shape[0]
    X = np.asarray(X0, float).reshape(T * J, 3).copy()
    R, t = np.asarray(R0, float).copy(), np.asarray(t0, float).copy()
    K = np.asarray(K, float)
    x = np.asarray(x2d, float).transpose(0, 2, 1, 3).reshape(T * J, C, 2)
    cf = np.asarray(conf, float).transpose(0, 2, 1).reshape(T * J, C)
    w = cf / (cf.sum() + 1e-6)
    free = free_mask(C, mode)
    bounds = np.linspace(0, T, shards + 1).astype(int) * J
    parts = [slice(bounds[i], bounds[i + 1]) for i in range(shards)]
    lam, nu = float(lam0), 2.0
    history = []
    for it in range(num_iters):
        lin = None
        for s in parts:
            li = linearise(X[s], R, t, K, x[s], w[s], lam)
            lin = li if lin is None else lin + li
        dc, pred_cam, ok = solve_reduced(lin, lam, free)
        Rn, tn = apply_camera_step(R, t, dc)
        Xn = X.copy()
        pred, Ft, ncl = pred_cam, 0.0, 0
        for s in parts:
            dp, pp = back_substitute(X[s], R, t, K, x[s], w[s], lam, dc)
            Xn[s] = X[s] + dp
            pred += pp
            c, k = cost_only(Xn[s], Rn, tn, K, x[s], w[s])
            Ft += c
            ncl += k
        F = lin.cost
        rho = (F - Ft) / pred if pred > 0 else 0.0
        accepted = bool(ok and np.isfinite(Ft) and Ft < F)
        history.append(dict(iter=it, cost=F, trial_cost=Ft, lam=lam, rho=rho, accepted=accepted, n_clamped=lin.n_clamped, pred=pred))
        lam, nu = nielsen_update(lam, nu, rho, accepted)
        if accepted:
            X, R, t = Xn, Rn, tn
    return R, t, X.reshape(T, J, 3), history


def dense_step(X, R, t, K, x2d, w, lam, free):
    """Cross-check of the Schur algebra: assemble the full J and solve (J^T W J + lam D) d = -J^T W r
    directly.  Small problems only."""
    N, C = w.shape
    e, A, B, _ = residual_blocks(X, R, t, K, x2d, w)
    n_c, n_p = 6 * C, 3 * N
    Jf = np.zeros((N, C, 2, n_c + n_p))
    for c in range(C):
        Jf[:, c, :, 6 * c : 6 * c + 6] = B[:, c]
    for i in range(N):
        Jf[i, :, :, n_c + 3 * i : n_c + 3 * i + 3] = A[i]
    Jm = Jf.reshape(-1, n_c + n_p)
    sw = np.sqrt(np.repeat(w.reshape(-1), 2))
    Jw = Jm * sw[:, None]
    r = e.reshape(-1) * sw
    keep = np.concatenate([free.reshape(-1), np.ones(n_p, bool)])
    Jw = Jw[:, keep]
    H = Jw.T @ Jw
    g = Jw.T @ r
    d = np.linalg.solve(H + lam * np.diag(np.diag(H)), -g)
    full = np.zeros(n_c + n_p)
    full[keep] = d
    return full[:n_c].reshape(C, 6), full[n_c:].reshape(N, 3)
