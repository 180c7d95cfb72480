"""fp64 Schur-complement LM with FREE INTRINSICS AND DISTORTION - the oracle of the calibrating BA path
(BASELINE config 3: "2 cameras, Rodrigues extrinsics + intrinsics/distortion"; SURVEY.md section 8d "p = 15").

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY STATUS: "parity unpinned" for the LM trajectory, exactly like oracle/lm.py (the reference's
``run_local_ba`` is an undefined symbol, vggt/multi_view_process.py:553-564).  What IS pinned:
  * the projection model is cv2.projectPoints' 5-coefficient one, the model the reference reprojects with
    (triangulation/reproject.py:77-78, bundle_adjustment/reproject.py:147-148): values and the
    [tvec, f, c, dist] Jacobian columns are checked against cv2.projectPoints itself in tests/test_oracle_lm_calib.py;
  * with zero distortion, zero-skew K and only the extrinsics free this file reproduces oracle/lm.py's
    trajectory (and therefore the reference's ``reprojection_loss`` values, golden G3/G6) to rounding.

Model (per camera c, 15 parameters [d_omega(3), d_t(3), fx, fy, cx, cy, k1, k2, p1, p2, k3]):
  X_c = R X + t;  x = X_c.x/z, y = X_c.y/z; an observation with z < 1e-6 (loss.py:67's clamp threshold: the point is at /
  behind the camera plane) is EXCLUDED - zero residual and Jacobians, counted in n_clamped
  r2 = x^2 + y^2;  rad = 1 + k1 r2 + k2 r2^2 + k3 r2^3
  x" = x rad + 2 p1 x y + p2 (r2 + 2 x^2);  y" = y rad + p1 (r2 + 2 y^2) + 2 p2 x y
  u = fx x" + cx;  v = fy y" + cy
Cost  F = sum_tcj w_tcj |pi_c(X_tj) - x_tcj|^2 + sum_c sum_k rho_ck (theta_ck - theta0_ck)^2,
      w = conf / (sum conf + 1e-6); theta = the 9 intrinsic parameters; rho >= 0 is an optional Gaussian prior
      (free intrinsics on a fixating rig leave near-null directions; the prior pins them).
Update R <- exp([d_omega]x) R, t <- t + d_t, theta <- theta + d_theta; camera 0's extrinsics are the gauge.
Step / control: identical to oracle/lm.py (Marquardt damping lam*diag(H) on both blocks with the prior inside H,
Schur complement, Cholesky, back-substitution, Nielsen gain ratio).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import geometry as G
from .lm import LAMBDA0, ZMIN, _safe_inverse, nielsen_update

P = 15   # parameters per camera
NI = 9   # intrinsic parameters: fx fy cx cy k1 k2 p1 p2 k3

CALIB_MODES = ("full", "extr_focal", "intr_only")


def free_mask(C: int, mode: str = "full") -> np.ndarray:
    """(C,15) bool.  Camera 0's extrinsics are always fixed (gauge).
    full: everything else free; extr_focal: extrinsics + fx, fy, cx, cy (the "p = 10" block of SURVEY 8d);
    intr_only: the 9 intrinsic parameters of every camera, extrinsics fixed."""
    m = np.zeros((C, P), bool)
    if mode == "full":
        m[:, 6:] = True
        m[1:, :6] = True
    elif mode == "extr_focal":
        m[:, 6:10] = True
        m[1:, :6] = True
    elif mode == "intr_only":
        m[:, 6:] = True
    else:
        raise ValueError(f"unknown mode {mode!r}")
    return m


def intr_from_K(K, dist=None) -> np.ndarray:
    """(C,3,3) K (zero skew) + optional (C,5)/(5,) [k1 k2 p1 p2 k3] -> (C,9) intrinsic vectors."""
    K = np.asarray(K, float)
    C = K.shape[0]
    th = np.zeros((C, NI))
    th[:, 0], th[:, 1], th[:, 2], th[:, 3] = K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2]
    if dist is not None:
        th[:, 4:] = np.broadcast_to(np.asarray(dist, float), (C, 5))
    return th


def project(X, R, t, th):
    """X (N,3) -> pixels (N,C,2), clamped (N,C)."""
    Xc = np.einsum("cab,nb->nca", R, X) + t[None]
    z = Xc[..., 2]
    Z = np.maximum(z, ZMIN)
    x, y = Xc[..., 0] / Z, Xc[..., 1] / Z
    fx, fy, cx, cy, k1, k2, p1, p2, k3 = (th[None, :, i] for i in range(NI))
    r2 = x * x + y * y
    rad = 1 + r2 * (k1 + r2 * (k2 + r2 * k3))
    xd = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    return np.stack([fx * xd + cx, fy * yd + cy], -1), z < ZMIN


def residual_blocks(X, R, t, th, x2d):
    """e (N,C,2), A (N,C,2,3) = d(u,v)/dX, B (N,C,2,15) = d(u,v)/d[omega, t, theta], clamped (N,C)."""
    Xc = np.einsum("cab,nb->nca", R, X) + t[None]
    z = Xc[..., 2]
    clamped = z < ZMIN
    live = (~clamped).astype(float)
    iz = 1.0 / np.maximum(z, ZMIN)
    x, y = Xc[..., 0] * iz, Xc[..., 1] * iz
    fx, fy, cx, cy, k1, k2, p1, p2, k3 = (th[None, :, i] for i in range(NI))
    r2 = x * x + y * y
    rad = 1 + r2 * (k1 + r2 * (k2 + r2 * k3))
    drad = k1 + r2 * (2 * k2 + 3 * k3 * r2)
    xd = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    e = np.stack([fx * xd + cx, fy * yd + cy], -1) - x2d
    a11 = rad + 2 * x * x * drad + 2 * p1 * y + 6 * p2 * x
    a12 = 2 * x * y * drad + 2 * p1 * x + 2 * p2 * y
    a22 = rad + 2 * y * y * drad + 6 * p1 * y + 2 * p2 * x
    ju = np.stack([fx * a11 * iz, fx * a12 * iz, -fx * (a11 * x + a12 * y) * iz * live], -1)
    jv = np.stack([fy * a12 * iz, fy * a22 * iz, -fy * (a12 * x + a22 * y) * iz * live], -1)
    Jx = np.stack([ju, jv], -2)  # (N,C,2,3)
    A = np.einsum("ncij,cjk->ncik", Jx, R)
    p = Xc - t[None]
    Bw = np.cross(p[..., None, :], Jx)
    zero, one = np.zeros_like(x), np.ones_like(x)
    r4, r6 = r2 * r2, r2 * r2 * r2
    Bu = np.stack([xd, zero, one, zero, fx * x * r2, fx * x * r4, fx * 2 * x * y, fx * (r2 + 2 * x * x), fx * x * r6], -1)
    Bv = np.stack([zero, yd, zero, one, fy * y * r2, fy * y * r4, fy * (r2 + 2 * y * y), fy * 2 * x * y, fy * y * r6], -1)
    B = np.concatenate([Bw, Jx, np.stack([Bu, Bv], -2)], -1)
    # a depth-clamped observation is EXCLUDED (zero residual and Jacobians, still counted): the polynomial distortion of a
    # clamped projection (|x| ~ 1e6) is meaningless and overflows fp32 - see csrc/ska_ba_calib.cuh
    keep = ~clamped
    return np.where(keep[..., None], e, 0.0), np.where(keep[..., None, None], A, 0.0), np.where(keep[..., None, None], B, 0.0), clamped


def cost_only(X, R, t, th, x2d, w):
    uv, clamped = project(X, R, t, th)
    d = np.where(clamped[..., None], 0.0, uv - x2d)  # excluded observations
    return float((w[..., None] * d**2).sum()), int(clamped.sum())


def prior_cost(th, th0, rho):
    return float((rho * (th - th0) ** 2).sum())


@dataclass
class Linearisation:
    Hcc: np.ndarray  # (C,15,15) undamped, data term only
    gc: np.ndarray   # (C,15)
    Sw: np.ndarray   # (15C,15C)
    bw: np.ndarray   # (15C,)
    cost: float      # data term only
    n_clamped: int

    def __add__(self, o):
        return Linearisation(self.Hcc + o.Hcc, self.gc + o.gc, self.Sw + o.Sw, self.bw + o.bw, self.cost + o.cost,
                             self.n_clamped + o.n_clamped)


def linearise(X, R, t, th, x2d, w, lam) -> Linearisation:
    N, C = w.shape
    e, A, B, clamped = residual_blocks(X, R, t, th, x2d)
    wA = A * w[..., None, None]
    Hpp = np.einsum("ncij,ncik->njk", wA, A)
    gp = np.einsum("ncij,nci->nj", wA, e)
    wB = B * w[..., None, None]
    Hcc = np.einsum("ncij,ncik->cjk", wB, B)
    gc = np.einsum("ncij,nci->cj", wB, e)
    W = np.einsum("ncij,ncik->njck", wA, B).reshape(N, 3, P * C)
    Hd = Hpp + lam * np.einsum("nii,ij->nij", Hpp, np.eye(3))
    Hinv = _safe_inverse(Hd)
    Sw = np.einsum("nik,nij,njl->kl", W, Hinv, W)
    bw = np.einsum("nik,nij,nj->k", W, Hinv, gp)
    return Linearisation(Hcc, gc, Sw, bw, float((w[..., None] * e**2).sum()), int(clamped.sum()))


def solve_reduced(lin: Linearisation, lam, free, th, th0, rho):
    """S = (Hcc + prior)(1 + lam on the diagonal) - Sw on the free parameters, Cholesky.
    Returns delta (C,15), pred_cam, ok."""
    C = lin.Hcc.shape[0]
    n = P * C
    S = -lin.Sw.copy()
    g = lin.gc.copy()
    hd = np.zeros((C, P))
    for c in range(C):
        H = lin.Hcc[c].copy()
        H[6:, 6:] += np.diag(rho[c])
        g[c, 6:] += rho[c] * (th[c] - th0[c])
        hd[c] = np.diag(H)
        S[P * c : P * c + P, P * c : P * c + P] += H + lam * np.diag(hd[c])
    b = -g.reshape(n) + lin.bw
    f = free.reshape(n)
    S[~f, :] = 0.0
    S[:, ~f] = 0.0
    S[~f, ~f] = 1.0
    b = np.where(f, b, 0.0)
    try:
        L = np.linalg.cholesky(S)
    except np.linalg.LinAlgError:
        return np.zeros((C, P)), 0.0, False
    d = np.linalg.solve(L.T, np.linalg.solve(L, b))
    pred_cam = float(np.sum(d * (lam * hd.reshape(n) * d - g.reshape(n)) * f))
    return d.reshape(C, P), pred_cam, True


def back_substitute(X, R, t, th, x2d, w, lam, delta):
    e, A, B, _ = residual_blocks(X, R, t, th, x2d)
    wA = A * w[..., None, None]
    Hpp = np.einsum("ncij,ncik->njk", wA, A)
    gp = np.einsum("ncij,nci->nj", wA, e)
    lin_e = e + np.einsum("ncik,ck->nci", B, delta)
    rhs = np.einsum("ncij,nci->nj", wA, lin_e)
    diag = np.einsum("nii->ni", Hpp)
    Hd = Hpp + lam * np.einsum("ni,ij->nij", diag, np.eye(3))
    dp = -np.einsum("nij,nj->ni", _safe_inverse(Hd), rhs)
    return dp, float(np.sum(dp * (lam * diag * dp - gp)))


def apply_camera_step(R, t, th, delta):
    R2 = np.stack([G.so3_exp(delta[c, :3]) @ R[c] for c in range(len(R))])
    return R2, t + delta[:, 3:6], th + delta[:, 6:]


def run_lm(X0, R0, t0, th0, x2d, conf, num_iters=20, free=None, lam0=LAMBDA0, prior_theta=None, prior_rho=None, shards=1):
    """X0 (T,J,3); R0 (C,3,3); t0 (C,3); th0 (C,9); x2d (T,C,J,2); conf (T,C,J); free (C,15) bool.
    Returns R, t, theta, X (T,J,3), history (rows as oracle/lm.py; `cost` / `trial_cost` include the prior)."""
    T, J, _ = X0.shape
    C = R0.shape[0]
    X = np.asarray(X0, float).reshape(T * J, 3).copy()
    R, t, th = np.asarray(R0, float).copy(), np.asarray(t0, float).copy(), np.asarray(th0, float).copy()
    x = np.asarray(x2d, float).transpose(0, 2, 1, 3).reshape(T * J, C, 2)
    cf = np.asarray(conf, float).transpose(0, 2, 1).reshape(T * J, C)
    w = cf / (cf.sum() + 1e-6)
    free = free_mask(C) if free is None else np.asarray(free, bool).copy()
    free[0, :6] = False
    pth = th.copy() if prior_theta is None else np.asarray(prior_theta, float)
    rho = np.zeros((C, NI)) if prior_rho is None else np.broadcast_to(np.asarray(prior_rho, float), (C, NI)).copy()
    bounds = np.linspace(0, T, shards + 1).astype(int) * J
    parts = [slice(bounds[i], bounds[i + 1]) for i in range(shards)]
    lam, nu = float(lam0), 2.0
    history = []
    for it in range(num_iters):
        lin = None
        for s in parts:
            li = linearise(X[s], R, t, th, x[s], w[s], lam)
            lin = li if lin is None else lin + li
        dc, pred_cam, ok = solve_reduced(lin, lam, free, th, pth, rho)
        Rn, tn, thn = apply_camera_step(R, t, th, dc)
        Xn = X.copy()
        pred, Ft, ncl = pred_cam, 0.0, 0
        for s in parts:
            dp, pp = back_substitute(X[s], R, t, th, x[s], w[s], lam, dc)
            Xn[s] = X[s] + dp
            pred += pp
            c, k = cost_only(Xn[s], Rn, tn, thn, x[s], w[s])
            Ft += c
            ncl += k
        F = lin.cost + prior_cost(th, pth, rho)
        Ft += prior_cost(thn, pth, rho)
        rho_gain = (F - Ft) / pred if pred > 0 else 0.0
        accepted = bool(ok and np.isfinite(Ft) and Ft < F)
        history.append(dict(iter=it, cost=F, trial_cost=Ft, lam=lam, rho=rho_gain, accepted=accepted,
                            n_clamped=lin.n_clamped, pred=pred))
        lam, nu = nielsen_update(lam, nu, rho_gain, accepted)
        if accepted:
            X, R, t, th = Xn, Rn, tn, thn
    return R, t, th, X.reshape(T, J, 3), history


def dense_step(X, R, t, th, x2d, w, lam, free, th0, rho):
    """Cross-check of the Schur algebra: assemble the full Jacobian (data + prior rows) and solve
    (J^T J + lam diag(J^T J)) d = -J^T r directly.  Small problems only."""
    N, C = w.shape
    e, A, B, _ = residual_blocks(X, R, t, th, x2d)
    n_c, n_p = P * C, 3 * N
    Jf = np.zeros((N, C, 2, n_c + n_p))
    for c in range(C):
        Jf[:, c, :, P * c : P * c + P] = B[:, c]
    for i in range(N):
        Jf[i, :, :, n_c + 3 * i : n_c + 3 * i + 3] = A[i]
    sw = np.sqrt(np.repeat(w.reshape(-1), 2))
    Jw = Jf.reshape(-1, n_c + n_p) * sw[:, None]
    r = e.reshape(-1) * sw
    Jp = np.zeros((C * NI, n_c + n_p))
    rp = np.zeros(C * NI)
    for c in range(C):
        for k in range(NI):
            Jp[c * NI + k, P * c + 6 + k] = np.sqrt(rho[c, k])
            rp[c * NI + k] = np.sqrt(rho[c, k]) * (th[c, k] - th0[c, k])
    Jw = np.concatenate([Jw, Jp])
    r = np.concatenate([r, rp])
    keep = np.concatenate([free.reshape(-1), np.ones(n_p, bool)])
    Jw = Jw[:, keep]
    H = Jw.T @ Jw
    d = np.linalg.solve(H + lam * np.diag(np.diag(H)), -(Jw.T @ r))
    full = np.zeros(n_c + n_p)
    full[keep] = d
    return full[:n_c].reshape(C, P), full[n_c:].reshape(N, 3)


from skiing_analysis_pytorch_b200.synth import CALIB_PRIOR_RHO as PRIOR_RHO  # noqa: E402  (problem definition shared with bench.py)
from skiing_analysis_pytorch_b200.synth import perturb_intrinsics  # noqa: E402,F401


def make_problem(rig: str, T: int, J: int, seed: int = 0):
    """oracle/lm.py's problem (perturbed extrinsics, DLT points) plus perturbed intrinsics.
    Returns (clip, R0, t0, theta_init, X0)."""
    from . import lm

    clip, R0, t0, X0 = lm.make_problem(rig, T, J, seed)
    return clip, R0, t0, perturb_intrinsics(clip.K), X0
