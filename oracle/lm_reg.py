"""fp64 Levenberg-Marquardt over the reference's FULL configured objective - the oracle of the regularised
`run_local_ba` (second-order form; multi-GPU row e3).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Objective (the reference's own terms, per-frame cameras as the call site passes them, vggt/multi_view_process.py:546-564;
weights configs/vggt.yaml:46-50):
    F = w_r reprojection_loss  + w_s camera_smooth_loss + w_b baseline_reg_loss + w_l bone_length_loss + w_t pose_temporal_loss
                 loss.py:90-94            loss.py:103-106         loss.py:109-114          loss.py:134-150          loss.py:153-155
written as a sum of squares F = |r|^2 with residual groups
    reproj    sqrt(c_r conf) (pi(R X + t) - x),           c_r = w_r / (sum conf + 1e-6)
    bone      sqrt(c_l) (|X_i - X_j| - ref_b),             c_l = w_l / (T B),  ref_b = mean_t |X_i - X_j| (detached: a constant
                                                                                of the linearisation, loss.py:141-146)
    temporal  sqrt(c_t) (X[t+1] - X[t]),                   c_t = w_t / ((T-1) J 3)
    smooth    sqrt(c_s) (Cc[t+1] - Cc[t]),  Cc = -R^T t,   c_s = w_s / ((T-1) C 3)
    baseline  sqrt(c_b) (|Cc_0 - Cc_1| - mean_t |..|),     c_b = w_b / T   (the mean detached like loss.py:114)
PARITY: the VALUE F is the reference's (tests compare `cost` with oracle/torch_ref.py, itself pinned by the reference's
functions through goldens G3/G4/G9, and golden G12 holds the reference's loss at this file's iterates); the optimiser is
"parity unpinned" - run_local_ba is an undefined symbol in the reference - so this file is its specification:
  unknowns  X (T,J,3) always; t (T,C,3) in "pose_cam_t" and "full"; R (T,C,3,3) in "full", moved on SO(3) through the left
            tangent R <- exp([d_omega]x) R.  EVERY camera of every frame is free (no gauge camera: damping and the camera
            regularisers hold the gauge, as in the first-order form oracle/first_order.py).
  step      Gauss-Newton with Marquardt damping: (J^T J + lam diag(J^T J)) delta = -J^T r, solved exactly here (sparse LU);
            the CUDA path solves the same system by block-Jacobi preconditioned conjugate gradients to a relative
            residual of 1e-8 (the coupling terms make J^T J block-tridiagonal in time, not block-diagonal per point).
  control   as oracle/lm.py: accept iff F_trial < F; rho = (F - F_trial) / (delta^T (lam D delta - g)); Nielsen's lambda
            update, lam0 = 1e-3.  F_trial is the true objective at the trial point (its own bone / baseline means).
  pi        as loss.py:17-87: X_c = R X + t, z clamped at 1e-6 with zero derivative while the clamp is active.
Multi-shard: tests/oracle_engine.py (OracleRegularisedBundleAdjuster) runs this specification shard by shard - each rank
multiplies only its rows of the normal matrix with [halo | own | halo] entries - behind the product's sequencer over gloo.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import geometry as G

ZMIN = 1e-6
LAMBDA0 = 1e-3
MODES = ("pose_only", "pose_cam_t", "full")
DEFAULT_WEIGHTS = dict(reproj=1.0, smooth=0.1, baseline=0.01, bone_length=0.1, pose_temporal=0.1)  # configs/vggt.yaml:46-50
BONES = [(11, 13), (13, 15), (12, 14), (14, 16), (5, 7), (7, 9), (6, 8), (8, 10), (5, 6), (11, 12), (5, 11), (6, 12)]  # loss.py:118-131
TERMS = ("reproj", "smooth", "baseline", "bone_length", "pose_temporal")

# golden set G12 (tests/golden/g12_lm_reg.npz): name -> (rig, T, J, mode)
G12_CASES = {"c3_pose": ("2b", 60, 17, "pose_only"), "c3_t": ("2b", 40, 17, "pose_cam_t"), "c3_full": ("2b", 40, 17, "full"),
             "c5_pose": ("8", 12, 70, "pose_only"), "c4_full": ("4", 16, 17, "full")}


def bones_for(J):
    return [(i, j) for i, j in BONES if i < J and j < J]


def centres(R, t):
    return -np.einsum("tcba,tcb->tca", R, t)


def coefficients(T, J, C, conf_sum, weights):
    w = dict(DEFAULT_WEIGHTS, **(weights or {}))
    nb = len(bones_for(J))
    return dict(
        reproj=w["reproj"] / (conf_sum + 1e-6),
        bone_length=w["bone_length"] / (T * nb) if nb else 0.0,
        pose_temporal=w["pose_temporal"] / ((T - 1) * J * 3) if T > 1 else 0.0,
        smooth=w["smooth"] / ((T - 1) * C * 3) if T > 1 else 0.0,
        baseline=w["baseline"] / T if C >= 2 else 0.0,
    )


def project(X, R, t, K):
    """X (T,J,3); R (T,C,3,3); t (T,C,3); K (C,3,3) -> Xc (T,C,J,3), uv (T,C,J,2), Jx (T,C,J,2,3) = d(u,v)/dX_c, clamped."""
    Xc = np.einsum("tcab,tjb->tcja", R, X) + t[:, :, None, :]
    z = Xc[..., 2]
    clamped = z < ZMIN
    iz = 1.0 / np.maximum(z, ZMIN)
    x, y = Xc[..., 0] * iz, Xc[..., 1] * iz
    k = K[None, :, None]
    u = k[..., 0, 0] * x + k[..., 0, 1] * y + k[..., 0, 2]
    v = k[..., 1, 0] * x + k[..., 1, 1] * y + k[..., 1, 2]
    live = (~clamped).astype(float)
    ju = np.stack([k[..., 0, 0] * iz, k[..., 0, 1] * iz, -(k[..., 0, 0] * x + k[..., 0, 1] * y) * iz * live], -1)
    jv = np.stack([k[..., 1, 0] * iz, k[..., 1, 1] * iz, -(k[..., 1, 0] * x + k[..., 1, 1] * y) * iz * live], -1)
    return Xc, np.stack([u, v], -1), np.stack([ju, jv], -2), clamped


def cost_terms(X, R, t, K, x2d, conf, coef):
    """The five terms of F (each already weighted) at one point; the bone / baseline references are the means AT this point."""
    T, J, _ = X.shape
    _, uv, _, clamped = project(X, R, t, K)
    d = ((uv - x2d) ** 2).sum(-1)
    out = dict(reproj=coef["reproj"] * float((conf * d).sum()))
    bl = bones_for(J)
    if bl and coef["bone_length"]:
        L = np.stack([np.linalg.norm(X[:, i] - X[:, j], axis=-1) for i, j in bl], -1)
        out["bone_length"] = coef["bone_length"] * float(((L - L.mean(0, keepdims=True)) ** 2).sum())
    else:
        out["bone_length"] = 0.0
    out["pose_temporal"] = coef["pose_temporal"] * float(((X[1:] - X[:-1]) ** 2).sum())
    Cc = centres(R, t)
    out["smooth"] = coef["smooth"] * float(((Cc[1:] - Cc[:-1]) ** 2).sum())
    if R.shape[1] >= 2 and coef["baseline"]:
        b = np.linalg.norm(Cc[:, 0] - Cc[:, 1], axis=-1)
        out["baseline"] = coef["baseline"] * float(((b - b.mean()) ** 2).sum())
    else:
        out["baseline"] = 0.0
    return out, int(clamped.sum())


def free_columns(T, J, C, mode):
    """Boolean mask over the stacked unknown vector [X (T,J,3) | cams (T,C,6) = d_omega, d_t]."""
    m = np.zeros(T * J * 3 + T * C * 6, bool)
    m[: T * J * 3] = True
    cam = np.zeros((T, C, 6), bool)
    if mode == "full":
        cam[:] = True
    elif mode == "pose_cam_t":
        cam[..., 3:] = True
    elif mode != "pose_only":
        raise ValueError(f"unknown mode {mode!r}")
    m[T * J * 3:] = cam.reshape(-1)
    return m


def jacobian(X, R, t, K, x2d, conf, coef, refs=None):
    """Residual vector r and sparse Jacobian J over ALL unknowns [X | cams] (columns masked later).
    refs: None = the bone / baseline references are the means AT this point (what the LM uses: constants of the linearisation);
    (ref_bone (B,), ref_base) = given constants (finite-difference checks hold them fixed while the point moves)."""
    T, J, _ = X.shape
    C = R.shape[1]
    nX = T * J * 3
    n = nX + T * C * 6
    rows, cols, vals, res = [], [], [], []
    nrow = 0

    def add(r_idx, c_idx, v):
        rows.append(np.asarray(r_idx).reshape(-1))
        cols.append(np.asarray(c_idx).reshape(-1))
        vals.append(np.asarray(v, float).reshape(-1))

    ix = lambda tt, jj: (tt * J + jj) * 3  # noqa: E731
    ic = lambda tt, cc: nX + (tt * C + cc) * 6  # noqa: E731
    # ---- reprojection
    Xc, uv, Jx, _ = project(X, R, t, K)
    s = np.sqrt(coef["reproj"] * conf)  # (T,C,J)
    e = (uv - x2d) * s[..., None]
    A = np.einsum("tcjik,tckl->tcjil", Jx, R) * s[..., None, None]  # d/dX
    p = Xc - t[:, :, None, :]
    Bw = np.cross(p[..., None, :], Jx) * s[..., None, None]  # d/d_omega rows p x ju, p x jv
    Bt = Jx * s[..., None, None]
    tt, cc, jj = np.meshgrid(np.arange(T), np.arange(C), np.arange(J), indexing="ij")
    r0 = nrow + ((tt * C + cc) * J + jj) * 2
    for i in range(2):
        for k in range(3):
            add(r0 + i, ix(tt, jj) + k, A[..., i, k])
            add(r0 + i, ic(tt, cc) + k, Bw[..., i, k])
            add(r0 + i, ic(tt, cc) + 3 + k, Bt[..., i, k])
    res.append(e.reshape(-1))
    nrow += T * C * J * 2
    # ---- bone length (reference length = the mean at this point, a constant of the linearisation)
    bl = bones_for(J)
    if bl and coef["bone_length"]:
        sq = np.sqrt(coef["bone_length"])
        tv = np.arange(T)
        for bi, bj in bl:
            d = X[:, bi] - X[:, bj]
            L = np.linalg.norm(d, axis=-1)
            u = d / L[:, None]
            rr = nrow + tv
            for k in range(3):
                add(rr, ix(tv, bi) + k, sq * u[:, k])
                add(rr, ix(tv, bj) + k, -sq * u[:, k])
            res.append(sq * (L - (L.mean() if refs is None else refs[0][len(res) - 1])))
            nrow += T
    # ---- pose temporal
    if T > 1 and coef["pose_temporal"]:
        sq = np.sqrt(coef["pose_temporal"])
        tv, jv, kv = np.meshgrid(np.arange(T - 1), np.arange(J), np.arange(3), indexing="ij")
        rr = nrow + (tv * J + jv) * 3 + kv
        add(rr, ix(tv + 1, jv) + kv, np.full(rr.shape, sq))
        add(rr, ix(tv, jv) + kv, np.full(rr.shape, -sq))
        res.append((sq * (X[1:] - X[:-1])).reshape(-1))
        nrow += (T - 1) * J * 3
    # ---- camera centres: Cc = -R^T t;  dCc/d_t = -R^T;  dCc/d_omega = -R^T [t]x   (R <- exp([w]x) R)
    Cc = centres(R, t)
    Jt = -np.swapaxes(R, -1, -2)  # (T,C,3,3)
    Jw = np.einsum("tcab,tcbk->tcak", Jt, G.hat(t))
    JC = np.concatenate([Jw, Jt], -1)  # (T,C,3,6)
    if T > 1 and coef["smooth"]:
        sq = np.sqrt(coef["smooth"])
        tv, cv, av = np.meshgrid(np.arange(T - 1), np.arange(C), np.arange(3), indexing="ij")
        rr = nrow + (tv * C + cv) * 3 + av
        for k in range(6):
            add(rr, ic(tv + 1, cv) + k, sq * JC[tv + 1, cv, av, k])
            add(rr, ic(tv, cv) + k, -sq * JC[tv, cv, av, k])
        res.append((sq * (Cc[1:] - Cc[:-1])).reshape(-1))
        nrow += (T - 1) * C * 3
    if C >= 2 and coef["baseline"]:
        sq = np.sqrt(coef["baseline"])
        d = Cc[:, 0] - Cc[:, 1]
        b = np.linalg.norm(d, axis=-1)
        nh = d / b[:, None]
        tv = np.arange(T)
        rr = nrow + tv
        g0 = np.einsum("ta,tak->tk", nh, JC[:, 0])
        g1 = -np.einsum("ta,tak->tk", nh, JC[:, 1])
        for k in range(6):
            add(rr, ic(tv, 0) + k, sq * g0[:, k])
            add(rr, ic(tv, 1) + k, sq * g1[:, k])
        res.append(sq * (b - (b.mean() if refs is None else refs[1])))
        nrow += T
    Jm = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nrow, n)).tocsr()
    return np.concatenate(res), Jm


def apply_step(X, R, t, delta, mode):
    T, J, _ = X.shape
    C = R.shape[1]
    Xn = X + delta[: T * J * 3].reshape(T, J, 3)
    dc = delta[T * J * 3:].reshape(T, C, 6)
    tn = t + dc[..., 3:] if mode != "pose_only" else t
    if mode == "full":
        Rn = np.stack([np.stack([G.so3_exp(dc[a, c, :3]) @ R[a, c] for c in range(C)]) for a in range(T)])
    else:
        Rn = R
    return Xn, Rn, tn


def normal_system(X, R, t, K, x2d, conf, coef, mode):
    """H = J^T J over the free unknowns, g = J^T r, the free mask, and the residual norm."""
    r, Jm = jacobian(X, R, t, K, x2d, conf, coef)
    free = free_columns(X.shape[0], X.shape[1], R.shape[1], mode)
    Jf = Jm[:, np.flatnonzero(free)]
    return (Jf.T @ Jf).tocsc(), Jf.T @ r, free, float(r @ r)


def pcg(H, D, lam, g, blocks, tol=1e-8, max_iter=500):
    """The solver the CUDA path runs: conjugate gradients on (H + lam D) d = -g, block-Jacobi preconditioner (the diagonal
    blocks of H + lam D listed in `blocks` as index arrays).  Returns d and the iteration count."""
    A = (H + lam * sp.diags(D)).tocsr()
    Minv = [(b, np.linalg.inv(A[b][:, b].toarray())) for b in blocks]

    def prec(v):
        z = np.zeros_like(v)
        for b, Mi in Minv:
            z[b] = Mi @ v[b]
        return z

    x = np.zeros_like(g)
    r = -g
    z = prec(r)
    p = z.copy()
    rz = rz0 = float(r @ z)
    k = 0
    while k < max_iter and rz > tol * tol * rz0:
        y = A @ p
        alpha = rz / float(p @ y)
        x += alpha * p
        r -= alpha * y
        z = prec(r)
        rz_new = float(r @ z)
        p = z + (rz_new / rz) * p
        rz = rz_new
        k += 1
    return x, k


def jacobi_blocks(T, J, C, mode):
    """Index arrays (into the FREE unknown vector) of the block-Jacobi preconditioner: one 3x3 block per point, one block per
    (frame, camera) over its free parameters."""
    blocks = [np.arange(3 * i, 3 * i + 3) for i in range(T * J)]
    nX = T * J * 3
    if mode == "pose_cam_t":
        blocks += [nX + np.arange(3 * i, 3 * i + 3) for i in range(T * C)]
    elif mode == "full":
        blocks += [nX + np.arange(6 * i, 6 * i + 6) for i in range(T * C)]
    return blocks


def run_lm(X0, R0, t0, K, x2d, conf, num_iters=10, mode="pose_only", weights=None, lam0=LAMBDA0, solver="direct", cg_tol=1e-8,
           cg_max_iter=500):
    """X0 (T,J,3); R0 (T,C,3,3) | (C,3,3); t0 (T,C,3) | (C,3); K (C,3,3); x2d (T,C,J,2); conf (T,C,J).
    Returns R (T,C,3,3), t (T,C,3), X (T,J,3), history (list of dict: iter, cost, trial_cost, lam, rho, accepted, pred,
    n_clamped, cg_iters, one entry per term of the cost before the step)."""
    if mode not in MODES:
        raise ValueError(f"unknown mode {mode!r}")
    X = np.asarray(X0, float).copy()
    T, J, _ = X.shape
    R = np.asarray(R0, float)
    t = np.asarray(t0, float)
    if R.ndim == 3:
        R, t = np.broadcast_to(R, (T,) + R.shape).copy(), np.broadcast_to(t, (T,) + t.shape).copy()
    R, t = R.copy(), t.copy()
    C = R.shape[1]
    K = np.broadcast_to(np.asarray(K, float), (C, 3, 3))
    x2d, conf = np.asarray(x2d, float), np.asarray(conf, float)
    coef = coefficients(T, J, C, float(conf.sum()), weights)
    lam, nu = float(lam0), 2.0
    hist = []
    blocks = jacobi_blocks(T, J, C, mode) if solver == "pcg" else None
    for it in range(num_iters):
        terms, ncl = cost_terms(X, R, t, K, x2d, conf, coef)
        F = sum(terms.values())
        H, g, free, r2 = normal_system(X, R, t, K, x2d, conf, coef, mode)
        # every residual group is in r, so |r|^2 == F whatever the mode (tested)
        D = H.diagonal()
        if solver == "direct":
            d_free = spla.splu((H + lam * sp.diags(D)).tocsc()).solve(-g)
            cg_iters = 0
        else:
            d_free, cg_iters = pcg(H, D, lam, g, blocks, cg_tol, cg_max_iter)
        delta = np.zeros(free.shape[0])
        delta[free] = d_free
        pred = float(d_free @ (lam * D * d_free - g))
        Xn, Rn, tn = apply_step(X, R, t, delta, mode)
        tterms, _ = cost_terms(Xn, Rn, tn, K, x2d, conf, coef)
        Ft = sum(tterms.values())
        rho = (F - Ft) / pred if pred > 0 else 0.0
        accepted = bool(np.isfinite(Ft) and Ft < F)
        row = dict(iter=it, cost=F, trial_cost=Ft, lam=lam, rho=rho, accepted=accepted, pred=pred, n_clamped=ncl, cg_iters=cg_iters, r2=r2)
        row.update(terms)
        hist.append(row)
        if accepted:
            lam, nu = lam * max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3), 2.0
            X, R, t = Xn, Rn, tn
        else:
            lam, nu = lam * nu, 2.0 * nu
    return R, t, X, hist


def make_problem(rig: str, T: int, J: int, seed: int = 0, cam_jitter: float = 0.0):
    """The BA test problem of oracle/lm.py with the cameras broadcast over the frames (the call site's layout); cam_jitter > 0
    adds a per-frame random walk to the camera translations so that the smoothness / baseline terms are active."""
    from . import lm

    clip, R0, t0, X0 = lm.make_problem(rig, T, J, seed)
    C = len(R0)
    R = np.broadcast_to(R0, (T, C, 3, 3)).copy()
    t = np.broadcast_to(t0, (T, C, 3)).copy()
    if cam_jitter:
        rng = np.random.default_rng(seed + 7)
        t += np.cumsum(rng.normal(0.0, cam_jitter, (T, C, 3)), axis=0)
    return clip, R, t, X0
