"""CPU checks of the calibrating-BA oracle (oracle/lm_calib.py: free intrinsics + distortion, BASELINE config 3's
"Rodrigues extrinsics + intrinsics/distortion"): its projection and its [tvec, f, c, dist] Jacobian columns are
cv2.projectPoints' own (the model the reference reprojects with, triangulation/reproject.py:77-78), every Jacobian
column equals finite differences, the Schur algebra equals a dense normal-equation solve (prior rows included), it
reduces to oracle/lm.py (whose cost is pinned by the reference's reprojection_loss) when only extrinsics are free, and
it does not depend on the shard count."""
import numpy as np
import pytest

from oracle import geometry as G
from oracle import lm, lm_calib as lc


def _flat(clip):
    C = len(clip.R)
    x = clip.x_fm.astype(float).transpose(0, 2, 1, 3).reshape(-1, C, 2)
    cf = clip.conf_fm.astype(float).transpose(0, 2, 1).reshape(-1, C)
    return x, cf / (cf.sum() + 1e-6)


def test_projection_and_jacobian_are_cv2s():
    cv2 = pytest.importorskip("cv2")
    clip, R0, t0, th, X0 = lc.make_problem("2b", 3, 6)
    x, _ = _flat(clip)
    X = X0.reshape(-1, 3)
    uv, _ = lc.project(X, R0, t0, th)
    e, A, B, _ = lc.residual_blocks(X, R0, t0, th, x)
    for c in range(2):
        K = np.array([[th[c, 0], 0, th[c, 2]], [0, th[c, 1], th[c, 3]], [0, 0, 1.0]])
        rvec = cv2.Rodrigues(R0[c])[0]
        p, jac = cv2.projectPoints(X.reshape(-1, 1, 3), rvec, t0[c].reshape(3, 1), K, th[c, 4:].copy())
        np.testing.assert_allclose(uv[:, c], p.reshape(-1, 2), rtol=0, atol=1e-9)
        jac = jac.reshape(-1, 2, jac.shape[-1])  # (N, 2, 3 rvec + 3 tvec + 2 f + 2 c + 5 dist)
        np.testing.assert_allclose(B[:, c, :, 3:15], jac[:, :, 3:15], rtol=1e-9, atol=1e-9)


def test_jacobians_match_finite_differences():
    clip, R0, t0, th, X0 = lc.make_problem("2b", 3, 4)
    x, _ = _flat(clip)
    X = X0.reshape(-1, 3)
    e, A, B, _ = lc.residual_blocks(X, R0, t0, th, x)
    h = 1e-6
    for k in range(3):
        Xp = X.copy()
        Xp[:, k] += h
        ep = lc.residual_blocks(Xp, R0, t0, th, x)[0]
        np.testing.assert_allclose((ep - e) / h, A[..., k], rtol=1e-4, atol=1e-3)
    for c in range(2):
        for k in range(lc.P):
            d = np.zeros((2, lc.P))
            hk = h * (1.0 if k < 6 or k >= 10 else 1e3)
            d[c, k] = hk
            Rp, tp, thp = lc.apply_camera_step(R0, t0, th, d)
            ep = lc.residual_blocks(X, Rp, tp, thp, x)[0]
            np.testing.assert_allclose((ep - e)[:, c] / hk, B[:, c, :, k], rtol=2e-4, atol=2e-2)
            assert np.abs((ep - e)[:, 1 - c]).max() == 0.0


@pytest.mark.parametrize("rho_scale", [0.0, 1.0])
def test_schur_equals_dense_normal_equations(rho_scale):
    clip, R0, t0, th, X0 = lc.make_problem("3", 5, 6)
    x, w = _flat(clip)
    X = X0.reshape(-1, 3)
    th0 = lc.intr_from_K(clip.K)
    rho = np.broadcast_to(lc.PRIOR_RHO * rho_scale, (3, lc.NI))
    for mode in ("full", "extr_focal", "intr_only"):
        free = lc.free_mask(3, mode)
        for lam in (1e-3, 0.5):
            dc, dp = lc.dense_step(X, R0, t0, th, x, w, lam, free, th0, rho)
            lin = lc.linearise(X, R0, t0, th, x, w, lam)
            dc2, _, ok = lc.solve_reduced(lin, lam, free, th, th0, rho)
            dp2, _ = lc.back_substitute(X, R0, t0, th, x, w, lam, dc2)
            assert ok
            np.testing.assert_allclose(dc2, dc, rtol=1e-6, atol=1e-7 * max(1.0, np.abs(dc).max()))
            np.testing.assert_allclose(dp2, dp, rtol=1e-6, atol=1e-7 * max(1.0, np.abs(dp).max()))


def test_reduces_to_the_pinned_extrinsics_only_oracle():
    """Zero distortion + only extrinsics free == oracle/lm.py, whose cost is the reference's reprojection_loss."""
    clip, R0, t0, X0 = lm.make_problem("2b", 40, 17)
    base = lm.run_lm(X0, R0, t0, clip.K, clip.x_fm, clip.conf_fm, num_iters=8)
    free = np.zeros((2, lc.P), bool)
    free[1, :6] = True
    R, t, th, X, hist = lc.run_lm(X0, R0, t0, lc.intr_from_K(clip.K), clip.x_fm, clip.conf_fm, num_iters=8, free=free)
    for a, b in zip(base[3], hist):
        assert a["accepted"] == b["accepted"]
        assert abs(a["cost"] - b["cost"]) <= 1e-9 * a["cost"]
        assert abs(a["trial_cost"] - b["trial_cost"]) <= 1e-9 * a["trial_cost"]
    np.testing.assert_allclose(R, base[0], atol=1e-9)
    np.testing.assert_allclose(X, base[2], atol=1e-7)


@pytest.mark.parametrize("rig,T,J", [("2b", 60, 17), ("3", 30, 17)])
def test_lm_removes_the_intrinsic_error(rig, T, J):
    clip, R0, t0, th, X0 = lc.make_problem(rig, T, J)
    th_gt = lc.intr_from_K(clip.K)
    R, t, th1, X, hist = lc.run_lm(X0, R0, t0, th, clip.x_fm, clip.conf_fm, num_iters=25, prior_theta=th_gt,
                                   prior_rho=lc.PRIOR_RHO)
    final = min(h["trial_cost"] if h["accepted"] else h["cost"] for h in hist)
    assert final < 0.05 * hist[0]["cost"]
    costs = [h["cost"] for h in hist]
    assert all(b <= a for a, b in zip(costs, costs[1:]))
    # the extrinsics-only optimum with the WRONG intrinsics is worse than the calibrating one
    Rb, tb, Xb, hb = lm.run_lm(X0, R0, t0, np.stack([np.array([[q[0], 0, q[2]], [0, q[1], q[3]], [0, 0, 1.0]]) for q in th]),
                               clip.x_fm, clip.conf_fm, num_iters=25)
    assert final < min(h["trial_cost"] if h["accepted"] else h["cost"] for h in hb)


def test_shard_invariance():
    clip, R0, t0, th, X0 = lc.make_problem("2b", 16, 10)
    kw = dict(num_iters=5, prior_theta=lc.intr_from_K(clip.K), prior_rho=lc.PRIOR_RHO)
    base = lc.run_lm(X0, R0, t0, th, clip.x_fm, clip.conf_fm, **kw)[4]
    for shards in (2, 4):
        h = lc.run_lm(X0, R0, t0, th, clip.x_fm, clip.conf_fm, shards=shards, **kw)[4]
        for a, b in zip(base, h):
            assert abs(a["trial_cost"] - b["trial_cost"]) <= 1e-10 * a["trial_cost"]
            assert a["accepted"] == b["accepted"]
