"""CPU checks of the fp64 Schur-LM oracle (oracle/lm.py): its cost IS the reference's
reprojection_loss (golden G3/G6, values produced by bundle_adjustment/loss.py itself), the Schur
algebra equals a dense normal-equation solve, the analytic Jacobians equal finite differences, the
result does not depend on the shard count, the golden history is reproduced, and the optimum agrees
with scipy's least_squares (the only second-order BA code in the reference uses it:
VideoPose3D/slove_rt_from_3d.py:140-244)."""
import numpy as np
import pytest

from oracle import geometry as G
from oracle import lm


def _flat(clip):
    C = len(clip.R)
    x = clip.x_fm.astype(float).transpose(0, 2, 1, 3).reshape(-1, C, 2)
    cf = clip.conf_fm.astype(float).transpose(0, 2, 1).reshape(-1, C)
    return x, cf / (cf.sum() + 1e-6)


def test_cost_is_reference_reprojection_loss(golden):
    g = golden("g3_g4_loss.npz")
    X, R, t, K, x2d, conf = g["X"], g["R_c"], g["t_c"], g["K_c"], g["x2d"], g["conf"]
    T, J, _ = X.shape
    C = R.shape[0]
    x = x2d.astype(float).transpose(0, 2, 1, 3).reshape(T * J, C, 2)
    cf = conf.astype(float).transpose(0, 2, 1).reshape(T * J, C)
    w = cf / (cf.sum() + 1e-6)
    c, _ = lm.cost_only(X.reshape(-1, 3), R, t, K, x, w)
    ref = float(g["loss_static_f64"])  # bundle_adjustment/loss.py:90-94 run in the authoring container
    assert abs(c - ref) <= 1e-12 * ref
    lin = lm.linearise(X.reshape(-1, 3), R, t, K, x, w, 1e-3)
    assert abs(lin.cost - ref) <= 1e-12 * ref


def test_golden_history_and_reference_cost(golden):
    g = golden("g6_lm_history.npz")
    for name, (rig, T, J, mode) in lm.G6_CASES.items():
        clip, R0, t0, X0 = lm.make_problem(rig, T, J)
        R, t, X, hist = lm.run_lm(X0, R0, t0, clip.K, clip.x_fm, clip.conf_fm, num_iters=10, mode=mode)
        gh = g[f"{name}_hist"]
        # the reference's own reprojection_loss at the initial and the final state
        assert abs(hist[0]["cost"] - float(g[f"{name}_ref_loss_init"])) <= 1e-11 * hist[0]["cost"]
        final = hist[-1]["trial_cost"] if hist[-1]["accepted"] else hist[-1]["cost"]
        assert abs(final - float(g[f"{name}_ref_loss_final"])) <= 1e-9 * final
        np.testing.assert_allclose([h["cost"] for h in hist], gh[:, 1], rtol=1e-9)
        np.testing.assert_allclose([h["trial_cost"] for h in hist[:6]], gh[:6, 2], rtol=1e-9)
        np.testing.assert_allclose(R, g[f"{name}_R"], atol=1e-7)
        np.testing.assert_allclose(t, g[f"{name}_t"], atol=1e-6)
        assert hist[0]["cost"] > 10 * final  # it actually optimises


def test_schur_equals_dense_normal_equations():
    clip, R0, t0, X0 = lm.make_problem("3", 5, 4)
    x, w = _flat(clip)
    X = X0.reshape(-1, 3)
    for mode in ("full", "pose_cam_t"):
        free = lm.free_mask(3, mode)
        for lam in (1e-3, 0.5):
            dc, dp = lm.dense_step(X, R0, t0, clip.K, x, w, lam, free)
            lin = lm.linearise(X, R0, t0, clip.K, x, w, lam)
            dc2, _, ok = lm.solve_reduced(lin, lam, free)
            dp2, _ = lm.back_substitute(X, R0, t0, clip.K, x, w, lam, dc2)
            assert ok
            np.testing.assert_allclose(dc2, dc, atol=1e-8 * max(1.0, np.abs(dc).max()))
            np.testing.assert_allclose(dp2, dp, atol=1e-8 * max(1.0, np.abs(dp).max()))


def test_jacobians_match_finite_differences():
    clip, R0, t0, X0 = lm.make_problem("2b", 3, 4)
    x, w = _flat(clip)
    X = X0.reshape(-1, 3)
    e, A, B, _ = lm.residual_blocks(X, R0, t0, clip.K, x, w)
    h = 1e-6
    for k in range(3):
        Xp = X.copy()
        Xp[:, k] += h
        ep = lm.residual_blocks(Xp, R0, t0, clip.K, x, w)[0]
        np.testing.assert_allclose((ep - e) / h, A[..., k], rtol=1e-4, atol=1e-3)
    for k in range(6):
        d = np.zeros((2, 6))
        d[1, k] = h
        Rp, tp = lm.apply_camera_step(R0, t0, d)
        ep = lm.residual_blocks(X, Rp, tp, clip.K, x, w)[0]
        np.testing.assert_allclose((ep - e)[:, 1] / h, B[:, 1, :, k], rtol=1e-4, atol=2e-2)
        assert np.abs((ep - e)[:, 0]).max() == 0.0


def test_gradient_matches_autograd_of_loss():
    torch = pytest.importorskip("torch")
    clip, R0, t0, X0 = lm.make_problem("2b", 4, 5)
    x, w = _flat(clip)
    X = torch.tensor(X0, dtype=torch.float64, requires_grad=True)
    # loss.py:17-94 restated in torch for autograd (the reference module itself does not travel)
    Rt, tt, Kt = (torch.tensor(a, dtype=torch.float64) for a in (R0, t0, clip.K))
    Xc = torch.einsum("cab,tjb->tcja", Rt, X) + tt[None, :, None, :]
    Z = Xc[..., 2].clamp(min=1e-6)
    xy = Xc[..., :2] / Z[..., None]
    u = Kt[None, :, None, 0, 0] * xy[..., 0] + Kt[None, :, None, 0, 1] * xy[..., 1] + Kt[None, :, None, 0, 2]
    v = Kt[None, :, None, 1, 0] * xy[..., 0] + Kt[None, :, None, 1, 1] * xy[..., 1] + Kt[None, :, None, 1, 2]
    cf = torch.tensor(clip.conf_fm, dtype=torch.float64)
    d = (u - torch.tensor(clip.x_fm[..., 0], dtype=torch.float64)) ** 2 + (v - torch.tensor(clip.x_fm[..., 1], dtype=torch.float64)) ** 2
    loss = (cf * d).sum() / (cf.sum() + 1e-6)
    loss.backward()
    lin = lm.linearise(X0.reshape(-1, 3), R0, t0, clip.K, x, w, 0.0, keep_points=True)
    np.testing.assert_allclose(2.0 * lin.gp.reshape(X0.shape), X.grad.numpy(), rtol=1e-9, atol=1e-12)
    assert abs(lin.cost - loss.item()) < 1e-12 * loss.item()


def test_shard_invariance():
    clip, R0, t0, X0 = lm.make_problem("8", 16, 10)
    base = lm.run_lm(X0, R0, t0, clip.K, clip.x_fm, clip.conf_fm, num_iters=5)[3]
    for shards in (2, 4, 8):
        h = lm.run_lm(X0, R0, t0, clip.K, clip.x_fm, clip.conf_fm, num_iters=5, shards=shards)[3]
        for a, b in zip(base, h):
            assert abs(a["trial_cost"] - b["trial_cost"]) <= 1e-10 * a["trial_cost"]
            assert a["accepted"] == b["accepted"]


def test_optimum_agrees_with_scipy_least_squares():
    opt = pytest.importorskip("scipy.optimize")
    clip, R0, t0, X0 = lm.make_problem("2b", 12, 6)
    x, w = _flat(clip)
    R, t, X, hist = lm.run_lm(X0, R0, t0, clip.K, clip.x_fm, clip.conf_fm, num_iters=45)  # the free-scale direction is nearly flat: slow tail
    n = X0.size

    def fun(p):
        Rc = np.stack([R0[0], G.so3_exp(p[:3]) @ R0[1]])
        tc = np.stack([t0[0], t0[1] + p[3:6]])
        e = lm.residual_blocks(X0.reshape(-1, 3) + p[6:].reshape(-1, 3), Rc, tc, clip.K, x, w)[0]
        return (e * np.sqrt(w)[..., None]).ravel()

    sol = opt.least_squares(fun, np.zeros(6 + n), method="trf", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=200)
    final = min(h["trial_cost"] if h["accepted"] else h["cost"] for h in hist)
    assert abs(2.0 * sol.cost - final) <= 1e-8 * final


def test_unobserved_points_have_no_step_and_no_schur_term():
    """A point with all confidences 0 has a singular block: the specification (shared with the CUDA kernels) is
    'no step, no contribution' - the solve stays finite and the point stays where it was."""
    clip, R0, t0, X0 = lm.make_problem("3", 20, 17)
    conf = clip.conf_fm.copy()
    conf[4, :, 2] = 0.0
    conf[9] = 0.0
    R, t, X, hist = lm.run_lm(X0, R0, t0, clip.K, clip.x_fm, conf, num_iters=5)
    assert np.isfinite(X).all() and np.isfinite(R).all() and all(np.isfinite(h["trial_cost"]) for h in hist)
    np.testing.assert_array_equal(X[4, 2], X0[4, 2])
    np.testing.assert_array_equal(X[9], X0[9])
    assert hist[-1]["cost"] < 0.1 * hist[0]["cost"]
