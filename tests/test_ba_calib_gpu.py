"""GPU parity of the CALIBRATING bundle adjustment (free intrinsics + distortion, BASELINE config 3's
"Rodrigues extrinsics + intrinsics/distortion") through the C ABI against the fp64 oracle oracle/lm_calib.py.

Tolerances: LM cost trajectory 1e-4 relative per trial (north star); the packed reduced system of one linearisation
entry by entry at 3e-5 of its block's scale (fp32 per-point arithmetic, fp64 reductions)."""
import numpy as np
import pytest
import torch

from oracle import lm, lm_calib as lc
from skiing_analysis_pytorch_b200 import _cabi, ba

pytestmark = pytest.mark.gpu

COST_TOL = 1e-4


def _dev(clip, X0, cuda, layout="TCJ2"):
    x, c = (clip.x_fm, clip.conf_fm) if layout == "TCJ2" else (clip.x_vm, clip.conf_vm)
    return (torch.from_numpy(np.ascontiguousarray(x)).to(cuda), torch.from_numpy(np.ascontiguousarray(c)).to(cuda),
            torch.from_numpy(X0).to(cuda))


def _K_dist(th):
    K = np.zeros((len(th), 3, 3))
    K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2], K[:, 2, 2] = th[:, 0], th[:, 1], th[:, 2], th[:, 3], 1.0
    return K, th[:, 4:].copy()


def _unpack(red, C):
    """GPU payload -> the oracle's (Hcc (C,15,15), gc (C,15), bw (15C,), Sw (15C,15C), cost, n_clamped)."""
    L = _cabi.calib_red_layout(C)
    n, P = L["n"], _cabi.BA_CALIB_PARAMS
    S = np.zeros((n, n))
    iu = np.triu_indices(n)
    S[iu] = red[: n * (n + 1) // 2]
    S = S + np.triu(S, 1).T
    cols = [(c, r) for c in range(C) for r in range(P) if not (c == 0 and r < 6)]
    idx = np.array([P * c + r for c, r in cols])
    assert [_cabi.calib_col(c, r) for c, r in cols] == list(range(n))
    Sw = np.zeros((P * C, P * C))
    Sw[np.ix_(idx, idx)] = S
    Hcc, gc, bw = np.zeros((C, P, P)), np.zeros((C, P)), np.zeros(P * C)
    cost = ncl = 0.0
    for c in range(C):
        blk = red[L["cam"] + _cabi.BA_CALIB_CAM_BLOCK * c: L["cam"] + _cabi.BA_CALIB_CAM_BLOCK * (c + 1)]
        for r in range(P):
            for s in range(r, P):
                Hcc[c, r, s] = Hcc[c, s, r] = blk[_cabi.calib_tri(r, s)]
            gc[c, r] = blk[_cabi.calib_tri(r, 15)]
            bw[P * c + r] = blk[_cabi.calib_tri(r, 16)]
        cost += blk[_cabi.calib_tri(15, 15)]
        ncl += blk[153]
        assert np.all(blk[154:] == 0.0)
    return Hcc, gc, bw, Sw, cost, ncl


@pytest.mark.parametrize("T,J,layout", [(200, 17, "TCJ2"), (77, 17, "CTJ2"), (13, 5, "TCJ2"), (1000, 17, "TCJ2")])
def test_linearisation_matches_oracle(cuda, T, J, layout):
    clip, R0, t0, th, X0 = lc.make_problem("2b", T, J)
    K, dist = _K_dist(th)
    x, c, X = _dev(clip, X0, cuda, layout)
    s = ba.CalibratingBundleAdjuster(x, c, K, R0, t0, X, dist=dist, layout=layout)
    s.linearize()
    torch.cuda.synchronize()
    Hcc, gc, bw, Sw, cost, ncl = _unpack(s.red.cpu().numpy(), 2)
    xo = clip.x_fm.astype(float).transpose(0, 2, 1, 3).reshape(T * J, 2, 2)
    wo = clip.conf_fm.astype(float).transpose(0, 2, 1).reshape(T * J, 2)
    lin = lc.linearise(X0.astype(np.float32).astype(float).reshape(-1, 3), R0, t0, th, xo, wo, 1e-3)
    assert abs(cost - lin.cost) <= 3e-5 * lin.cost
    assert ncl == lin.n_clamped == 0
    # per parameter pair the natural scale is sqrt(H_rr H_ss): the blocks mix px/px (focal) and px/unit (distortion) columns
    for cam in range(2):
        sc = np.sqrt(np.diag(lin.Hcc[cam]))
        np.testing.assert_allclose(Hcc[cam] / np.outer(sc, sc), lin.Hcc[cam] / np.outer(sc, sc), rtol=0, atol=3e-5)
        np.testing.assert_allclose(gc[cam] / sc, lin.gc[cam] / sc, rtol=0, atol=3e-5 * np.abs(lin.gc[cam] / sc).max())
    free = np.ones(30, bool)
    free[:6] = False
    sc = np.sqrt(np.concatenate([np.diag(lin.Hcc[0]), np.diag(lin.Hcc[1])]))
    ref = lin.Sw / np.outer(sc, sc)
    got = Sw / np.outer(sc, sc)
    np.testing.assert_allclose(got[np.ix_(free, free)], ref[np.ix_(free, free)], rtol=0, atol=3e-5)
    np.testing.assert_allclose((bw / sc)[free], (lin.bw / sc)[free], rtol=0,
                               atol=3e-5 * max(np.abs(lin.bw / sc).max(), np.abs(lin.gc.reshape(-1) / sc).max()))


def _check_history(hist, ref, n):
    for k in range(n):
        g, o = hist[k], ref[k]
        assert abs(g["cost"] - o["cost"]) <= COST_TOL * o["cost"], (k, g, o)
        assert abs(g["trial_cost"] - o["trial_cost"]) <= COST_TOL * o["trial_cost"], (k, g, o)
        assert g["n_clamped"] == o["n_clamped"]


@pytest.mark.parametrize("calib,prior", [("full", True), ("extr_focal", True), ("intr_only", True), ("full", False)])
def test_lm_trajectory_matches_oracle(cuda, calib, prior):
    T, J = 200, 17
    clip, R0, t0, th, X0 = lc.make_problem("2b", T, J)
    K, dist = _K_dist(th)
    th_gt = lc.intr_from_K(clip.K)
    x, c, X = _dev(clip, X0, cuda)
    kw = dict(prior_rho=lc.PRIOR_RHO, prior_theta=th_gt) if prior else {}
    n_it = 10 if prior else 4  # without the prior the problem has near-null directions: compare the decisive trials only
    s = ba.ba_calibrate(x, c, K, R0, t0, X, num_iters=n_it, dist=dist, calib=calib, **kw)
    R, t, th1, Xo, ref = lc.run_lm(X0, R0, t0, th, clip.x_fm, clip.conf_fm, num_iters=n_it, free=lc.free_mask(2, calib),
                                   prior_theta=th_gt if prior else None, prior_rho=lc.PRIOR_RHO if prior else None)
    hist = s.history
    _check_history(hist, ref, n_it)
    checked = 0
    for k in range(n_it):
        o = ref[k]
        if (o["cost"] - o["trial_cost"]) <= 1e-3 * o["cost"]:
            break
        assert hist[k]["accepted"] == bool(o["accepted"])
        assert abs(hist[k]["lam"] - o["lam"]) <= 1e-2 * o["lam"]
        assert abs(hist[k]["rho"] - o["rho"]) <= 5e-3
        checked += 1
    assert checked >= 2
    if prior:
        final = min(h["trial_cost"] if h["accepted"] else h["cost"] for h in ref)
        assert abs(s.cost - final) <= COST_TOL * final
        # the prior pins the intrinsics: parameters agree with the oracle's, and the initial 1 % focal error is gone
        np.testing.assert_allclose(s.theta[:, :4], th1[:, :4], atol=0.05)
        np.testing.assert_allclose(s.theta[:, 4:], th1[:, 4:], atol=2e-3)
        if calib == "full":
            assert np.abs(s.theta[:, :2] - th_gt[:, :2]).max() < 1.0 < np.abs(th[:, :2] - th_gt[:, :2]).min()
        np.testing.assert_allclose(s.R, R, atol=2e-4)
        np.testing.assert_allclose(s.t, t, atol=2e-3)


def test_extrinsics_only_mask_equals_the_six_parameter_engine(cuda):
    """Zero distortion + only camera 1's extrinsics free: the calibrating kernels must walk the same trajectory as the
    6-parameter engine (whose cost is the reference's reprojection_loss)."""
    clip, R0, t0, X0 = lm.make_problem("2b", 300, 17)
    x, c, X = _dev(clip, X0, cuda)
    a = ba.ba_solve(x, c, clip.K, R0, t0, X, num_iters=8)
    b = ba.ba_calibrate(x, c, clip.K, R0, t0, X, num_iters=8, calib=0x3F << 15)
    for ha, hb in zip(a.history, b.history):
        assert abs(ha["cost"] - hb["cost"]) <= 1e-5 * ha["cost"]
        assert abs(ha["trial_cost"] - hb["trial_cost"]) <= 1e-5 * ha["trial_cost"]
    np.testing.assert_allclose(a.R, b.R, atol=2e-4)  # cost-equivalent end states: the free scale is a nearly flat direction
    np.testing.assert_array_equal(b.theta, lc.intr_from_K(clip.K))


def test_graph_replay_equals_eager_and_deterministic(cuda):
    clip, R0, t0, th, X0 = lc.make_problem("2b", 300, 17)
    K, dist = _K_dist(th)
    x, c, X = _dev(clip, X0, cuda)
    kw = dict(dist=dist, prior_rho=lc.PRIOR_RHO, prior_theta=lc.intr_from_K(clip.K))
    a = ba.ba_calibrate(x, c, K, R0, t0, X, num_iters=8, **kw)
    b = ba.ba_calibrate(x, c, K, R0, t0, X, num_iters=8, graph=True, **kw)
    torch.cuda.synchronize()
    assert a.history == b.history
    assert torch.equal(a.X, b.X)
    np.testing.assert_array_equal(a.theta, b.theta)


def test_unobserved_and_ragged(cuda):
    """Zero-confidence points take no step; a clip that is not a multiple of the 384-point tile is handled."""
    clip, R0, t0, th, X0 = lc.make_problem("2b", 41, 17)
    K, dist = _K_dist(th)
    conf = clip.conf_fm.copy()
    conf[7, :, 2] = 0.0
    x = torch.from_numpy(clip.x_fm).to(cuda)
    s = ba.CalibratingBundleAdjuster(x, torch.from_numpy(conf).to(cuda), K, R0, t0, torch.from_numpy(X0).to(cuda), dist=dist,
                                     prior_rho=lc.PRIOR_RHO)
    s.run(4)
    h = s.history
    assert all(np.isfinite(r["trial_cost"]) for r in h)
    np.testing.assert_array_equal(s.X.cpu().numpy()[7, 2], X0[7, 2].astype(np.float32))
    ref = lc.run_lm(X0, R0, t0, th, clip.x_fm, conf, num_iters=4, prior_theta=th, prior_rho=lc.PRIOR_RHO)[4]
    _check_history(h, ref, 4)


def test_clamped_points_are_counted_and_harmless(cuda):
    """A point behind camera 1 hits loss.py's depth clamp: counted per camera in slot 153 of the camera blocks, zero
    z-derivative, and the first trial's costs still match the oracle."""
    clip, R0, t0, th, X0 = lc.make_problem("2b", 40, 17)
    K, dist = _K_dist(th)
    X0 = X0.copy()
    X0[3, 5] = [0.0, 0.0, -5.0]
    X0[9, 2] = [30.0, 0.0, 10.0]
    x, c, X = _dev(clip, X0, cuda)
    s = ba.CalibratingBundleAdjuster(x, c, K, R0, t0, X, dist=dist, prior_rho=lc.PRIOR_RHO)
    s.run(2)
    ref = lc.run_lm(X0.astype(np.float32).astype(float), R0, t0, th, clip.x_fm, clip.conf_fm, num_iters=2, prior_theta=th,
                    prior_rho=lc.PRIOR_RHO)[4]
    h = s.history
    assert h[0]["n_clamped"] == ref[0]["n_clamped"] >= 2
    assert abs(h[0]["cost"] - ref[0]["cost"]) <= 1e-4 * ref[0]["cost"]
    assert all(np.isfinite(r["trial_cost"]) for r in h)


def test_unsupported_camera_count(cuda):
    clip, R0, t0, X0 = lm.make_problem("3", 20, 17)
    x, c, X = _dev(clip, X0, cuda)
    with pytest.raises(ValueError):
        ba.CalibratingBundleAdjuster(x, c, clip.K, R0, t0, X)


def test_full_size_config3_properties(cuda):
    """BASELINE config 3 (100k frames x 17 joints x 2 cameras) with free intrinsics / distortion: shard linearity of the
    packed payload, monotone accepted costs, the intrinsic error is removed, and a sub-sampled oracle agrees on the cost."""
    from skiing_analysis_pytorch_b200 import api, synth

    T, J = 100_000, 17
    clip = synth.make_clip("2b", T, J, seed=0)
    R0, t0 = synth.perturb_cameras(clip.R, clip.t, seed=1)
    th = lc.perturb_intrinsics(clip.K)
    th_gt = lc.intr_from_K(clip.K)
    K, dist = _K_dist(th)
    x = torch.from_numpy(clip.x_fm).to(cuda)
    c = torch.from_numpy(clip.conf_fm).to(cuda)
    X0 = api.triangulate_reproject(torch.from_numpy(clip.x_vm).to(cuda), K, R0, t0, want=("X",)).X
    kw = dict(dist=dist, prior_rho=lc.PRIOR_RHO, prior_theta=th_gt)
    full = ba.CalibratingBundleAdjuster(x, c, K, R0, t0, X0, **kw)
    full.linearize()
    halves = []
    for a, b in ((0, 37_001), (37_001, T)):
        h = ba.CalibratingBundleAdjuster(x[a:b].contiguous(), c[a:b].contiguous(), K, R0, t0, X0[a:b].contiguous(), **kw)
        h.linearize()
        halves.append(h.red.clone())
    tot = halves[0] + halves[1]
    rel = ((tot - full.red).abs() / full.red.abs().clamp_min(1e-30))[full.red.abs() > 1e-6 * full.red.abs().max()]
    assert rel.max().item() < 1e-5
    full.run(14)
    h = full.history
    acc = [r for r in h if r["accepted"]]
    assert len(acc) >= 6 and all(r["trial_cost"] < r["cost"] for r in acc)
    assert h[0]["cost"] > 20.0 and full.cost < 0.5
    assert np.abs(full.theta[:, :2] - th_gt[:, :2]).max() < 1.0
    sub = slice(0, T, 50)
    xs = clip.x_fm[sub].astype(float).transpose(0, 2, 1, 3).reshape(-1, 2, 2)
    ws = clip.conf_fm[sub].astype(float).transpose(0, 2, 1).reshape(-1, 2)
    Xs = full.X[sub].cpu().numpy().astype(float).reshape(-1, 3)
    cs, _ = lc.cost_only(Xs, full.R, full.t, full.theta, xs, ws / (ws.sum() + 1e-6))
    data_cost = full.cost - lc.prior_cost(full.theta, th_gt, np.broadcast_to(lc.PRIOR_RHO, (2, 9)))
    assert abs(cs - data_cost) < 0.03 * data_cost
