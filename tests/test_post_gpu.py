"""GPU parity of the post-triangulation triage / smoothing kernels (row N2) against the REFERENCE's own
outputs (tests/golden/g7_post_triage.npz), the fp64 oracle on larger seeded inputs, and size-independent
properties at the full config-2 size."""
import warnings

import numpy as np
import pytest
import torch

from oracle import geometry as G
from oracle import postprocess as OP
from skiing_analysis_pytorch_b200 import dropin, post, synth

pytestmark = pytest.mark.gpu
KEYS = ["rmse_px", "median_err_px", "pos_depth_ratio", "kept_ratio", "kept_count"]
SMOOTH_RTOL = 5e-6   # fp32 window sums (9-25 terms) vs scipy's float64 arithmetic on the same float32 samples


def _cases(g):
    return {"plain": dict(), "dist": dict(dist1=g["dist"], dist2=g["dist"]),
            "conf_smooth": dict(confL=g["confL"], confR=g["confR"], smooth=True),
            "tight_smooth6": dict(err_thresh_px=1.0, smooth=True, sg_win=6, sg_poly=3)}


def test_golden_g7_through_the_shim(cuda, golden):
    pp = dropin.shim("triangulation.postprocess")
    g = golden("g7_post_triage.npz")
    X, kL, kR, K, R, t = g["X"], g["kptL"], g["kptR"], g["K"], g["R"], g["t"]
    for name, kw in _cases(g).items():
        Xc, st = pp.post_triage_sequence(X, kL, kR, K, K, R, t, **kw)
        ref = g[f"{name}_X"]
        assert Xc.dtype == np.float32 and Xc.shape == ref.shape and len(st) == 80
        np.testing.assert_array_equal(np.isnan(Xc), np.isnan(ref))          # identical accept / reject decisions
        np.testing.assert_allclose(Xc, ref, rtol=SMOOTH_RTOL if kw.get("smooth") else 0, atol=2e-6 if kw.get("smooth") else 0, equal_nan=True)
        got = np.array([[s[k] for k in KEYS] for s in st])
        np.testing.assert_allclose(got[:, 2:], g[f"{name}_stats"][:, 2:], rtol=0, atol=1e-12)   # ratios and counts: exact
        np.testing.assert_allclose(got[:, :2], g[f"{name}_stats"][:, :2], rtol=1e-6, equal_nan=True)  # rmse / median from f32 errors
        assert isinstance(st[0]["kept_count"], int)
    Xc1, rep1, keep1 = pp.post_triage_single(X[20], kL[20], kR[20], K, K, R, t, confL=g["confL"][20], confR=g["confR"][20], return_masks=True)
    np.testing.assert_array_equal(keep1, g["single_keep"])
    np.testing.assert_array_equal(np.isnan(Xc1), np.isnan(g["single_X"]))
    np.testing.assert_allclose([rep1[k] for k in KEYS], g["single_stats"], rtol=1e-6)
    assert len(pp.post_triage_single(X[0], kL[0], kR[0], K, K, R, t)) == 2
    np.testing.assert_allclose(pp.smooth_skeleton(g["smooth_in"], win=9, poly=2), g["smooth_out_9_2"], rtol=SMOOTH_RTOL, atol=2e-6, equal_nan=True)
    np.testing.assert_allclose(pp.smooth_skeleton(g["smooth_in"], win=8, poly=3), g["smooth_out_8_3"], rtol=SMOOTH_RTOL, atol=2e-6, equal_nan=True)
    # helpers
    P1, P2 = pp.build_P(K), pp.build_P(K, R, t)
    np.testing.assert_allclose(P2, G.make_P(K, R, t), atol=1e-12)
    Xf = np.nan_to_num(X[1].astype(np.float64), nan=1.0)
    e1, e2, em = pp.reproj_errors(P1, P2, Xf, kL[1], kR[1])
    np.testing.assert_allclose(e1, np.linalg.norm(OP.project(P1, Xf) - kL[1], axis=1), atol=2.5e-4)
    np.testing.assert_array_equal(pp.positive_depth_mask(R, t, X[10]), (X[10][:, 2] > 0) & ((X[10] @ R.T + t)[:, 2] > 0))


@pytest.mark.parametrize("J,with_dist,with_conf", [(17, True, True), (70, False, True), (17, True, False)])
def test_triage_matches_oracle_on_seeded_clips(cuda, J, with_dist, with_conf):
    clip = synth.make_clip("2b", 400, J, seed=31)
    K, R, t = clip.K[0], clip.R[1], clip.t[1]
    P = np.stack([G.make_P(K, np.eye(3), np.zeros(3)), G.make_P(K, R, t)])
    X = G.dlt_triangulate(P, clip.x_vm.reshape(2, -1, 2)).reshape(400, J, 3).astype(np.float32)
    rng = np.random.default_rng(3)
    X[rng.uniform(size=X.shape[:2]) < 0.02] = np.nan
    X[rng.uniform(size=X.shape[:2]) < 0.01] *= -1.0
    kw = dict(dist1=synth.DIST_CALIB, dist2=synth.DIST_CALIB) if with_dist else {}
    cf = dict(confL=clip.conf_vm[0], confR=clip.conf_vm[1]) if with_conf else {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Xo, so = OP.post_triage_sequence(X, clip.x_vm[0], clip.x_vm[1], K, K, R, t, err_thresh_px=1.5, **kw, **cf)
    res = post.post_triage(torch.from_numpy(X).to(cuda), torch.from_numpy(clip.x_vm).to(cuda), K, K, R, t,
                           kw.get("dist1"), kw.get("dist2"), conf=torch.from_numpy(clip.conf_vm).to(cuda) if with_conf else None,
                           err_thresh_px=1.5)
    Xg = res.X_clean.cpu().numpy()
    np.testing.assert_array_equal(np.isnan(Xg), np.isnan(Xo))
    np.testing.assert_array_equal(Xg[~np.isnan(Xg)], Xo[~np.isnan(Xo)])
    rep = res.report.cpu().numpy()
    ref = np.array([[s[k] for k in KEYS] for s in so])
    np.testing.assert_allclose(rep[:, 2:], ref[:, 2:], atol=1e-12)
    np.testing.assert_allclose(rep[:, :2], ref[:, :2], rtol=1e-6, equal_nan=True)
    fl = res.flags.cpu().numpy()
    assert ((fl & 8) != 0).sum() == int(ref[:, 4].sum())
    assert (((fl & 8) != 0) == (((fl & 1) != 0) & ((fl & 2) != 0) & ((fl & 4) != 0))).all()


def test_smoothing_matches_oracle_with_gaps_and_odd_sizes(cuda):
    rng = np.random.default_rng(7)
    for T, J, win, poly in ((1000, 17, 9, 2), (334, 70, 15, 3), (130, 3, 25, 4), (333, 5, 9, 2), (9, 2, 9, 2), (5, 2, 9, 2)):  # odd T -> window 3 (postprocess.py:58)
        X = synth.skeleton_clip(T, J, rng).astype(np.float32)
        X[rng.uniform(size=X.shape) < 0.1] = np.nan
        if J > 2:
            X[:, 1, 0] = np.nan
            X[: T - 3, 2, 1] = np.nan
        ref = OP.smooth_skeleton(X, win, poly)
        out = post.smooth_skeleton(torch.from_numpy(X).to(cuda), win=win, poly=poly).cpu().numpy()
        np.testing.assert_array_equal(np.isnan(out), np.isnan(ref))
        np.testing.assert_allclose(out, ref, rtol=SMOOTH_RTOL, atol=2e-6, equal_nan=True)


def test_smoothing_argument_errors(cuda):
    X = torch.zeros(40, 3, 3, device=cuda)
    with pytest.raises(ValueError, match="polyorder"):   # scipy's message for the same mistake
        post.smooth_skeleton(X, win=5, poly=5)
    with pytest.raises(RuntimeError, match="no CPU path"):
        post.smooth_skeleton(X.cpu())
    assert tuple(post.smooth_skeleton(X[:0]).shape) == (0, 3, 3)


def test_full_size_properties(cuda):
    """1M frames x 17 joints: (1) shards give bit-identical triage; (2) smoothing a polynomial of degree <= poly
    is the identity (the defining property of the filter), gaps included; (3) smoothing is linear."""
    T, J = 1_000_000, 17
    d = synth.make_clip_device("2b", T, J, cuda, seed=5, layout="CTJ2")
    K, R, t = d["K"][0], d["R"][1], d["t"][1]
    X = d["X"].float()
    res = post.post_triage(X, d["x2d"], K, K, R, t, synth.DIST_CALIB, synth.DIST_CALIB, conf=d["conf"], conf_thr=0.3)
    for a, b in ((0, 777), (500_000, 500_321), (999_000, 1_000_000)):
        part = post.post_triage(X[a:b].contiguous(), d["x2d"][:, a:b].contiguous(), K, K, R, t, synth.DIST_CALIB, synth.DIST_CALIB,
                                conf=d["conf"][:, a:b].contiguous(), conf_thr=0.3)
        assert torch.equal(part.flags, res.flags[a:b])
        assert torch.equal(torch.nan_to_num(part.X_clean, nan=-1.0), torch.nan_to_num(res.X_clean[a:b], nan=-1.0))
    kept = ((res.flags & 8) != 0)
    assert abs(((res.flags & 4) != 0).float().mean().item() - 0.875 ** 2) < 0.01   # both confidences U(0.2,1) >= 0.3
    assert ((res.flags & 1) != 0).all()                                           # ground-truth points are in front of both cameras
    assert 0.6 < kept.float().mean().item() < 0.766                               # minus the 2 px error gate under 1 px noise
    assert res.report[:, 4].sum().item() == kept.sum().item()
    tt = torch.arange(T, device=cuda, dtype=torch.float64) / T
    poly = torch.stack([1.0 + 2.0 * tt - 3.0 * tt * tt, 0.5 - tt, 4.0 * tt * tt], -1)[:, None, :].expand(T, J, 3).float().contiguous()
    gaps = poly.clone()
    gaps[torch.rand(T, J, device=cuda, generator=torch.Generator(device=cuda).manual_seed(1)) < 0.2] = float("nan")
    # with gaps the compacted abscissa is no longer uniform, so only the gap-free clip reproduces the polynomial
    sm = post.smooth_skeleton(poly, win=9, poly=2)
    assert (sm - poly).abs().max().item() < 2e-5
    sg = post.smooth_skeleton(gaps, win=9, poly=2)
    assert torch.equal(torch.isnan(sg), torch.isnan(gaps))
    a = post.smooth_skeleton(X, win=11, poly=3)
    b = post.smooth_skeleton(2.0 * X, win=11, poly=3)
    assert (b - 2.0 * a).abs().max().item() < 1e-4
