"""GPU parity of the two-view 3D fusion + EMA kernels (row N3, through the C ABI) against golden G8 - outputs of the
reference's own fuse/main_raw.py, fuse/confidence.py, fuse/fuse.py - and the fp64 oracle (oracle/fusion.py).

Tolerance: fp64 arithmetic on both sides, different summation orders and a Jacobi instead of a LAPACK SVD: 1e-9
relative on positions / qualities (held: ~1e-13); the sequential EMA mode is compared at 1e-14."""
import numpy as np
import pytest
import torch

from oracle import fusion as F
from skiing_analysis_pytorch_b200 import _cabi, fusion, synth

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _dev(cuda, *arrs):
    return [torch.from_numpy(np.ascontiguousarray(a)).to(cuda) for a in arrs]


def test_fuse_clip_matches_reference_golden(cuda, golden):
    g = golden("g8_fusion.npz")
    r = fusion.fuse_clip(*_dev(cuda, g["Xl"], g["Xr"], g["Ul"], g["Ur"]))
    np.testing.assert_allclose(r.aligned.cpu().numpy(), g["aligned"], rtol=TOL, atol=TOL, equal_nan=True)
    np.testing.assert_allclose(r.q_l.cpu().numpy(), g["q_l"], rtol=TOL, atol=TOL)
    np.testing.assert_allclose(r.q_r.cpu().numpy(), g["q_r"], rtol=TOL, atol=TOL)
    np.testing.assert_allclose(r.fused.cpu().numpy(), g["fused"], rtol=TOL, atol=TOL, equal_nan=True)
    assert int(r.status.sum()) == 0


def test_unity_pipeline_matches_reference_golden(cuda, golden):
    """fuse/main_unity.py: no rigid alignment, 15 target joints, the Unity joint ids select the EMA's alpha classes."""
    g = golden("g8_fusion.npz")
    r = fusion.fuse_clip(*_dev(cuda, g["unity_Xl"], g["unity_Xr"], g["unity_Ul"], g["unity_Ur"]), align=False)
    np.testing.assert_allclose(r.fused.cpu().numpy(), g["unity_fused"], rtol=TOL, atol=TOL, equal_nan=True)
    assert int(r.status.sum()) == 0
    Y = fusion.temporal_smooth_ema(r.fused, [int(i) for i in g["unity_ids"]], exact=True).cpu().numpy()
    np.testing.assert_allclose(Y, g["unity_smooth"], rtol=TOL, atol=TOL, equal_nan=True)


@pytest.mark.parametrize("J,scale_mode", [(70, "hip"), (70, "torso"), (17, "hip"), (33, "hip"), (96, "torso")])
def test_fuse_clip_matches_oracle(cuda, J, scale_mode):
    d = synth.make_fusion_clip(257, J, seed=J, nan_frac=0.08)
    r = fusion.fuse_clip(*_dev(cuda, d["Xl"], d["Xr"], d["Ul"], d["Ur"]), scale_mode=scale_mode, sigma_px=9.0, sigma_3d=0.1)
    fused, ql, qr, Xa = F.fuse_clip(d["Xl"], d["Xr"], d["Ul"], d["Ur"], sigma_px=9.0, sigma_3d=0.1, scale_mode=scale_mode)
    np.testing.assert_allclose(r.fused.cpu().numpy(), fused, rtol=TOL, atol=TOL, equal_nan=True)
    np.testing.assert_allclose(r.q_l.cpu().numpy(), ql, rtol=TOL, atol=TOL)
    np.testing.assert_allclose(r.q_r.cpu().numpy(), qr, rtol=TOL, atol=TOL)
    al = r.aligned.cpu().numpy()
    ok = np.isfinite(Xa).all(-1)
    np.testing.assert_allclose(al[ok], Xa[ok], rtol=TOL, atol=TOL)
    assert np.isnan(al[~ok]).all()
    assert (ql == 0).any() and (ql > 0.5).any()  # frames with a missing key joint have zero confidence; good joints are trusted


def test_polar_fast_path_equals_jacobi_path_and_reflections_fall_back(cuda):
    """The Newton polar fast path and the Jacobi SVD path give the same rotation; planar / mirrored point sets (where the
    reference's det fix matters, main_raw.py:62-65) take the Jacobi path and still match the oracle."""
    d = synth.make_fusion_clip(300, 70, seed=9, nan_frac=0.05)
    Xr = d["Xr"]
    Xr[100:150] = Xr[100:150] * np.array([1.0, 1.0, -1.0])        # mirrored right view: reflection case
    Xr[150:200, :, 1] = 0.25                                        # planar right view
    args = _dev(cuda, d["Xl"], Xr, d["Ul"], d["Ur"])
    a = fusion.fuse_clip(*args)
    b = fusion.fuse_clip(*args, force_jacobi=True)
    np.testing.assert_allclose(a.aligned.cpu().numpy(), b.aligned.cpu().numpy(), rtol=1e-11, atol=1e-11, equal_nan=True)
    np.testing.assert_allclose(a.fused.cpu().numpy(), b.fused.cpu().numpy(), rtol=1e-11, atol=1e-11, equal_nan=True)
    fused, ql, qr, Xa = F.fuse_clip(d["Xl"], Xr, d["Ul"], d["Ur"])
    al = a.aligned.cpu().numpy()
    ok = np.isfinite(Xa).all(-1)
    np.testing.assert_allclose(al[ok], Xa[ok], rtol=TOL, atol=TOL)
    np.testing.assert_allclose(a.fused.cpu().numpy(), fused, rtol=TOL, atol=TOL, equal_nan=True)


def test_batched_kernel_equals_per_frame_kernel(cuda):
    """The product path (moments -> per-frame parameters -> per-joint fusion, centred moments from raw ones) and the
    single warp-per-frame kernel agree far inside the parity tolerance, statuses included, for ragged clip lengths."""
    for T in (1, 15, 16, 17, 333):
        d = synth.make_fusion_clip(T, 70, seed=T, nan_frac=0.06)
        if T >= 17:
            d["Xr"][3, 2:] = np.nan
            d["Ur"][5] = np.nan
        args = _dev(cuda, d["Xl"], d["Xr"], d["Ul"], d["Ur"])
        a = fusion.fuse_clip(*args, strict=False)
        b = fusion.fuse_clip(*args, strict=False, per_frame_kernel=True)
        assert torch.equal(a.status, b.status)
        for x, y in ((a.fused, b.fused), (a.q_l, b.q_l), (a.q_r, b.q_r), (a.aligned, b.aligned)):
            np.testing.assert_allclose(x.cpu().numpy(), y.cpu().numpy(), rtol=1e-10, atol=1e-10, equal_nan=True)


def test_fuse_edge_cases(cuda):
    d = synth.make_fusion_clip(6, 70, seed=1, nan_frac=0.0)
    Xl, Xr, Ul, Ur = d["Xl"], d["Xr"], d["Ul"], d["Ur"]
    Xr[1, 2:] = np.nan      # < 3 common joints: right view stays unaligned (main_raw.py:83-84); its fit also fails (< 8 points)
    Xl[2, :65] = np.nan     # left fit impossible (confidence.py:31-32)
    Ur[3] = np.nan          # right fit impossible
    Xl[4, 30] = np.nan      # only the right view has joint 30
    Xr[4, 31] = np.nan      # only the left view has joint 31
    Xl[4, 32] = np.nan
    Xr[4, 32] = np.nan      # nobody has joint 32
    args = _dev(cuda, Xl, Xr, Ul, Ur)
    with pytest.raises(ValueError):
        fusion.fuse_clip(*args)
    r = fusion.fuse_clip(*args, strict=False)
    st = r.status.cpu().numpy()
    assert st[0] == 0 and st[4] == 0 and st[5] == 0
    assert st[1] == _cabi.FUSE_NO_ALIGN | _cabi.FUSE_FIT_RIGHT_FAILED
    assert st[2] & _cabi.FUSE_FIT_LEFT_FAILED and st[3] == _cabi.FUSE_FIT_RIGHT_FAILED
    fz = r.fused.cpu().numpy()
    assert np.isnan(fz[1:4]).all()
    for t in (0, 4, 5):
        ref = F.fuse_clip(Xl[t:t + 1], Xr[t:t + 1], Ul[t:t + 1], Ur[t:t + 1])
        np.testing.assert_allclose(fz[t], ref[0][0], rtol=TOL, atol=TOL, equal_nan=True)
    al = r.aligned.cpu().numpy()
    np.testing.assert_allclose(fz[4, 30], al[4, 30], rtol=0, atol=0)
    np.testing.assert_array_equal(fz[4, 31], Xl[4, 31])
    assert np.isnan(fz[4, 32]).all()
    # empty clip, one frame, bad arguments
    e = fusion.fuse_clip(*[a[:0] for a in args])
    assert e.fused.shape == (0, 70, 3)
    with pytest.raises(RuntimeError):
        fusion.fuse_clip(*[a.cpu() for a in args])
    with pytest.raises(ValueError):
        fusion.fuse_clip(args[0], args[1][:, :60], args[2], args[3])
    with pytest.raises(ValueError):
        fusion.fuse_clip(*args, scale_mode="arm")
    big = torch.zeros((2, 97, 3), dtype=torch.float64, device=cuda)
    with pytest.raises(ValueError):
        fusion.fuse_clip(big, big, big[..., :2], big[..., :2])


@pytest.mark.parametrize("key,kw", [
    ("ema_adaptive", dict(alpha=0.7, adaptive=True, alpha_min=0.45, alpha_max=0.92, speed_gain=0.25)),
    ("ema_fixed", dict(alpha=0.7, adaptive=False)),
    ("ema_gain", dict(alpha=0.6, adaptive=True, alpha_min=0.3, alpha_max=0.95, speed_gain=2.0)),
    ("ema_sparse", dict(alpha=0.7, adaptive=True, alpha_min=0.45, alpha_max=0.92, speed_gain=0.25)),
])
def test_ema_matches_reference_golden(cuda, golden, key, kw):
    g = golden("g8_fusion.npz")
    X = g["fused_sparse"] if key == "ema_sparse" else g["fused"]
    (Xd,) = _dev(cuda, X)
    Ye = fusion.temporal_smooth_ema(Xd, exact=True, **kw).cpu().numpy()
    np.testing.assert_allclose(Ye, g[key], rtol=1e-14, atol=0, equal_nan=True)
    Yc = fusion.temporal_smooth_ema(Xd, chunk=16, **kw).cpu().numpy()  # 60 frames in 4 chunks, each replaying its halo
    np.testing.assert_allclose(Yc, g[key], rtol=1e-13, atol=0, equal_nan=True)


def test_ema_chunked_equals_sequential_with_long_gaps(cuda):
    rng = np.random.default_rng(5)
    T, J = 20_000, 70
    X = synth.skeleton_clip(T, J, rng)
    X[rng.random((T, J)) < 0.25] = np.nan
    X[3000:3900, :10] = np.nan     # a gap longer than a chunk: the state is held across it
    X[:700, 20] = np.nan           # a joint that appears late
    X[:, 21] = np.nan              # a joint that never appears
    (Xd,) = _dev(cuda, X)
    ids = list(range(J))
    Ye = fusion.temporal_smooth_ema(Xd, ids, exact=True).cpu().numpy()
    ref = F.temporal_smooth_ema(X[:2500], ids)
    np.testing.assert_allclose(Ye[:2500], ref, rtol=1e-14, atol=0, equal_nan=True)
    for chunk in (512, 64, 4096):
        Yc = fusion.temporal_smooth_ema(Xd, ids, chunk=chunk).cpu().numpy()
        np.testing.assert_allclose(Yc, Ye, rtol=1e-13, atol=0, equal_nan=True)
    assert np.isnan(Ye[:, 21]).all() and np.isnan(Ye[:700, 20]).all() and np.isfinite(Ye[701:, 20]).all()
    assert np.isfinite(Ye[3000:3900, :10]).all()  # held


def test_dict_interfaces_match_reference_golden(cuda, golden):
    g = golden("g8_fusion.npz")
    T, J = g["Xl"].shape[:2]
    frames = {}
    for t in range(T):
        frames[t] = {k: {"pred": {j: v[t, j] for j in range(J)}} for k, v in
                     (("L_3D", g["Xl"]), ("R_3D", g["Xr"]), ("L_2D", g["Ul"]), ("R_2D", g["Ur"]))}
    fused_seq, smooth_seq, ids = fusion.fuse_person(frames)
    assert ids == list(range(J)) and len(fused_seq) == len(smooth_seq) == T
    for t in (0, 17, T - 1):
        have = {j for j in range(J) if np.isfinite(g["fused"][t, j]).all()}
        assert set(fused_seq[t]) == have
        for j in have:
            np.testing.assert_allclose(fused_seq[t][j], g["fused"][t, j], rtol=TOL, atol=TOL)
        for j, v in smooth_seq[t].items():
            np.testing.assert_allclose(v, g["ema_adaptive"][t, j], rtol=TOL, atol=TOL)
    out = fusion.temporal_smooth_ema_dicts(fused_seq, ids, alpha=0.7, adaptive=False)
    np.testing.assert_allclose(out[30][14], g["ema_fixed"][30, 14], rtol=TOL, atol=TOL)
    assert fusion.temporal_smooth_ema_dicts([], ids) == []


def test_full_size_properties(cuda):
    """1M frames x 70 joints: rigid-motion equivariance of the fusion (moving the right view's frame does not change the
    result), identical views fuse to themselves, EMA of a constant is the constant, chunked EMA == sequential."""
    T, J = 1_000_000, 70
    g = torch.Generator(device=cuda).manual_seed(0)
    f64 = dict(dtype=torch.float64, device=cuda)
    base = torch.randn(1, J, 3, generator=g, **f64) * 0.4
    Xl = base + torch.randn(T, J, 3, generator=g, **f64) * 0.02 + torch.tensor([0.0, 0.0, 10.0], **f64)
    Xr_same = Xl + torch.randn(T, J, 3, generator=g, **f64) * 0.02
    Ul = Xl[..., :2] / Xl[..., 2:3] * 1100.0 + 960.0
    Ur = Xr_same[..., :2] / Xr_same[..., 2:3] * 1100.0 + 960.0
    miss = torch.rand(T, J, generator=g, device=cuda) < 0.02
    miss[:, [14, 11, 12, 5, 6]] = False
    Xr_same[miss] = float("nan")
    R = torch.tensor(synth.rot_y(1.1), **f64)
    Xr_moved = Xr_same @ R.T + torch.tensor([3.0, -1.0, 2.0], **f64)
    a = fusion.fuse_clip(Xl, Xr_same, Ul, Ur, want=())
    b = fusion.fuse_clip(Xl, Xr_moved, Ul, Ur, want=())
    assert int(a.status.max()) == 0
    d = (a.fused - b.fused).abs()
    assert torch.isnan(a.fused).sum() == 0 and float(d.max()) < 1e-9
    same = fusion.fuse_clip(Xl, Xl, Ul, Ul, want=()).fused
    assert float(((same - Xl).abs() / Xl.abs().clamp_min(1.0)).max()) < 1e-7  # (wl + wr) / (wl + wr + 1e-8)
    const = Xl[:1].expand(200_000, J, 3).contiguous()
    assert float((fusion.temporal_smooth_ema(const) - const).abs().max()) < 1e-14  # a x + (1 - a) x rounds, in numpy too
    Ye = fusion.temporal_smooth_ema(a.fused, exact=True)
    Yc = fusion.temporal_smooth_ema(a.fused)
    assert float(((Ye - Yc).abs() / Ye.abs().clamp_min(1.0)).max()) < 1e-13


@pytest.mark.parametrize("name", ["plain", "weighted", "scaled"])
def test_rigid_transform_3d_matches_reference_golden(cuda, golden, name):
    """bundle_adjustment/fuse/fuse.py:rigid_transform_3D through ska_rigid_fuse_f64 against golden G11 (the reference's own
    outputs): fused joints, per-frame R / t / s and the diagnostics."""
    g = golden("g11_rigid_fuse.npz")
    ok = g[f"{name}_ok"]
    kw = {"plain": {}, "weighted": dict(wL=g["wL"][ok], wR=g["wR"][ok], tau=0.05), "scaled": dict(allow_scale=True, wL=g["wL"][0], wR=g["wR"][0])}[name]
    L, R = _dev(cuda, g["L"][ok], g["R"][ok])
    r = fusion.rigid_fuse_clip(L, R, **kw)
    np.testing.assert_allclose(r.fused.cpu().numpy(), g[f"{name}_fused"], rtol=TOL, atol=TOL, equal_nan=True)
    np.testing.assert_allclose(r.R.cpu().numpy(), g[f"{name}_R"], atol=TOL)
    np.testing.assert_allclose(r.t.cpu().numpy(), g[f"{name}_t"], atol=TOL)
    np.testing.assert_allclose(r.s.cpu().numpy(), g[f"{name}_s"], rtol=TOL)
    np.testing.assert_allclose(r.diag.cpu().numpy(), g[f"{name}_diag"], rtol=TOL, atol=TOL, equal_nan=True)
    # the reference's own signature: numpy in, (fused, diag dict) out; single frame (J,3) too
    fused, diag = fusion.rigid_transform_3D(g["L"][ok], g["R"][ok], **kw)
    np.testing.assert_allclose(fused, g[f"{name}_fused"], rtol=TOL, atol=TOL, equal_nan=True)
    assert abs(diag["mean_gain"] - float(g[f"{name}_mean_gain"])) < 1e-9 and diag["bad_frames"] == list(g[f"{name}_bad"])
    assert set(diag["per_frame"][0]) == {"frame", "LR_before", "Fused_vs_L", "Fused_vs_R", "gain", "s", "R", "t"}
    if name == "plain":
        from skiing_analysis_pytorch_b200 import dropin

        fs, _ = dropin.shim("front_side.side.fuse.fuse").rigid_transform_3D(target=g["L"][0], source=g["R"][0], wL=None, wR=None,
                                                                          return_diagnostics=True)  # the call of front_side/side/run.py:81
        np.testing.assert_allclose(fs, g["single_fused"], rtol=TOL, atol=TOL, equal_nan=True)
        f1, d1 = fusion.rigid_transform_3D(g["L"][0], g["R"][0])
        np.testing.assert_allclose(f1, g["single_fused"], rtol=TOL, atol=TOL, equal_nan=True)
        assert f1.shape == (70, 3) and len(d1["per_frame"]) == 1
        assert fusion.rigid_transform_3D(g["L"][0], g["R"][0], return_diagnostics=False)[1] is None


def test_rigid_fuse_edges_and_oracle(cuda):
    d = synth.make_fusion_clip(500, 70, seed=31, nan_frac=0.04)
    full = synth.make_fusion_clip(500, 70, seed=31, nan_frac=0.0)
    L, R = d["Xl"], d["Xr"]
    L[:, fusion.TORSO_IDX[:4]], R[:, fusion.TORSO_IDX[:4]] = full["Xl"][:, fusion.TORSO_IDX[:4]], full["Xr"][:, fusion.TORSO_IDX[:4]]  # >= 3 torso joints everywhere
    R[100:150] = R[100:150] * np.array([1.0, 1.0, -1.0])   # mirrored right view: the reflection fix of fuse_check.py:57-62
    tau = np.linspace(0.02, 0.3, 70)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fo, Ro, to, so, do = F.rigid_transform_3D(L, R, tau=tau, allow_scale=True)
    r = fusion.rigid_fuse_clip(*_dev(cuda, L, R), tau=tau, allow_scale=True)
    np.testing.assert_allclose(r.fused.cpu().numpy(), fo, rtol=TOL, atol=TOL, equal_nan=True)
    np.testing.assert_allclose(r.R.cpu().numpy(), Ro, atol=TOL)
    np.testing.assert_allclose(r.s.cpu().numpy(), so, rtol=TOL)
    np.testing.assert_allclose(r.diag.cpu().numpy(), do, rtol=TOL, atol=TOL, equal_nan=True)
    bad = L.copy()
    bad[7, [69, 9, 10]] = np.nan
    with pytest.raises(ValueError):
        fusion.rigid_fuse_clip(*_dev(cuda, bad, R))
    st = fusion.rigid_fuse_clip(*_dev(cuda, bad, R), strict=False).status.cpu().numpy()
    assert st[7] == 1 and st.sum() == 1
    with pytest.raises(ValueError):
        fusion.rigid_fuse_clip(*_dev(cuda, L[:, :60], R[:, :60]))
    with pytest.raises(ValueError):
        fusion.rigid_fuse_clip(*_dev(cuda, L, R), wL=np.ones(5))
    e = fusion.rigid_fuse_clip(*_dev(cuda, L[:0], R[:0]))
    assert e.fused.shape == (0, 70, 3)
