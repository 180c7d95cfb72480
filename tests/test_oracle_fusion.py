"""The fusion oracle (oracle/fusion.py, array form) against golden G8 = outputs of the reference's own dict-based
fuse/main_raw.py, fuse/confidence.py and fuse/fuse.py.  Same LAPACK calls in the same order: 1e-12."""
import numpy as np
import pytest

from oracle import fusion as F


def test_fuse_clip_matches_reference(golden):
    g = golden("g8_fusion.npz")
    fused, ql, qr, Xa = F.fuse_clip(g["Xl"], g["Xr"], g["Ul"], g["Ur"])
    np.testing.assert_allclose(Xa, g["aligned"], rtol=1e-12, atol=1e-12, equal_nan=True)
    np.testing.assert_allclose(ql, g["q_l"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(qr, g["q_r"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(fused, g["fused"], rtol=1e-12, atol=1e-12, equal_nan=True)
    assert np.isnan(g["fused"]).any() and (g["q_l"] > 0.5).any() and (g["q_l"] < 0.1).any()
    c1, _ = F.weakpersp_reproj_confidence(g["Xl"][5], g["Ul"][5])
    c2, _ = F.crossview_consistency_confidence(g["Xl"][5], g["Xr"][5])
    np.testing.assert_allclose(c1, g["conf1_l"][5], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(c2, g["conf2"][5], rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("key,kw", [
    ("ema_adaptive", dict(alpha=0.7, adaptive=True, alpha_min=0.45, alpha_max=0.92, speed_gain=0.25)),
    ("ema_fixed", dict(alpha=0.7, adaptive=False)),
    ("ema_gain", dict(alpha=0.6, adaptive=True, alpha_min=0.3, alpha_max=0.95, speed_gain=2.0)),
])
def test_ema_matches_reference(golden, key, kw):
    g = golden("g8_fusion.npz")
    Y = F.temporal_smooth_ema(g["fused"], **kw)
    np.testing.assert_allclose(Y, g[key], rtol=1e-13, atol=1e-13, equal_nan=True)


def test_ema_hold_and_reset_branches_match_reference(golden):
    g = golden("g8_fusion.npz")
    Y = F.temporal_smooth_ema(g["fused_sparse"], alpha=0.7, adaptive=True, alpha_min=0.45, alpha_max=0.92, speed_gain=0.25)
    np.testing.assert_allclose(Y, g["ema_sparse"], rtol=1e-13, atol=1e-13, equal_nan=True)
    assert np.isnan(g["fused_sparse"][..., 0]).mean() > 0.2 and np.isnan(g["ema_sparse"][..., 0]).mean() < 0.05


def test_unity_pipeline_without_alignment_matches_reference(golden):
    """fuse/main_unity.py:_fuse_pair (no rigid alignment, 15 target joints, key indices as array positions) + EMA with
    the Unity joint ids selecting the per-joint alpha classes."""
    g = golden("g8_fusion.npz")
    fused, _, _, _ = F.fuse_clip(g["unity_Xl"], g["unity_Xr"], g["unity_Ul"], g["unity_Ur"], align=False)
    np.testing.assert_allclose(fused, g["unity_fused"], rtol=1e-12, atol=1e-12, equal_nan=True)
    Y = F.temporal_smooth_ema(fused, list(g["unity_ids"]))
    np.testing.assert_allclose(Y, g["unity_smooth"], rtol=1e-13, atol=1e-13, equal_nan=True)


def test_reference_errors_and_fallbacks():
    d = np.full((70, 3), np.nan)
    with pytest.raises(ValueError):
        F.weakpersp_reproj_confidence(d, np.zeros((70, 2)))
    X = np.random.default_rng(0).normal(size=(70, 3))
    Xr = X.copy()
    Xr[2:] = np.nan  # fewer than 3 common joints: right view returned unchanged (main_raw.py:83-84)
    np.testing.assert_array_equal(F.align_right_to_left(X, Xr), Xr)
    Xk = X.copy()
    Xk[14] = np.nan  # a key joint missing: canonicalisation undefined -> confidence 0 everywhere (confidence.py:150-153)
    c, dist = F.crossview_consistency_confidence(Xk, X)
    assert (c == 0).all() and np.isnan(dist).all()


@pytest.mark.parametrize("name", ["plain", "weighted", "scaled"])
def test_rigid_transform_3d_matches_reference(golden, name):
    """bundle_adjustment/fuse/fuse.py:rigid_transform_3D (Umeyama on the torso joints + threshold fusion + diagnostics)."""
    import warnings

    g = golden("g11_rigid_fuse.npz")
    ok = g[f"{name}_ok"]
    kw = {"plain": {}, "weighted": dict(wL=g["wL"][ok], wR=g["wR"][ok], tau=0.05), "scaled": dict(allow_scale=True, wL=g["wL"][0], wR=g["wR"][0])}[name]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fused, Rh, th, sh, dg = F.rigid_transform_3D(g["L"][ok], g["R"][ok], **kw)
    np.testing.assert_allclose(fused, g[f"{name}_fused"], rtol=1e-12, atol=1e-12, equal_nan=True)
    np.testing.assert_allclose(Rh, g[f"{name}_R"], atol=1e-12)
    np.testing.assert_allclose(th, g[f"{name}_t"], atol=1e-12)
    np.testing.assert_allclose(sh, g[f"{name}_s"], rtol=1e-12)
    np.testing.assert_allclose(dg, g[f"{name}_diag"], rtol=1e-12, atol=1e-14, equal_nan=True)
    assert np.isfinite(g[f"{name}_diag"][:12]).all() and np.isnan(g[f"{name}_diag"][12:]).any()
    if name == "scaled":
        assert np.abs(sh - 1.0).max() > 1e-4
    with pytest.raises(ValueError):
        bad = g["L"][:1].copy()
        bad[0, [69, 9, 10]] = np.nan
        F.rigid_transform_3D(bad, g["R"][:1])
