"""Pin the post-processing restatement (oracle/postprocess.py, row N2): against cv2.undistortPoints and
scipy.signal.savgol_filter directly (the third-party algorithms it restates) and against outputs of the
REFERENCE's own triangulation.postprocess functions (tests/golden/g7_post_triage.npz)."""
import warnings

import numpy as np
import pytest

from oracle import postprocess as OP
from skiing_analysis_pytorch_b200 import synth

KEYS = ["rmse_px", "median_err_px", "pos_depth_ratio", "kept_ratio", "kept_count"]


def test_undistort_is_cv2_bit_for_bit():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    x = np.stack([rng.uniform(0, 1920, 500), rng.uniform(0, 1080, 500)], 1).astype(np.float32)
    for d in (synth.DIST_CALIB, synth.DIST_CALIB[:5], np.r_[synth.DIST_CALIB[:8], 1e-3, -2e-3, 1e-3, 5e-4]):
        ref = cv2.undistortPoints(x.reshape(-1, 1, 2), synth.K_CALIB, np.asarray(d), P=synth.K_CALIB).reshape(-1, 2)
        np.testing.assert_array_equal(OP.undistort_points(x, synth.K_CALIB, d), ref)


def test_savgol_is_scipy():
    sig = pytest.importorskip("scipy.signal")
    v = np.random.default_rng(1).normal(size=73)
    for w, p in ((9, 2), (5, 2), (11, 3), (25, 4), (3, 1)):
        np.testing.assert_allclose(OP.savgol_interp(v, w, p), sig.savgol_filter(v, w, p), atol=1e-11)
    assert [OP.effective_window(T, w) for T, w in ((80, 9), (80, 8), (6, 9), (7, 9), (2, 9))] == [9, 9, 5, 3, 3]


def test_golden_g7_reference_outputs(golden):
    g = golden("g7_post_triage.npz")
    X, kL, kR, K, R, t = g["X"], g["kptL"], g["kptR"], g["K"], g["R"], g["t"]
    cases = {"plain": dict(), "dist": dict(dist1=g["dist"], dist2=g["dist"]),
             "conf_smooth": dict(confL=g["confL"], confR=g["confR"], smooth=True),
             "tight_smooth6": dict(err_thresh_px=1.0, smooth=True, sg_win=6, sg_poly=3)}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name, kw in cases.items():
            Xc, st = OP.post_triage_sequence(X, kL, kR, K, K, R, t, **kw)
            ref = g[f"{name}_X"]
            assert Xc.dtype == np.float32
            np.testing.assert_array_equal(np.isnan(Xc), np.isnan(ref))
            np.testing.assert_allclose(Xc, ref, rtol=2e-6, atol=2e-6 if kw.get("smooth") else 0, equal_nan=True)  # float32 outputs
            np.testing.assert_allclose(np.array([[s[k] for k in KEYS] for s in st]), g[f"{name}_stats"], rtol=1e-12, atol=0, equal_nan=True)
        np.testing.assert_allclose(OP.smooth_skeleton(g["smooth_in"], 9, 2), g["smooth_out_9_2"], rtol=2e-6, atol=2e-6, equal_nan=True)
        np.testing.assert_allclose(OP.smooth_skeleton(g["smooth_in"], 8, 3), g["smooth_out_8_3"], rtol=2e-6, atol=2e-6, equal_nan=True)
    # the scenario actually exercises every gate
    keep = ~np.isnan(g["plain_X"][..., 0])
    assert not keep[3, 4] and not keep[10, 2] and not keep[20, 5] and keep.mean() > 0.5
