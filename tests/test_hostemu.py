"""CPU-side numerics gate for the kernel arithmetic: the exact __host__ __device__ code the GPU
runs (tests/hostemu compiles it with g++) against the fp64 oracle.  Tolerances are the north-star
ones: X within 1e-4 relative (we hold 1e-6), reprojection RMSE within 1e-5 px, per point 2e-4 px."""
import numpy as np
import pytest

from oracle import geometry as G
from skiing_analysis_pytorch_b200 import _cabi, synth
from tests import hostemu

X_REL_TOL = 1e-4      # north star
X_REL_HELD = 2e-6     # what the fp32 secular solver actually holds on these rigs
RMSE_TOL = 1e-5       # px, north star
POINT_TOL = 2e-4      # px per point (the reference's own f32 noise floor)


def _oracle(clip, conf, dist):
    V, T, J, _ = clip.x_vm.shape
    P = np.stack([G.make_P(clip.K[v], clip.R[v], clip.t[v]) for v in range(V)])
    x = clip.x_vm.reshape(V, -1, 2)
    w = None if conf is None else conf.reshape(V, -1)
    X = G.dlt_triangulate(P, x, w)
    err = np.stack([np.linalg.norm(G.project_cv(X, clip.R[v], clip.t[v], clip.K[v], dist) - x[v], axis=1) for v in range(V)])
    return X, err


CASES = [
    ("2a", 64, 17, False, None),
    ("2a", 64, 17, False, synth.DIST_CALIB),
    ("2b", 64, 17, False, synth.DIST_CALIB),
    ("2b", 64, 17, True, None),
    ("3", 32, 17, True, synth.DIST_CALIB),
    ("4", 32, 17, True, synth.DIST_CALIB[:5]),
    ("8", 16, 70, True, synth.DIST_CALIB),
]


@pytest.mark.parametrize("rig,T,J,use_conf,dist", CASES)
@pytest.mark.parametrize("solver", ["secular", "jacobi64"])
def test_hostemu_matches_oracle(rig, T, J, use_conf, dist, solver):
    clip = synth.make_clip(rig, T, J, seed=0)
    V = len(clip.R)
    conf = clip.conf_vm if use_conf else None
    Xo, eo = _oracle(clip, conf, dist)
    cams = _cabi.make_cameras(clip.K, clip.R, clip.t, dist)
    X, err, st = hostemu.triangulate(cams, V, clip.x_vm.reshape(V, -1, 2), None if conf is None else conf.reshape(V, -1),
                                     flags=_cabi.SOLVERS[solver])
    rel = np.linalg.norm(X - Xo, axis=1) / np.linalg.norm(Xo, axis=1)
    assert rel.max() < X_REL_HELD < X_REL_TOL
    assert np.abs(err - eo).max() < POINT_TOL
    rm = lambda e: np.sqrt(np.mean(np.asarray(e, np.float64) ** 2))
    assert abs(rm(err) - rm(eo)) < RMSE_TOL
    # the near-degenerate FIXED rig sends its worst-conditioned points (<2%) to the fp64 path
    assert (st <= 1).all() and (st == 1).mean() <= (0.02 if rig == "2a" else 0.0)


def test_hostemu_fallback_certificate():
    """Near-degenerate geometry + large noise: the interlacing certificate must fail for some
    points and the fp64 Jacobi fallback must reproduce the exact-mode answer there."""
    clip = synth.make_clip("2a", 300, 17, seed=0, noise_px=20.0)
    cams = _cabi.make_cameras(clip.K, clip.R, clip.t)
    x = clip.x_vm.reshape(2, -1, 2)
    Xs, _, st = hostemu.triangulate(cams, 2, x, flags=_cabi.SOLVER_SECULAR)
    Xj, _, _ = hostemu.triangulate(cams, 2, x, flags=_cabi.SOLVER_JACOBI64)
    assert (st == 1).sum() > 0
    fb = st == 1
    np.testing.assert_array_equal(Xs[fb], Xj[fb])
    ok = st == 0
    rel = np.linalg.norm(Xs[ok] - Xj[ok], axis=1) / np.linalg.norm(Xj[ok], axis=1)
    assert rel.max() < 1e-4


def test_hostemu_centre_invariance():
    """The conditioning origin must not change the answer beyond rounding."""
    clip = synth.make_clip("2b", 32, 17, seed=2)
    cams = _cabi.make_cameras(clip.K, clip.R, clip.t, synth.DIST_CALIB)
    x = clip.x_vm.reshape(2, -1, 2)
    Xa, ea, _ = hostemu.triangulate(cams, 2, x)
    Xb, eb, _ = hostemu.triangulate(cams, 2, x, centre=[0.0, 0.0, 0.0])
    Xc, ec, _ = hostemu.triangulate(cams, 2, x, centre=[3.0, -2.0, 14.0])
    for Xo, eo in ((Xb, eb), (Xc, ec)):
        assert (np.linalg.norm(Xa - Xo, axis=1) / np.linalg.norm(Xa, axis=1)).max() < 5e-6
        assert np.abs(ea - eo).max() < 2e-3


def test_hostemu_nan_propagates():
    clip = synth.make_clip("2b", 4, 17, seed=2)
    cams = _cabi.make_cameras(clip.K, clip.R, clip.t)
    x = clip.x_vm.reshape(2, -1, 2).copy()
    x[1, 5, 0] = np.nan
    X, err, st = hostemu.triangulate(cams, 2, x)
    assert np.isnan(X[5]).all() and st[5] == 2
    assert np.isfinite(np.delete(X, 5, axis=0)).all()
