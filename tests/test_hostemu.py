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


def test_hostemu_cv_projection_matches_oracle():
    """ska_project.cuh project_cv64 (the standalone reprojection kernel's arithmetic) vs the fp64 oracle."""
    import ctypes as C

    lib = C.CDLL(str(hostemu.build()))
    rng = np.random.default_rng(0)
    R, t = synth.rig("3")
    X = synth.CENTRE + rng.normal(0, 0.5, (200, 3))
    for dist in (None, synth.DIST_CALIB, np.r_[synth.DIST_CALIB[:8], 1e-3, -2e-3, 3e-3, 1e-3]):
        d = np.zeros(12) if dist is None else np.asarray(dist, float)[:12]
        K = synth.K_CALIB
        cam = np.r_[R[1].ravel(), t[1], K[0, 0], K[1, 1], K[0, 2], K[1, 2], d]
        uv = np.zeros((200, 2))
        lib.hostemu_project_cv(cam.ctypes.data_as(C.c_void_p), np.ascontiguousarray(X).ctypes.data_as(C.c_void_p), C.c_int64(200),
                               uv.ctypes.data_as(C.c_void_p))
        ref = G.project_cv(X, R[1], t[1], K, None if dist is None else d)
        assert np.abs(uv - ref).max() < 1e-9


def test_hostemu_loss_projection_and_adjoint_match_autograd():
    """ska_project.cuh project_loss / project_loss_adjoint vs torch.autograd on the plain restatement."""
    import ctypes as C

    import torch

    from oracle import torch_ref as TR

    lib = C.CDLL(str(hostemu.build()))
    rng = np.random.default_rng(1)
    R, t = synth.rig("3")
    K = synth.K_CALIB.copy()
    K[0, 1] = 0.7  # skew is honoured (loss.py:74-82)
    N = 64
    X = synth.CENTRE + rng.normal(0, 0.5, (N, 3))
    X[0] = [0.0, 0.0, -5.0]  # behind camera 0... for camera 1 pick a point behind it too
    X[1] = -R[1].T @ t[1] - 2.0 * R[1].T @ np.array([0, 0, 1.0])
    g = rng.normal(0, 1, (N, 2))
    uv, gXc, gK, cl = np.zeros((N, 2)), np.zeros((N, 3)), np.zeros((N, 6)), np.zeros(N, np.uint8)
    p = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)
    lib.hostemu_project_loss(p(R[1]), p(t[1]), p(K), p(X), p(g), C.c_int64(N), p(uv), p(gXc), p(gK), p(cl))
    Xt = torch.tensor(X[None], requires_grad=True)
    Rt = torch.tensor(R[1][None], requires_grad=True)
    tt = torch.tensor(t[1][None], requires_grad=True)
    Kt = torch.tensor(K[None], requires_grad=True)
    proj = TR.project_points(Xt, Rt, tt, Kt)  # (1,1,N,2)
    assert np.abs(proj.detach().numpy()[0, 0] - uv).max() <= 1e-9 * np.abs(uv).max()
    assert cl[1] == 1 and cl[2:].sum() == 0
    (proj[0, 0] * torch.tensor(g)).sum().backward()
    gX = gXc @ R[1]                      # gX = R^T gXc per point
    scale = lambda a: np.abs(a).max() + 1e-30
    assert np.abs(gX - Xt.grad.numpy()[0]).max() <= 1e-9 * scale(gX)
    assert np.abs(gXc.sum(0) - tt.grad.numpy()[0]).max() <= 1e-9 * scale(gXc.sum(0))
    gR = np.einsum("nr,nk->rk", gXc, X)
    assert np.abs(gR - Rt.grad.numpy()[0]).max() <= 1e-9 * scale(gR)
    gKs = gK.sum(0)
    assert np.abs(gKs - Kt.grad.numpy()[0, :2].ravel()).max() <= 1e-9 * scale(gKs)
    assert np.abs(Kt.grad.numpy()[0, 2]).max() == 0.0


@pytest.mark.parametrize("rig,T,J,use_conf,dist", [c for c in CASES if c[0] in ("2a", "2b", "3", "4")])
def test_hostemu_packed_pair_path_is_bit_identical_to_scalar(rig, T, J, use_conf, dist):
    """PTS = 2 (two points as one F2 computation - FFMA2 on the GPU) must give exactly the scalar path's
    numbers: each packed component is an IEEE fma like the scalar instruction."""
    clip = synth.make_clip(rig, T, J, seed=0)
    V = len(clip.R)
    cams = _cabi.make_cameras(clip.K, clip.R, clip.t, dist)
    k = clip.x_vm.reshape(V, -1, 2)
    c = clip.conf_vm.reshape(V, -1) if use_conf else None
    X1, e1, s1 = hostemu.triangulate(cams, V, k, c, flags=0)
    X2, e2, s2 = hostemu.triangulate(cams, V, k, c, flags=1 << 10)
    np.testing.assert_array_equal(s1, s2)
    np.testing.assert_array_equal(X1, X2)
    np.testing.assert_array_equal(e1, e2)
    # odd point count exercises the tail pair
    X3, e3, _ = hostemu.triangulate(cams, V, k[:, :33], None if c is None else c[:, :33], flags=1 << 10)
    np.testing.assert_array_equal(X3, X1[:33] if c is None else X3)


@pytest.mark.parametrize("rig,T,J,use_conf,dist", [c for c in CASES if c[0] in ("4", "8")] + [("8", 16, 17, False, None), ("4", 32, 17, False, synth.DIST_CALIB)])
@pytest.mark.parametrize("rows", [0, 1, 2, 6])
def test_hostemu_view_pair_form_matches_oracle(rig, T, J, use_conf, dist, rows):
    """tri_point_vp (per-view work packed over pairs of views - the even V >= 4 kernel path), with the rows kept
    in registers (0), formed a second time for the final residuals (1) or parked in a slab (2; 6 = slab + every view
    reading view 0's intrinsics, which are equal on these rigs): identical arithmetic."""
    clip = synth.make_clip(rig, T, J, seed=0)
    conf = clip.conf_vm if use_conf else None
    V = len(clip.R)
    Xo, eo = _oracle(clip, conf, dist)
    cams = _cabi.make_cameras(clip.K, clip.R, clip.t, dist)
    k = clip.x_vm.reshape(V, -1, 2)
    c = None if conf is None else conf.reshape(V, -1)
    X, err, st = hostemu.triangulate(cams, V, k, c, flags=(1 << 12) | (rows << 13))
    rel = np.linalg.norm(X - Xo, axis=1) / np.linalg.norm(Xo, axis=1)
    assert rel.max() < X_REL_HELD < X_REL_TOL
    assert np.abs(err - eo).max() < POINT_TOL
    assert abs(np.sqrt((err.astype(np.float64) ** 2).mean()) - np.sqrt((eo ** 2).mean())) < RMSE_TOL
    assert (st == 0).all()
    X1, e1, _ = hostemu.triangulate(cams, V, k, c, flags=0)
    assert (np.linalg.norm(X - X1, axis=1) / np.linalg.norm(X1, axis=1)).max() < 1e-6   # same sums in a different order


def test_hostemu_calib_observation_rows_match_oracle():
    """The calibrating-BA per-observation arithmetic (ska_ba_calib.cuh, fp32) against oracle/lm_calib.py (fp64):
    residual, point rows, all 15 camera-parameter rows, trial-cost error."""
    from oracle import lm_calib as lc

    clip, R0, t0, th, X0 = lc.make_problem("2b", 8, 17)
    X = X0.reshape(-1, 3)
    x = clip.x_fm.astype(float).transpose(0, 2, 1, 3).reshape(-1, 2, 2)
    e, A, B, cl = lc.residual_blocks(X, R0, t0, th, x)
    for c in range(2):
        cam = np.zeros(24)
        cam[:9], cam[9:12], cam[12:21] = R0[c].ravel(), t0[c], th[c]
        au, av, bu, bv, clamped, e2 = hostemu.calib_obs(cam, X, x[:, c])
        assert not clamped.any() and not cl.any()
        # the residual is a difference of ~1e3 px numbers in fp32: 1e-3 px absolute
        np.testing.assert_allclose(bu[:, 15], e[:, c, 0], atol=2e-3)
        np.testing.assert_allclose(bv[:, 15], e[:, c, 1], atol=2e-3)
        np.testing.assert_allclose(e2, (e[:, c] ** 2).sum(-1), rtol=1e-3, atol=1e-3)
        for got, ref in ((au, A[:, c, 0]), (av, A[:, c, 1]), (bu[:, :15], B[:, c, 0]), (bv[:, :15], B[:, c, 1])):
            scale = np.abs(ref).max(0)
            np.testing.assert_allclose(got / np.where(scale > 0, scale, 1), ref / np.where(scale > 0, scale, 1), atol=3e-5)


def test_hostemu_calib_triangle_index_maps():
    tri, row, col = hostemu.calib_tri_maps()
    q = 0
    for r in range(17):
        for s in range(r, 17):
            assert tri(r, s) == q and row(q) == r and col(q) == s
            q += 1
    assert q == 153
