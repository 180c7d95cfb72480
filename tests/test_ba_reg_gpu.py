"""GPU parity of the regularised LM (csrc/ska_ba_reg.cu through the C ABI) against its fp64 specification oracle/lm_reg.py:
the pieces of one linearisation entry by entry (gradient, diagonal, matrix-vector product, frame-Schur preconditioner), then
the LM trajectory (cost per trial within 1e-4 relative - the north-star tolerance - and identical accept / reject decisions)
against the exact-solve oracle and the committed golden history (tests/golden/g12_lm_reg.npz)."""
from pathlib import Path

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

from oracle import lm_reg
from skiing_analysis_pytorch_b200 import _cabi, ba_reg

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden" / "g12_lm_reg.npz"
COST_TOL = 1e-4


def _problem(name, cuda):
    rig, T, J, mode = lm_reg.G12_CASES[name]
    clip, R, t, X0 = lm_reg.make_problem(rig, T, J, cam_jitter=0.01)
    x2d, conf = clip.x_fm, clip.conf_fm
    s = ba_reg.RegularisedBundleAdjuster(torch.from_numpy(x2d).to(cuda), torch.from_numpy(conf).to(cuda), clip.K, R, t, X0, mode=mode,
                                         max_iters=16, cg_iters=80)
    return s, clip, R, t, X0, mode


def _rows_to_oracle(v, T, J, C):
    """(T, 3J + 6C) frame rows -> the oracle's stacked [X | cams] vector."""
    v = np.asarray(v)
    return np.concatenate([v[:, : 3 * J].reshape(-1), v[:, 3 * J:].reshape(-1)])


def _oracle_to_rows(u, T, J, C):
    return np.concatenate([u[: T * J * 3].reshape(T, 3 * J), u[T * J * 3:].reshape(T, 6 * C)], axis=1)


@pytest.mark.parametrize("name", sorted(lm_reg.G12_CASES))
def test_linearisation_pieces_match_oracle(cuda, name):
    s, clip, R, t, X0, mode = _problem(name, cuda)
    T, J, _ = X0.shape
    C = R.shape[1]
    x2d, conf = clip.x_fm.astype(float), clip.conf_fm.astype(float)
    coef = lm_reg.coefficients(T, J, C, conf.sum(), None)
    terms, _ = lm_reg.cost_terms(X0, R, t, clip.K, x2d, conf, coef)
    assert abs(s.cost_value - sum(terms.values())) <= 1e-12 * sum(terms.values())
    H, g, free, _ = lm_reg.normal_system(X0, R, t, clip.K, x2d, conf, coef, mode)
    lam = 1e-3
    s.linearize()
    torch.cuda.synchronize()
    g_gpu = _rows_to_oracle(s.vec[0, 1:-1].cpu().numpy(), T, J, C)
    D_gpu = _rows_to_oracle(s.vec[1, 1:-1].cpu().numpy(), T, J, C)
    np.testing.assert_allclose(g_gpu, _expand(g, free), rtol=0, atol=1e-11 * np.abs(g).max())
    np.testing.assert_allclose(D_gpu, _expand(H.diagonal(), free), rtol=0, atol=1e-11 * H.diagonal().max())
    # matrix-vector product
    rng = np.random.default_rng(5)
    p_free = rng.normal(size=g.shape[0])
    A = (H + lam * sp.diags(H.diagonal())).tocsr()
    s.vec[5, 1:-1] = torch.from_numpy(_oracle_to_rows(_expand(p_free, free), T, J, C)).to(cuda)
    s.cg(_cabi.BA_REG_CG_MATVEC)
    torch.cuda.synchronize()
    y_gpu = _rows_to_oracle(s.vec[6, 1:-1].cpu().numpy(), T, J, C)
    y_ref = _expand(A @ p_free, free)
    np.testing.assert_allclose(y_gpu, y_ref, rtol=0, atol=1e-11 * np.abs(y_ref).max())
    assert abs(float(s.dot.item()) - p_free @ (A @ p_free)) <= 1e-11 * abs(p_free @ (A @ p_free))
    # preconditioner: the frame's system without the within-frame bone / baseline coupling, solved exactly
    nX = T * J * 3
    k = {"pose_only": 0, "pose_cam_t": 3, "full": 6}[mode]
    fr = np.concatenate([np.repeat(np.arange(T), J * 3), np.repeat(np.arange(T), C * k)])
    blk = np.concatenate([np.repeat(np.arange(T * J), 3), T * J + np.repeat(np.arange(T * C), k)])
    ispt = np.concatenate([np.ones(nX, bool), np.zeros(T * C * k, bool)])
    Ac = A.tocoo()
    keep = (blk[Ac.row] == blk[Ac.col]) | ((fr[Ac.row] == fr[Ac.col]) & (ispt[Ac.row] != ispt[Ac.col]))
    M = sp.coo_matrix((Ac.data[keep], (Ac.row[keep], Ac.col[keep])), shape=Ac.shape).tocsc()
    r_free = rng.normal(size=g.shape[0])
    s.vec[3, 1:-1] = torch.from_numpy(_oracle_to_rows(_expand(r_free, free), T, J, C)).to(cuda)
    s.cg(_cabi.BA_REG_CG_BEGIN)
    torch.cuda.synchronize()
    z_gpu = _rows_to_oracle(s.vec[4, 1:-1].cpu().numpy(), T, J, C)
    z_ref = _expand(spla.splu(M).solve(r_free), free)
    np.testing.assert_allclose(z_gpu, z_ref, rtol=0, atol=1e-8 * np.abs(z_ref).max())


def _expand(v_free, free):
    out = np.zeros(free.shape[0])
    out[free] = v_free
    return out


def _check(hist, ref, n):
    """Trial by trial until the oracle has converged to machine precision (|F - F_trial| < 1e-10 F): from there the
    accept / reject decision is a coin toss of the last bit and the damping sequences part ways."""
    for k in range(n):
        g, o = hist[k], ref[k]
        assert abs(g["cost"] - o["cost"]) <= COST_TOL * o["cost"], (k, g, o)
        assert abs(g["trial_cost"] - o["trial_cost"]) <= COST_TOL * o["trial_cost"], (k, g, o)
        assert g["n_clamped"] == o["n_clamped"]
        if abs(o["cost"] - o["trial_cost"]) < 1e-10 * o["cost"]:
            return k
        assert g["accepted"] == o["accepted"], (k, g, o)
    return n


@pytest.mark.parametrize("name", sorted(lm_reg.G12_CASES))
def test_trajectory_matches_oracle_and_golden(cuda, name):
    s, clip, R, t, X0, mode = _problem(name, cuda)
    n = 6
    s.run(n)
    hist = s.history
    _, _, Xo, ref = lm_reg.run_lm(X0, R, t, clip.K, clip.x_fm.astype(float), clip.conf_fm.astype(float), num_iters=n, mode=mode)
    n_cmp = _check(hist, ref, n)
    assert n_cmp >= 3
    for k in range(n_cmp):
        for term in lm_reg.TERMS:
            assert abs(hist[k][term] - ref[k][term]) <= COST_TOL * max(ref[k]["cost"] * 1e-3, abs(ref[k][term])), (k, term)
        assert hist[k]["cg_residual"] <= 1e-7 or hist[k]["cg_iters"] == 80
    np.testing.assert_allclose(s.X.cpu().numpy(), Xo, rtol=0, atol=1e-5)
    gold = np.load(GOLDEN)
    gc = gold[f"{name}_cost"]
    for k in range(min(n, len(gc))):
        assert abs(hist[k]["cost"] - gc[k]) <= COST_TOL * gc[k]
    # the reference's own loss functions evaluated at the oracle's final iterate (generated by oracle/make_golden.py)
    assert abs(gold[f"{name}_ref_loss_final"] - gold[f"{name}_final_cost"]) <= 1e-9 * gold[f"{name}_final_cost"]


def test_graph_replay_equals_eager(cuda):
    a, *_ = _problem("c3_full", cuda)
    b, *_ = _problem("c3_full", cuda)
    a.run(5)
    b.run(5, graph=True)
    for ha, hb in zip(a.history, b.history):
        assert ha["trial_cost"] == hb["trial_cost"] and ha["accepted"] == hb["accepted"] and ha["cg_iters"] == hb["cg_iters"]


def test_run_local_ba_lm_signature_and_modes(cuda):
    """The reference's call (vggt/multi_view_process.py:553-564) with optimizer='lm': per-frame cameras are kept per frame,
    pose_only returns the cameras untouched, the configured objective decreases."""
    from skiing_analysis_pytorch_b200 import ba

    clip, R, t, X0 = lm_reg.make_problem("2b", 30, 17, cam_jitter=0.01)
    args = dict(K_torch=torch.from_numpy(clip.K).float(), R_init_torch=torch.from_numpy(R), t_init_torch=torch.from_numpy(t),
                X3d_init_torch=torch.from_numpy(X0), x2d_torch=torch.from_numpy(clip.x_fm).float(),
                conf2d_torch=torch.from_numpy(clip.conf_fm).float(), num_iters=5, lr=1e-3, device="cuda")
    for mode in ba_reg.MODES:
        Ro, to, Xo, hist = ba.run_local_ba(mode=mode, optimizer="lm", **args)
        assert Ro.shape == (30, 2, 3, 3) and to.shape == (30, 2, 3) and Xo.shape == (30, 17, 3) and Xo.dtype == torch.float64
        assert hist[-1]["trial_cost"] < hist[0]["cost"]
        if mode == "pose_only":
            assert torch.equal(Ro.cpu(), torch.from_numpy(R)) and torch.equal(to.cpu(), torch.from_numpy(t))
        if mode == "pose_cam_t":
            assert torch.equal(Ro.cpu(), torch.from_numpy(R)) and not torch.equal(to.cpu(), torch.from_numpy(t))


@pytest.mark.parametrize("rig,T,J,mode", [("2b", 1, 17, "full"), ("2b", 2, 5, "pose_cam_t"), ("2b", 7, 96, "pose_only"), ("3", 5, 17, "full"),
                                          ("8", 3, 96, "full")])
def test_edge_shapes(cuda, rig, T, J, mode):
    """One frame (no temporal / smoothness pairs), a skeleton too small for any bone (J = 5), the largest skeleton (J = 96:
    three joints per lane), three cameras, and the largest frame the kernels take (96 joints x 8 free cameras: 60 KB of
    shared memory per warp, two warps per block): trajectories against the exact-solve oracle."""
    clip, R, t, X0 = lm_reg.make_problem(rig, T, J, cam_jitter=0.01)
    s = ba_reg.RegularisedBundleAdjuster(torch.from_numpy(clip.x_fm).to(cuda), torch.from_numpy(clip.conf_fm).to(cuda), clip.K, R, t, X0,
                                         mode=mode, max_iters=8, cg_iters=120)
    s.run(4)
    _, _, _, ref = lm_reg.run_lm(X0, R, t, clip.K, clip.x_fm.astype(float), clip.conf_fm.astype(float), num_iters=4, mode=mode)
    assert _check(s.history, ref, 4) >= 2


def test_unobserved_points_and_input_checks(cuda):
    """A joint nobody observes (all confidences 0) has no reprojection rows: its point block comes from the regularisers
    alone and the solve stays finite and on the oracle's trajectory.  Bad arguments raise before anything is launched."""
    clip, R, t, X0 = lm_reg.make_problem("2b", 12, 17, cam_jitter=0.01)
    conf = clip.conf_fm.copy()
    conf[:, :, 3] = 0.0
    conf[5] = 0.0
    s = ba_reg.RegularisedBundleAdjuster(torch.from_numpy(clip.x_fm).to(cuda), torch.from_numpy(conf).to(cuda), clip.K, R, t, X0, mode="pose_only",
                                         max_iters=8, cg_iters=60)
    s.run(4)
    _, _, _, ref = lm_reg.run_lm(X0, R, t, clip.K, clip.x_fm.astype(float), conf.astype(float), num_iters=4, mode="pose_only")
    assert _check(s.history, ref, 4) >= 2 and torch.isfinite(s.X).all()
    x = torch.from_numpy(clip.x_fm).to(cuda)
    c = torch.from_numpy(clip.conf_fm).to(cuda)
    with pytest.raises(ValueError):
        ba_reg.RegularisedBundleAdjuster(x, c, clip.K, R, t, X0, mode="bogus")
    with pytest.raises(ValueError):
        ba_reg.RegularisedBundleAdjuster(x, c[:, :1], clip.K, R, t, X0)
    with pytest.raises(ValueError):
        ba_reg.RegularisedBundleAdjuster(x, c, clip.K, R[:5], t[:5], X0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ba_reg.RegularisedBundleAdjuster(x.cpu(), c.cpu(), clip.K, R, t, X0)
    s2 = ba_reg.RegularisedBundleAdjuster(x, c, clip.K, R, t, X0, max_iters=2)
    with pytest.raises(ValueError, match="history"):
        s2.run(3)


@pytest.mark.parametrize("name", ["c3_pose", "c3_full", "c5_pose"])
def test_bitwise_repeatable(cuda, name):
    """Every reduction is a fixed-order sum (per-warp partial rows, column reduction): two runs are bit-identical - which a
    data race between the warp-cooperative steps of a frame would break."""
    runs = []
    for _ in range(2):
        s, *_ = _problem(name, cuda)
        s.run(4)
        runs.append((s.history, s.X.clone(), s.t.clone()))
    assert runs[0][0] == runs[1][0]
    assert torch.equal(runs[0][1], runs[1][1]) and torch.equal(runs[0][2], runs[1][2])
