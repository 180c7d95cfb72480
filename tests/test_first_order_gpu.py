"""GPU parity of the first-order (Adam) regularised bundle adjustment (row N1, first-order form: ba.run_local_ba with
optimizer="adam") against golden G9 - oracle/first_order.py driven by the REFERENCE's own bundle_adjustment/loss.py.
float64 on the GPU: loss trajectory 1e-8 relative per iteration; float32: 1e-4 (north-star LM tolerance) over the first
iterations."""
import numpy as np
import pytest
import torch

from skiing_analysis_pytorch_b200 import _lib, ba

pytestmark = pytest.mark.gpu
TERMS = ba.FIRST_ORDER_TERMS


def _args(g, dtype):
    t = lambda k: torch.tensor(g[k], dtype=dtype)
    return dict(K_torch=t("K"), R_init_torch=t("R0"), t_init_torch=t("t0"), X3d_init_torch=t("X0"), x2d_torch=t("x2d").float(),
                conf2d_torch=t("conf").float())


@pytest.mark.parametrize("mode", ba.MODES)
def test_f64_trajectory_matches_reference_loss_golden(cuda, golden, mode):
    g = golden("g9_first_order.npz")
    R, t, X, hist = ba.run_local_ba(**_args(g, torch.float64), num_iters=25, lr=1e-2, device="cuda", mode=mode, optimizer="adam")
    got = np.array([[h["loss"]] + [h[k] for k in TERMS] for h in hist])
    np.testing.assert_allclose(got, g[f"{mode}_hist"], rtol=1e-8, atol=1e-13)
    assert R.dtype == torch.float64 and X.shape == g["X0"].shape
    np.testing.assert_allclose(X.cpu().numpy(), g[f"{mode}_X"], atol=1e-8)
    np.testing.assert_allclose(t.cpu().numpy(), g[f"{mode}_t"], atol=1e-8)
    np.testing.assert_allclose(R.cpu().numpy(), g[f"{mode}_R"], atol=1e-9)


def test_graph_replay_equals_eager(cuda, golden):
    g = golden("g9_first_order.npz")
    a = ba.run_local_ba(**_args(g, torch.float64), num_iters=25, lr=1e-2, device="cuda", mode="full", optimizer="adam", graph=False)
    b = ba.run_local_ba(**_args(g, torch.float64), num_iters=25, lr=1e-2, device="cuda", mode="full", optimizer="adam", graph=True)
    assert [h["loss"] for h in a[3]] == [h["loss"] for h in b[3]]      # same kernels in the same order: bit-identical
    assert torch.equal(a[2], b[2]) and torch.equal(a[0], b[0])


@pytest.mark.parametrize("mode", ba.MODES)
def test_fused_iteration_equals_autograd_iteration(cuda, golden, mode):
    """The product iteration (raw loss kernels, weights folded into the multi-term Adam kernel) against the same
    optimiser driven through the autograd wrappers of losses.py."""
    g = golden("g9_first_order.npz")
    a = ba.run_local_ba(**_args(g, torch.float64), num_iters=20, lr=1e-2, device="cuda", mode=mode, optimizer="adam", fused=True)
    b = ba.run_local_ba(**_args(g, torch.float64), num_iters=20, lr=1e-2, device="cuda", mode=mode, optimizer="adam", fused=False, graph=False)
    np.testing.assert_allclose([h["loss"] for h in a[3]], [h["loss"] for h in b[3]], rtol=1e-10)
    for k in TERMS:
        np.testing.assert_allclose([h[k] for h in a[3]], [h[k] for h in b[3]], rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(a[2].cpu().numpy(), b[2].cpu().numpy(), atol=1e-10)
    np.testing.assert_allclose(a[0].cpu().numpy(), b[0].cpu().numpy(), atol=1e-11)


def test_f32_trajectory_within_tolerance(cuda, golden):
    g = golden("g9_first_order.npz")
    R, t, X, hist = ba.run_local_ba(**_args(g, torch.float32), num_iters=12, lr=1e-2, device="cuda", mode="pose_cam_t", optimizer="adam")
    got = np.array([h["loss"] for h in hist])
    np.testing.assert_allclose(got, g["pose_cam_t_hist"][:12, 0], rtol=1e-4)
    assert X.dtype == torch.float32


def test_update_kernels(cuda):
    lib = _lib.load()
    import ctypes as C
    rng = np.random.default_rng(0)
    n = 1000
    p, gr = rng.normal(size=n), rng.normal(size=n)
    m, v = rng.normal(size=n) * 0.1, rng.random(n) * 0.1
    P, G, M, V = (torch.tensor(a, device=cuda) for a in (p, gr, m, v))
    k, lr, b1, b2, eps = 3, 1e-2, 0.9, 0.999, 1e-8
    ptr = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(lib.ska_adam_step_f64(ptr(P), ptr(G), ptr(M), ptr(V), n, lr / (1 - b1**k), b1, b2, eps, 1 / (1 - b2**k) ** 0.5, None, None, None))
    m2 = m + (1 - b1) * (gr - m)
    v2 = b2 * v + (1 - b2) * gr * gr
    np.testing.assert_allclose(P.cpu().numpy(), p - (lr / (1 - b1**k)) * m2 / (np.sqrt(v2) / (1 - b2**k) ** 0.5 + eps), rtol=1e-13)
    np.testing.assert_allclose(M.cpu().numpy(), m2, rtol=1e-13)  # fma contraction of m + c (g - m)
    # retraction: exp(-s) R stays orthonormal and matches scipy-free Rodrigues
    from oracle import geometry as Gm
    R0 = np.stack([Gm.so3_exp(rng.normal(size=3)) for _ in range(50)])
    s = rng.normal(size=(50, 3)) * 0.3
    s[0] = 0.0
    Rd, Sd = torch.tensor(R0, device=cuda), torch.tensor(s, device=cuda)
    _lib.check(lib.ska_so3_retract_f64(ptr(Rd), ptr(Sd), 50, None))
    ref = np.stack([Gm.so3_exp(-s[i]) @ R0[i] for i in range(50)])
    np.testing.assert_allclose(Rd.cpu().numpy(), ref, atol=1e-14)


def test_arguments(cuda, golden):
    g = golden("g9_first_order.npz")
    a = _args(g, torch.float64)
    with pytest.raises(ValueError):
        ba.run_local_ba(**a, num_iters=2, mode="bogus", optimizer="adam")
    with pytest.raises(ValueError):
        ba.run_local_ba(**a, num_iters=2, optimizer="sgd")
    with pytest.raises(RuntimeError):
        ba.run_local_ba(**a, num_iters=2, device="cpu", optimizer="adam")
    b = dict(a, R_init_torch=a["R_init_torch"][0])
    with pytest.raises(ValueError):
        ba.run_local_ba(**b, num_iters=2, optimizer="adam")
    # weights: switching every regulariser off leaves the reprojection term alone
    _, _, _, h = ba.run_local_ba(**a, num_iters=2, optimizer="adam", weights=dict(smooth=0, baseline=0, bone_length=0, pose_temporal=0))
    assert h[0]["loss"] == h[0]["reproj"] and h[0]["smooth"] == 0.0


def test_full_size_monotone_decrease(cuda):
    """100k frames x 17 joints x 2 per-frame cameras: the configured objective decreases and the rotations stay on SO(3)."""
    from skiing_analysis_pytorch_b200 import synth

    T, J = 100_000, 17
    d = synth.make_clip_device("2b", T, J, cuda, seed=3)
    R0, t0 = synth.perturb_cameras(d["R"], d["t"], seed=1)
    R = torch.tensor(R0, device=cuda)[None].expand(T, 2, 3, 3).contiguous()
    t = torch.tensor(t0, device=cuda)[None].expand(T, 2, 3).contiguous()
    X0 = d["X"] + 0.05 * torch.randn(T, J, 3, dtype=torch.float64, device=cuda, generator=torch.Generator(device=cuda).manual_seed(1))
    Ro, to, Xo, h = ba.run_local_ba(torch.tensor(d["K"]), R, t, X0, d["x2d"], d["conf"], num_iters=40, lr=1e-2, device="cuda", mode="full",
                                    optimizer="adam")
    loss = [r["loss"] for r in h]
    assert loss[-1] < 0.5 * loss[0]
    assert ((Ro @ Ro.transpose(-1, -2)) - torch.eye(3, dtype=Ro.dtype, device=cuda)).abs().max() < 1e-12
