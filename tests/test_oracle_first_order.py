"""The first-order regularised BA oracle (oracle/first_order.py, row N1): golden G9 was produced with the REFERENCE's own
bundle_adjustment/loss.py as the loss module; here the portable torch restatement (oracle/torch_ref.py) must walk the same
trajectory, the tangent-space rotation gradient must equal autograd through the exponential map, and the Adam step must
be torch.optim.Adam's."""
import numpy as np
import pytest
import torch

from oracle import first_order as FO
from oracle import torch_ref as TR


RefNames = TR.ReferenceNames


@pytest.mark.parametrize("mode", FO.MODES)
def test_trajectory_matches_reference_loss_golden(golden, mode):
    g = golden("g9_first_order.npz")
    R, t, X, hist = FO.run_adam(RefNames, g["K"], g["R0"], g["t0"], g["X0"], g["x2d"], g["conf"], num_iters=25, lr=1e-2, mode=mode)
    got = np.array([[h["loss"]] + [h[k] for k in FO.TERMS] for h in hist])
    np.testing.assert_allclose(got, g[f"{mode}_hist"], rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(X.numpy(), g[f"{mode}_X"], atol=1e-9)
    np.testing.assert_allclose(R.numpy(), g[f"{mode}_R"], atol=1e-10)
    assert got[-1, 0] < got[0, 0]
    if mode == "pose_only":
        np.testing.assert_array_equal(R.numpy(), g["R0"])
        np.testing.assert_array_equal(t.numpy(), g["t0"])
    if mode == "full":
        RtR = R @ R.transpose(-1, -2)
        assert (RtR - torch.eye(3, dtype=torch.float64)).abs().max() < 1e-12  # the retraction stays on SO(3)


def test_tangent_gradient_equals_autograd_through_the_exponential_map(golden):
    g = golden("g9_first_order.npz")
    f64 = lambda a: torch.tensor(a, dtype=torch.float64)
    R, t, X, K, x2d, cf = (f64(g[k]) for k in ("R0", "t0", "X0", "K", "x2d", "conf"))
    w = torch.zeros(R.shape[:2] + (3,), dtype=torch.float64, requires_grad=True)
    z = torch.zeros_like(w[..., 0])
    W = torch.stack([torch.stack([z, -w[..., 2], w[..., 1]], -1), torch.stack([w[..., 2], z, -w[..., 0]], -1),
                     torch.stack([-w[..., 1], w[..., 0], z], -1)], -2)
    FO.total_loss(RefNames, X, torch.matrix_exp(W) @ R, t, K, x2d, cf, FO.DEFAULT_WEIGHTS)[0].backward()
    Rv = R.clone().requires_grad_(True)
    FO.total_loss(RefNames, X, Rv, t, K, x2d, cf, FO.DEFAULT_WEIGHTS)[0].backward()
    np.testing.assert_allclose(FO.tangent_grad(Rv.grad, R).numpy(), w.grad.numpy(), rtol=1e-10, atol=1e-10)


def test_adam_is_torch_optim_adam(golden):
    g = golden("g9_first_order.npz")
    f64 = lambda a: torch.tensor(a, dtype=torch.float64)
    R, t, K, x2d, cf = (f64(g[k]) for k in ("R0", "t0", "K", "x2d", "conf"))
    Xp = f64(g["X0"]).requires_grad_(True)
    opt = torch.optim.Adam([Xp], lr=1e-2)
    for _ in range(25):
        opt.zero_grad()
        FO.total_loss(RefNames, Xp, R, t, K, x2d, cf, FO.DEFAULT_WEIGHTS)[0].backward()
        opt.step()
    np.testing.assert_allclose(Xp.detach().numpy(), g["pose_only_X"], atol=1e-12)
