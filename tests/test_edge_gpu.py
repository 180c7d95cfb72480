"""Edge cases of the standalone entry points on the GPU: empty / single-frame / single-joint inputs,
zero confidences, NaN propagation - compared with what the reference's expressions give (evaluated with
the plain torch restatement, oracle/torch_ref.py)."""
import math

import numpy as np
import pytest
import torch

from skiing_analysis_pytorch_b200 import api, ba, losses as L, synth
from oracle import torch_ref as TR

pytestmark = pytest.mark.gpu


def test_losses_on_degenerate_shapes(cuda):
    d = torch.float64
    X1 = torch.randn(1, 17, 3, dtype=d, device=cuda)
    # a single frame: mean over an empty difference tensor is nan in torch, and here
    assert math.isnan(L.pose_temporal_loss(X1).item()) and math.isnan(TR.pose_temporal(X1, 1e-2).item())
    assert L.bone_length_loss(X1).item() == pytest.approx(TR.bone_length(X1, None, 1e-2).item(), abs=1e-15)
    R = torch.eye(3, dtype=d, device=cuda).expand(1, 2, 3, 3).contiguous()
    t = torch.randn(1, 2, 3, dtype=d, device=cuda)
    assert math.isnan(L.camera_smooth_loss(R, t).item())
    assert L.baseline_reg_loss(R, t).item() == pytest.approx(0.0, abs=1e-30)
    # zero confidences: 0 / (0 + 1e-6) = 0, no NaN
    Xt = torch.randn(3, 5, 3, dtype=d, device=cuda) + torch.tensor([0.0, 0.0, 10.0], dtype=d, device=cuda)
    Rc = torch.eye(3, dtype=d, device=cuda).expand(2, 3, 3).contiguous()
    tc = torch.zeros(2, 3, dtype=d, device=cuda)
    K = torch.tensor(synth.K_CALIB, dtype=d, device=cuda).expand(2, 3, 3).contiguous()
    x2d = torch.zeros(3, 2, 5, 2, dtype=d, device=cuda)
    assert L.reprojection_loss(Xt, Rc, tc, K, x2d, torch.zeros(3, 2, 5, dtype=d, device=cuda)).item() == 0.0
    # one joint, one frame, 2-D X3d input
    p = L.project_points(Xt[0, :1].reshape(1, 3), Rc, tc, K)
    assert tuple(p.shape) == (1, 2, 1, 2)
    assert torch.allclose(p, TR.project_points(Xt[0, :1].reshape(1, 3), Rc, tc, K), atol=1e-9)
    # NaN joints propagate into the loss exactly like torch
    Xn = Xt.clone()
    Xn[1, 2, 0] = float("nan")
    assert math.isnan(L.reprojection_loss(Xn, Rc, tc, K, x2d, torch.ones(3, 2, 5, dtype=d, device=cuda)).item())
    # empty clip
    X0 = torch.zeros(0, 5, 3, dtype=d, device=cuda)
    assert tuple(L.project_points(X0, Rc, tc, K).shape) == (0, 2, 5, 2)
    assert L.reprojection_loss(X0, Rc, tc, K, x2d[:0], torch.zeros(0, 2, 5, dtype=d, device=cuda)).item() == 0.0


def test_reprojection_and_stats_edges(cuda):
    R, t = synth.rig("2b")
    X = torch.zeros(0, 17, 3, device=cuda)
    proj, err = api.reproject_points(X, synth.K_CALIB, R, t, None, kpts=torch.zeros(2, 0, 17, 2, device=cuda), want=("proj", "err"))
    assert tuple(proj.shape) == (2, 0, 17, 2) and tuple(err.shape) == (2, 0, 17)
    e = torch.full((2, 3, 1), float("nan"), device=cuda)
    e[0, 1, 0] = 2.5
    st = api.frame_stats(e).cpu().numpy()
    assert np.isnan(st[0]).all() and np.isnan(st[2]).all() and np.isnan(st[1, 1]).all()
    np.testing.assert_array_equal(st[1, 0], [2.5, 2.5, 2.5, 2.5])  # one value: rmse = mean = median = max
    # z = 0 follows cv2 (1/z replaced by 1), a point behind the camera projects through the mirror image: no NaN
    Xz = torch.tensor([[[1.0, 2.0, 0.0], [0.5, 0.5, -4.0]]], device=cuda)
    p, _ = api.reproject_points(Xz, synth.K_CALIB, np.eye(3)[None], np.zeros((1, 3)), None)
    assert torch.isfinite(p).all()
    assert p[0, 0, 0, 0].item() == pytest.approx(synth.K_CALIB[0, 0] * 1.0 + synth.K_CALIB[0, 2], rel=1e-6)


@pytest.mark.parametrize("J", [1, 5, 16, 17, 18, 25, 32, 33, 70])
def test_frame_stats_all_kernel_paths_match_numpy(cuda, J):
    """J <= 17 and J <= 32 run the thread-per-row register kernels, larger skeletons the warp-per-row kernel: all three
    must give numpy's nan-aware rmse / mean / median / max (triangulation/reproject.py:254-261), ties and NaNs included."""
    rng = np.random.default_rng(J)
    e = rng.uniform(0, 5, (2, 300, J)).astype(np.float32)
    e[rng.random(e.shape) < 0.2] = np.nan
    e[0, 7] = np.nan                      # an empty row
    e[1, 9] = 1.25                        # all ties
    if J >= 5:
        e[0, 11, :4] = 2.0                # partial ties around the median
    st = api.frame_stats(torch.from_numpy(e).to(cuda)).cpu().numpy()  # (T,V,4)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = np.stack([np.sqrt(np.nanmean(e.astype(np.float64) ** 2, -1)), np.nanmean(e.astype(np.float64), -1),
                        np.nanmedian(e.astype(np.float64), -1), np.nanmax(e, -1)], -1).transpose(1, 0, 2)
    np.testing.assert_allclose(st, ref, rtol=2e-7, atol=0, equal_nan=True)


def test_ba_with_unobserved_points_and_frozen_cameras(cuda):
    """conf = 0 for a whole point (nobody observes it): its block is singular, it must not move and must not
    poison the reduced system; mode='pose_only' keeps every camera fixed."""
    from oracle import lm

    clip, R0, t0, X0 = lm.make_problem("3", 40, 17)
    conf = clip.conf_fm.copy()
    conf[5, :, 3] = 0.0
    conf[17, :, :] = 0.0
    x = torch.from_numpy(clip.x_fm).to(cuda)
    c = torch.from_numpy(conf).to(cuda)
    X = torch.from_numpy(X0.astype(np.float32)).to(cuda)
    s = ba.ba_solve(x, c, clip.K, R0, t0, X, num_iters=6)
    _, _, Xo, ho = lm.run_lm(X0, R0, t0, clip.K, clip.x_fm, conf, num_iters=6)
    for h, o in zip(s.history[:4], ho[:4]):
        assert abs(h["cost"] - o["cost"]) <= 1e-4 * o["cost"]
    Xs = s.X.cpu().numpy()
    np.testing.assert_array_equal(Xs[5, 3], X0[5, 3].astype(np.float32))
    np.testing.assert_array_equal(Xs[17], X0[17].astype(np.float32))
    assert np.isfinite(Xs).all() and np.isfinite(s.R).all()
    f = ba.ba_solve(x, c, clip.K, R0, t0, X, num_iters=4, mode="pose_only")
    np.testing.assert_allclose(f.R, R0, atol=1e-15)
    np.testing.assert_allclose(f.t, t0, atol=1e-15)
    assert f.cost < f.history[0]["cost"]
