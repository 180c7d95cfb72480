"""Plain-PyTorch restatement of bundle_adjustment/loss.py:17-155 - TEST INFRASTRUCTURE (see oracle/__init__.py): the
autograd reference for the CUDA value+gradient kernels (values themselves are pinned by the reference's own outputs in
tests/golden/g3_g4_loss.npz and, through oracle/first_order.py, golden G9) and the CPU baseline of bench.py's first-order
bundle-adjustment leg.  Works on any device; the tests run it in float64."""
import torch

BONES = [(11, 13), (13, 15), (12, 14), (14, 16), (5, 7), (7, 9), (6, 8), (8, 10), (5, 6), (11, 12), (5, 11), (6, 12)]


def project_points(X, R, t, K):
    if X.dim() == 2:
        X = X[None]
    T = X.shape[0]
    if R.dim() == 3:
        R = R[None].expand(T, -1, -1, -1)
    if t.dim() == 2:
        t = t[None].expand(T, -1, -1)
    if K.dim() == 3:
        K = K[None].expand(T, -1, -1, -1)
    Xc = torch.einsum("tcab,tjb->tcja", R, X) + t[:, :, None, :]
    z = Xc[..., 2].clamp(min=1e-6)
    x, y = Xc[..., 0] / z, Xc[..., 1] / z
    u = K[:, :, None, 0, 0] * x + K[:, :, None, 0, 1] * y + K[:, :, None, 0, 2]
    v = K[:, :, None, 1, 0] * x + K[:, :, None, 1, 1] * y + K[:, :, None, 1, 2]
    return torch.stack([u, v], -1)


def reprojection_loss(X, R, t, K, x2d, conf, w=1.0):
    d = ((project_points(X, R, t, K) - x2d) ** 2).sum(-1)
    return w * (conf * d).sum() / (conf.sum() + 1e-6)


def centres(R, t):
    return -(R.transpose(-1, -2) @ t[..., None]).squeeze(-1)


def camera_smooth(R, t, w):
    C = centres(R, t)
    return w * ((C[1:] - C[:-1]) ** 2).mean()


def baseline_reg(R, t, w):
    C = centres(R, t)
    b = (C[:, 0] - C[:, 1]).norm(dim=-1)
    return w * ((b - b.mean().detach()) ** 2).mean()


def bone_length(X, ref, w):
    J = X.shape[1]
    L = torch.stack([(X[:, i] - X[:, j]).norm(dim=-1) for i, j in BONES if i < J and j < J], -1)
    r = L.mean(0, keepdim=True).detach() if ref is None else ref[None]
    return w * ((L - r) ** 2).mean()


def pose_temporal(X, w):
    return w * ((X[1:] - X[:-1]) ** 2).mean()


class ReferenceNames:
    """This module under the reference's function names and signatures (bundle_adjustment/loss.py:90-155), the interface
    oracle/first_order.py expects."""
    reprojection_loss = staticmethod(reprojection_loss)
    camera_smooth_loss = staticmethod(lambda R, t, w: camera_smooth(R, t, w))
    baseline_reg_loss = staticmethod(lambda R, t, w: baseline_reg(R, t, w))
    bone_length_loss = staticmethod(lambda X, ref, w: bone_length(X, ref, w))
    pose_temporal_loss = staticmethod(lambda X, w: pose_temporal(X, w))
