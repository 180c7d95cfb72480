"""Time the secondary kernels of the library at full clip sizes (CUDA events) to spot the ones far from their bound."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from skiing_analysis_pytorch_b200 import api, losses, synth  # noqa: E402


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda:0")
    T, J = 1_000_000, 17
    d = synth.make_clip_device("2b", T, J, dev, seed=0, layout="CTJ2")
    X = api.triangulate_reproject(d["x2d"], d["K"], d["R"], d["t"], want=("X",)).X
    N = T * J
    ms = timed(lambda: api.reproject_points(X, d["K"], d["R"], d["t"], synth.DIST_CALIB, kpts=d["x2d"], want=("proj", "err")))
    print(f"reproject_points (cv2 model, fp64) proj+err 2 views: {ms:.3f} ms  {N * (12 + 2 * (8 + 8 + 4)) / ms / 1e6:.0f} GB/s")
    err = api.reproject_points(X, d["K"], d["R"], d["t"], synth.DIST_CALIB, kpts=d["x2d"], want=("err",))[-1]
    ms = timed(lambda: api.frame_stats(err))
    print(f"frame_stats (2,T,17): {ms:.3f} ms  {err.numel() * 4 / ms / 1e6:.0f} GB/s")
    # loss.py kernels at config-3 size, shared and per-frame cameras, forward and forward+backward
    Tc = 100_000
    x2d = d["x2d"][:, :Tc].permute(1, 0, 2, 3).contiguous()
    cf = d["conf"][:, :Tc].permute(1, 0, 2).contiguous()
    Xc = X[:Tc].contiguous()
    f32 = lambda a: torch.tensor(a, dtype=torch.float32, device=dev)
    R, t, K = f32(d["R"]), f32(d["t"]), f32(d["K"])
    obs = Tc * J * 2
    ms = timed(lambda: losses.reprojection_loss(Xc, R, t, K, x2d, cf))
    print(f"reprojection_loss fwd (100k x 17 x 2, shared cams): {ms:.3f} ms  {obs * 12 / ms / 1e6:.0f} GB/s")
    Xg, Rg, tg = Xc.clone().requires_grad_(True), R.clone().requires_grad_(True), t.clone().requires_grad_(True)

    def fb():
        Xg.grad = Rg.grad = tg.grad = None
        losses.reprojection_loss(Xg, Rg, tg, K, x2d, cf).backward()
    ms = timed(fb)
    print(f"reprojection_loss fwd+bwd shared cams: {ms:.3f} ms")
    Rf, tf = R[None].expand(Tc, 2, 3, 3).contiguous().requires_grad_(True), t[None].expand(Tc, 2, 3).contiguous().requires_grad_(True)

    def fbf():
        Xg.grad = Rf.grad = tf.grad = None
        losses.reprojection_loss(Xg, Rf, tf, K, x2d, cf).backward()
    ms = timed(fbf)
    print(f"reprojection_loss fwd+bwd per-frame cams: {ms:.3f} ms")
    ms = timed(lambda: losses.project_points(Xc, R, t, K))
    print(f"project_points (100k x 17 x 2): {ms:.3f} ms  {Tc * J * (12 + 16) / ms / 1e6:.0f} GB/s")
    for name, fn in (("bone_length", lambda: losses.bone_length_loss(Xc)), ("pose_temporal", lambda: losses.pose_temporal_loss(Xc)),
                     ("camera_smooth", lambda: losses.camera_smooth_loss(Rf.detach(), tf.detach())),
                     ("baseline_reg", lambda: losses.baseline_reg_loss(Rf.detach(), tf.detach()))):
        print(f"{name} fwd (100k frames): {timed(fn):.3f} ms")
    X1 = X
    print(f"bone_length fwd (1M frames): {timed(lambda: losses.bone_length_loss(X1)):.3f} ms  {N * 12 / 1e6:.0f} MB in")
    print(f"pose_temporal fwd (1M frames): {timed(lambda: losses.pose_temporal_loss(X1)):.3f} ms")


if __name__ == "__main__":
    main()
