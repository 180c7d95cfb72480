"""Time the two-view fusion + EMA kernels (row N3) on 1M frames x 70 joints, next to the reference's per-frame numpy
path:  python tools/fusion_bench.py [T] [J]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from skiing_analysis_pytorch_b200 import fusion, synth  # noqa: E402


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
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    J = int(sys.argv[2]) if len(sys.argv) > 2 else 70
    dev = torch.device("cuda:0")
    small = synth.make_fusion_clip(2000, J, seed=0)
    reps = (T + 1999) // 2000
    d = {k: torch.from_numpy(v).to(dev).repeat(reps, 1, 1)[:T].contiguous() for k, v in small.items()}
    ms_fuse = timed(lambda: fusion.fuse_clip(d["Xl"], d["Xr"], d["Ul"], d["Ur"], want=(), strict=False))
    ms_old = timed(lambda: fusion.fuse_clip(d["Xl"], d["Xr"], d["Ul"], d["Ur"], want=(), strict=False, per_frame_kernel=True), n=3)
    fused = fusion.fuse_clip(d["Xl"], d["Xr"], d["Ul"], d["Ur"], want=(), strict=False).fused
    ms_ema = timed(lambda: fusion.temporal_smooth_ema(fused))
    ms_seq = timed(lambda: fusion.temporal_smooth_ema(fused, exact=True), n=2) if T <= 200_000 else float("nan")
    ms_rigid = timed(lambda: fusion.rigid_fuse_clip(d["Xl"], d["Xr"], strict=False)) if J >= 70 else float("nan")
    b_fuse = T * J * (24 + 24 + 16 + 16 + 24)
    halo = fusion.ema_halo(0.7, True, 0.45, 0.92)
    b_ema = T * J * 48
    print(f"fuse_frames (warp-per-frame kernel): {ms_old:.3f} ms")
    print(f"rigid_fuse (rigid_transform_3D): {ms_rigid:.3f} ms  alg {T * J * 72 / 1e9:.2f} GB -> {T * J * 72 / ms_rigid / 1e6:.0f} GB/s")
    print(f"fuse_frames: T={T} J={J}  {ms_fuse:.3f} ms  {T / ms_fuse * 1e3:.3e} frames/s  alg {b_fuse / 1e9:.2f} GB -> {b_fuse / ms_fuse / 1e6:.0f} GB/s")
    print(f"ema (chunk 512, halo {halo}): {ms_ema:.3f} ms  {T / ms_ema * 1e3:.3e} frames/s  alg {b_ema / 1e9:.2f} GB -> {b_ema / ms_ema / 1e6:.0f} GB/s"
          f"   sequential scan: {ms_seq:.1f} ms")
    # (the CPU baseline of this path is timed by bench.py's `fusion.cpu_baseline` leg)


if __name__ == "__main__":
    main()
