"""Time the post-triangulation triage + smoothing path (row N2) on 1M frames x 17 joints:  python tools/post_bench.py [T]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from skiing_analysis_pytorch_b200 import api, post, synth  # noqa: E402


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
    J = 17
    dev = torch.device("cuda:0")
    d = synth.make_clip_device("2b", T, J, dev, seed=0, layout="CTJ2")
    X = api.triangulate_reproject(d["x2d"], d["K"], d["R"], d["t"], want=("X",)).X
    K = d["K"][0]
    R, t = d["R"][1], d["t"][1]
    ms_plain = timed(lambda: post.post_triage(X, d["x2d"], K, K, R, t))
    ms_und = timed(lambda: post.post_triage(X, d["x2d"], K, K, R, t, dist1=synth.DIST_CALIB, dist2=synth.DIST_CALIB, conf=d["conf"]))
    ms_sg = timed(lambda: post.smooth_skeleton(X, 9, 2))
    N = T * J
    print(f"post_triage plain: {ms_plain:.3f} ms ({N * 45 / ms_plain / 1e6:.0f} GB/s)   with undistortion + conf: {ms_und:.3f} ms "
          f"({N * 53 / ms_und / 1e6:.0f} GB/s)   savgol: {ms_sg:.3f} ms")


if __name__ == "__main__":
    main()
