"""Time the per-frame-extrinsics triangulation path (process_triangulate's, row a2) on 1M frames x 17 joints x 2 views."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from skiing_analysis_pytorch_b200 import api, synth  # noqa: E402


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    dev = torch.device("cuda:0")
    d = synth.make_clip_device("2b", T, 17, dev, seed=0, layout="CTJ2")
    R = torch.tensor(np.broadcast_to(d["R"][None], (T, 2, 3, 3)).copy(), device=dev)
    t = torch.tensor(np.broadcast_to(d["t"][None], (T, 2, 3)).copy(), device=dev)
    fn = lambda: api.triangulate_reproject(d["x2d"], d["K"], R, t, dist=synth.DIST_CALIB, want=("X", "err"))
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"per-frame extrinsics, T={T}: {e0.elapsed_time(e1) / 10:.3f} ms per clip")


if __name__ == "__main__":
    main()
