"""Time the fused triangulate+reproject kernel on the headline shapes with CUDA events (one process, one GPU):
config 2 (1M x 17 x 2 views, distortion scoring), the same with confidences, the north star's 8-view shape
(1M x 17 x 8, confidence weighted), config 4's shard (500k x 70 x 8) and process_triangulate's per-frame-extrinsics path.
SKA_LIB_PATH selects the build (tools/variants.py)."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from skiing_analysis_pytorch_b200 import api, synth  # noqa: E402

PEAK = 6542.1
try:
    PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass


def timed(fn, n=20):
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


def shape(rig, T, J, conf, dev, label, n=20, frames=False):
    d = synth.make_clip_device(rig, T, J, dev, seed=0, layout="CTJ2")
    V = d["x2d"].shape[0]
    outs = {"X": torch.empty((T, J, 3), dtype=torch.float32, device=dev), "err": torch.empty((V, T, J), dtype=torch.float32, device=dev)}
    kw = dict(K=d["K"], R=d["R"], t=d["t"], dist=synth.DIST_CALIB, want=("X", "err"), out=outs)
    cf = d["conf"] if conf else None
    bpj = 8 * V + (4 * V if conf else 0) + 12 + 4 * V
    if frames:
        Rf = np.broadcast_to(d["R"], (T, V, 3, 3)).copy()
        tf = np.broadcast_to(d["t"], (T, V, 3)).copy()
        kw.update(R=api.pack_frame_extrinsics(Rf, tf, dev), t=None)  # packed once per clip, as a pipeline would
        bpj += V * 96 / J
    ms = timed(lambda: api.triangulate_reproject(d["x2d"], conf=cf, **kw), n)
    gbs = bpj * T * J / ms / 1e6
    print(f"{label:28s} {ms:8.4f} ms  {gbs:7.0f} GB/s  frac {gbs / PEAK:.3f}", flush=True)
    del d, outs
    torch.cuda.empty_cache()
    return ms


def main():
    dev = torch.device("cuda:0")
    which = sys.argv[1:] or ["c2", "c2conf", "v8", "c4", "frames"]
    if "c2" in which:
        shape("2b", 1_000_000, 17, False, dev, "config2 1Mx17x2")
    if "c2conf" in which:
        shape("2b", 1_000_000, 17, True, dev, "config2+conf 1Mx17x2")
    if "v4" in which:
        shape("4", 1_000_000, 17, True, dev, "4 views conf 1Mx17x4")
    if "v8" in which:
        shape("8", 1_000_000, 17, True, dev, "north-star 1Mx17x8 conf", n=10)
    if "c4" in which:
        shape("8", 500_000, 70, True, dev, "config4 shard 500kx70x8", n=5)
    if "frames" in which:
        shape("2b", 1_000_000, 17, False, dev, "per-frame [R|t] 1Mx17x2", n=10, frames=True)


if __name__ == "__main__":
    main()
