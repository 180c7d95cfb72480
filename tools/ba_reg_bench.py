"""Time the regularised LM (row N1) kernel by kernel on config 3's shape: python tools/ba_reg_bench.py [T] [J] [rig] [mode]"""
import math
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from skiing_analysis_pytorch_b200 import _cabi, api, ba_reg, synth  # noqa: E402


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    J = int(sys.argv[2]) if len(sys.argv) > 2 else 17
    rig = sys.argv[3] if len(sys.argv) > 3 else "2b"
    modes = sys.argv[4:] or ["pose_only", "full"]
    dev = torch.device("cuda:0")
    d = synth.make_clip_device(rig, T, J, dev, seed=100)
    R0, t0 = synth.perturb_cameras(d["R"], d["t"], seed=1)
    C = len(R0)
    X0 = api.triangulate_reproject(d["x2d"].permute(1, 0, 2, 3).contiguous(), d["K"], R0, t0, want=("X",)).X.double()
    tg = torch.arange(T, dtype=torch.float64, device=dev)
    drift = torch.stack([0.01 * torch.sin(2 * math.pi * tg / 200.0 + c) for c in range(C)], 1)[..., None] * torch.tensor([1.0, 0.5, 0.25], dtype=torch.float64, device=dev)
    R = torch.tensor(R0, device=dev)[None].expand(T, C, 3, 3).contiguous()
    t = (torch.tensor(t0, device=dev)[None] + drift).contiguous()
    k = _cabi
    for mode in modes:
        s = ba_reg.RegularisedBundleAdjuster(d["x2d"], d["conf"], d["K"], R, t, X0, mode=mode, max_iters=40, cg_iters=48, local_only=True)
        s.run(2)

        def timed(fn, n=5):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n

        parts = {"linearize": timed(s.linearize), "precond(begin)": timed(lambda: s.cg(k.BA_REG_CG_BEGIN))}
        s.cg(k.BA_REG_CG_INIT)
        s.cg(k.BA_REG_CG_DIR)
        parts["matvec"] = timed(lambda: s.cg(k.BA_REG_CG_MATVEC))
        s.cg(k.BA_REG_CG_ALPHA)
        parts["update+precond"] = timed(lambda: s.cg(k.BA_REG_CG_UPDATE), 1)
        parts["dir"] = timed(lambda: s.cg(k.BA_REG_CG_DIR))
        parts["apply"] = timed(s.apply)
        parts["cost"] = timed(lambda: s.cost(1))
        s2 = ba_reg.RegularisedBundleAdjuster(d["x2d"], d["conf"], d["K"], R, t, X0, mode=mode, max_iters=40, cg_iters=48, local_only=True)
        s2.run(3, graph=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s2.run(10, graph=True)
        e1.record()
        torch.cuda.synchronize()
        h = s2.history
        print(f"{mode}: T={T} J={J} C={C}  {e0.elapsed_time(e1) / 10:.3f} ms/trial (graph, 48 CG iterations captured)  cg iters {[x['cg_iters'] for x in h]}")
        print("   parts (ms): " + "  ".join(f"{n} {v:.3f}" for n, v in parts.items()))
        print("   cost: " + " ".join(f"{x['cost']:.5f}" for x in h[:8]))
        del s, s2


if __name__ == "__main__":
    main()
