"""Time the BA kernels on the full BASELINE shapes (one GPU): python tools/ba_bench.py [c3|c5|...] [T]"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from skiing_analysis_pytorch_b200 import api, ba, synth  # noqa: E402

CASES = {"c3": ("2b", 100_000, 17), "c5": ("8", 1_000_000, 70), "c5s": ("8", 125_000, 70), "c4v": ("4", 200_000, 17)}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    rig, T, J = CASES[name]
    if len(sys.argv) > 2:
        T = int(sys.argv[2])
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    graph = "--graph" in sys.argv
    dev = torch.device("cuda:0")
    d = synth.make_clip_device(rig, T, J, dev, seed=0)
    R0, t0 = synth.perturb_cameras(d["R"], d["t"], seed=1)
    C = len(R0)
    kv = d["x2d"].permute(1, 0, 2, 3).contiguous()
    X0 = api.triangulate_reproject(kv, d["K"], R0, t0, want=("X",)).X
    del kv
    if "--adam" in sys.argv:  # first-order form of the regularised objective with per-frame cameras (row N1)
        Rf = torch.tensor(R0, device=dev)[None].expand(T, C, 3, 3).contiguous()
        tf = torch.tensor(t0, device=dev)[None].expand(T, C, 3).contiguous()
        dt = torch.float64 if "--f64" in sys.argv else torch.float32
        for mode in ba.MODES:
            args = (torch.tensor(d["K"]), Rf.to(dt), tf.to(dt), X0.to(dt), d["x2d"], d["conf"])
            use_graph = "--no-graph" not in sys.argv
            only = [a.split("=")[1] for a in sys.argv if a.startswith("--mode=")]
            if only and mode not in only:
                continue
            ba.run_local_ba(*args, num_iters=5, lr=1e-2, mode=mode, optimizer="adam", graph=use_graph)
            torch.cuda.synchronize()
            t0_ = time.perf_counter()
            _, _, _, h = ba.run_local_ba(*args, num_iters=iters, lr=1e-2, mode=mode, optimizer="adam", graph=use_graph)
            torch.cuda.synchronize()
            ms = 1e3 * (time.perf_counter() - t0_) / iters
            print(f"{name} adam {mode:10s} {str(dt)[6:]}: T={T} J={J} C={C}  {ms:.3f} ms/iter (wall incl. set-up, graph={use_graph})  loss {h[0]['loss']:.4f} -> {h[-1]['loss']:.4f}")
        return
    if "--calib" in sys.argv:  # config 3 with free intrinsics + distortion (15 parameters per camera)
        K_init, dist_init = synth.theta_to_K_dist(synth.perturb_intrinsics(d["K"]))
        X0 = api.triangulate_reproject(d["x2d"].permute(1, 0, 2, 3).contiguous(), K_init, R0, t0, want=("X",)).X
        s = ba.CalibratingBundleAdjuster(d["x2d"], d["conf"], K_init, R0, t0, X0, dist=dist_init, prior_rho=synth.CALIB_PRIOR_RHO,
                                         prior_theta=synth.theta_from_K(d["K"]), max_iters=iters + 8)
    else:
        s = ba.BundleAdjuster(d["x2d"], d["conf"], d["K"], R0, t0, X0, max_iters=iters + 8, force_wide="--wide" in sys.argv, tensor_core="--tc" in sys.argv)
    s.run(3, graph=graph)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    # per-kernel timing of one trial
    ev[0].record(); s.linearize(); ev[1].record(); s.solve(); ev[2].record(); s.backsub(); ev[3].record(); s.control(); ev[4].record()
    torch.cuda.synchronize()
    parts = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
    s.iters_done += 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0_ = time.perf_counter()
    e0.record()
    s.run(iters, graph=graph)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0_
    ms = e0.elapsed_time(e1) / iters
    N = T * J
    alg = N * (2 * 12 * C + 36)
    print(f"{name}: T={T} J={J} C={C} graph={graph}  {ms:.4f} ms/iter  {1e3 / ms:.1f} it/s  wall {1e3 * wall / iters:.4f} ms/iter  "
          f"alg {alg / 1e6:.1f} MB/iter -> {alg / ms / 1e6:.1f} GB/s   parts(ms): lin {parts[0]:.4f} solve {parts[1]:.4f} back {parts[2]:.4f} ctl {parts[3]:.4f}")
    h = s.history
    print("  cost:", " ".join(f"{r['cost']:.5f}" for r in h[:8]), " accepted:", sum(r["accepted"] for r in h), "/", len(h))


if __name__ == "__main__":
    main()
