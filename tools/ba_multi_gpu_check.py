"""Multi-GPU BA check (run under torchrun on N GPUs): the frame-sharded LM (NCCL all-reduce of the packed
reduced camera system per trial) must follow the single-GPU trajectory.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ba_multi_gpu_check.py
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from skiing_analysis_pytorch_b200 import api, ba, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for rig, T, J in (("2b", 20_000, 17), ("8", 4_000, 70)):
        d = synth.make_clip_device(rig, T, J, dev, seed=7)  # same seed on every rank: identical clip
        R0, t0 = synth.perturb_cameras(d["R"], d["t"], seed=1)
        kv = d["x2d"].permute(1, 0, 2, 3).contiguous()
        X0 = api.triangulate_reproject(kv, d["K"], R0, t0, want=("X",)).X
        a, b = ba.frame_shard(T, world, rank)
        s = ba.BundleAdjuster(d["x2d"][a:b].contiguous(), d["conf"][a:b].contiguous(), d["K"], R0, t0, X0[a:b].contiguous(),
                              max_iters=12, group=dist.group.WORLD)
        s.run(10, graph=True)  # the captured trial contains the two NCCL all-reduces
        h = s.history
        if rank == 0:
            ref = ba.BundleAdjuster(d["x2d"], d["conf"], d["K"], R0, t0, X0, max_iters=12, local_only=True)
            ref.run(10)
            hr = ref.history
            dev_cost = max(abs(x["cost"] - y["cost"]) / y["cost"] for x, y in zip(h, hr))
            dev_trial = max(abs(x["trial_cost"] - y["trial_cost"]) / y["trial_cost"] for x, y in zip(h[:5], hr[:5]))
            # decisions are compared while they are decisive (at the optimum F_trial - F is rounding noise)
            same = [x["accepted"] == y["accepted"] for x, y in zip(h, hr) if abs(y["cost"] - y["trial_cost"]) > 1e-3 * y["cost"]]
            print(f"rig {rig}: world={world} cost dev {dev_cost:.2e} trial dev {dev_trial:.2e} decisions equal {all(same)} "
                  f"cost {hr[0]['cost']:.5f} -> {ref.cost:.5f} | sharded {s.cost:.5f}; R dev {np.abs(s.R - ref.R).max():.2e}")
            ok = ok and dev_cost < 1e-4 and dev_trial < 1e-4 and all(same)
        del s
    # the calibrating BA (15 parameters per camera) sharded by frames: same check against the single-GPU trajectory
    d = synth.make_clip_device("2b", 20_000, 17, dev, seed=7)
    R0, t0 = synth.perturb_cameras(d["R"], d["t"], seed=1)
    K_init, dist_init = synth.theta_to_K_dist(synth.perturb_intrinsics(d["K"]))
    X0 = api.triangulate_reproject(d["x2d"].permute(1, 0, 2, 3).contiguous(), K_init, R0, t0, want=("X",)).X
    kw = dict(dist=dist_init, prior_rho=synth.CALIB_PRIOR_RHO, prior_theta=synth.theta_from_K(d["K"]), max_iters=12)
    a, b = ba.frame_shard(20_000, world, rank)
    s = ba.CalibratingBundleAdjuster(d["x2d"][a:b].contiguous(), d["conf"][a:b].contiguous(), K_init, R0, t0, X0[a:b].contiguous(),
                                     group=dist.group.WORLD, **kw)
    s.run(10, graph=True)
    h = s.history
    if rank == 0:
        ref = ba.CalibratingBundleAdjuster(d["x2d"], d["conf"], K_init, R0, t0, X0, local_only=True, **kw)
        ref.run(10)
        hr = ref.history
        dev_cost = max(abs(x["cost"] - y["cost"]) / y["cost"] for x, y in zip(h, hr))
        same = [x["accepted"] == y["accepted"] for x, y in zip(h, hr) if abs(y["cost"] - y["trial_cost"]) > 1e-3 * y["cost"]]
        print(f"calibrating BA: world={world} cost dev {dev_cost:.2e} decisions equal {all(same)} cost {hr[0]['cost']:.5f} -> {ref.cost:.5f} "
              f"| sharded {s.cost:.5f}; theta dev {np.abs(s.theta - ref.theta).max():.2e}")
        ok = ok and dev_cost < 1e-4 and all(same)
    del s
    # every rank holds the same cameras / decisions
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_BA_OK" if flag.item() == 1.0 else "MULTI_GPU_BA_FAILED")
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
