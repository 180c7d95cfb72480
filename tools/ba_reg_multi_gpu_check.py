"""Multi-GPU check of the regularised LM (run under torchrun on N GPUs): the frame-sharded solve - one-frame
halo of the CG direction and of the trial point, all-reduced dot products / cost sums / bone + baseline means - must follow
the single-GPU trajectory, eagerly and as a replayed CUDA graph (which then contains the NCCL collectives).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ba_reg_multi_gpu_check.py
"""
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from skiing_analysis_pytorch_b200 import api, ba, ba_reg, synth  # noqa: E402


def problem(rig, T, J, dev):
    d = synth.make_clip_device(rig, T, J, dev, seed=7)  # same seed on every rank: identical clip
    R0, t0 = synth.perturb_cameras(d["R"], d["t"], seed=1)
    X0 = api.triangulate_reproject(d["x2d"].permute(1, 0, 2, 3).contiguous(), d["K"], R0, t0, want=("X",)).X
    C = len(R0)
    g = torch.Generator(device=dev).manual_seed(11)
    R = torch.tensor(R0, device=dev)[None].expand(T, C, 3, 3).contiguous()
    t = torch.tensor(t0, device=dev)[None].expand(T, C, 3) + torch.cumsum(0.002 * torch.randn(T, C, 3, generator=g, device=dev, dtype=torch.float64), 0)
    return d, R, t.contiguous(), X0.double()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    n = 6
    for rig, T, J, mode in (("2b", 20_000, 17, "pose_only"), ("2b", 20_000, 17, "full"), ("8", 2_000, 70, "pose_cam_t")):
        d, R, t, X0 = problem(rig, T, J, dev)
        a, b = ba.frame_shard(T, world, rank)
        for graph in (False, True):
            s = ba_reg.RegularisedBundleAdjuster(d["x2d"][a:b].contiguous(), d["conf"][a:b].contiguous(), d["K"], R[a:b], t[a:b], X0[a:b],
                                                 mode=mode, max_iters=n + 2, cg_iters=48, group=dist.group.WORLD)
            torch.cuda.synchronize()
            dist.barrier()
            t0_ = time.perf_counter()
            s.run(n, graph=graph)
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0_) / n
            h = s.history
            if rank == 0:
                ref = ba_reg.RegularisedBundleAdjuster(d["x2d"], d["conf"], d["K"], R, t, X0, mode=mode, max_iters=n + 2, cg_iters=48, local_only=True)
                ref.run(n)
                hr = ref.history
                dev_cost = max(abs(x["trial_cost"] - y["trial_cost"]) / y["trial_cost"] for x, y in zip(h, hr))
                same = [x["accepted"] == y["accepted"] for x, y in zip(h, hr) if abs(y["cost"] - y["trial_cost"]) > 1e-9 * y["cost"]]
                dx = float((s.X - ref.X[a:b]).abs().max())
                print(f"{rig} T={T} J={J} {mode:10s} world={world} graph={graph}: trial-cost dev {dev_cost:.2e}, decisions equal {all(same)}, "
                      f"cost {hr[0]['cost']:.5f} -> {hr[-1]['trial_cost']:.5f} | sharded {h[-1]['trial_cost']:.5f}; max |dX| {dx:.2e}; "
                      f"cg iters {[x['cg_iters'] for x in h]}; {1e3 * wall:.2f} ms/trial", flush=True)
                ok = ok and dev_cost < 1e-9 and all(same) and dx < 1e-7
                del ref
            del s
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_BA_REG_OK" if flag.item() == 1.0 else "MULTI_GPU_BA_REG_FAILED")
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
