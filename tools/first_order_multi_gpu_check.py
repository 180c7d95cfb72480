"""Multi-GPU check of the first-order (Adam) form of the configured objective (run under torchrun on N GPUs): the frame-sharded
solve (one-frame halos, all-reduced sums and means) must reproduce the single-GPU loss history and iterates.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/first_order_multi_gpu_check.py
"""
import os
import sys
from pathlib import Path

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
    T, J, rig, iters = 20_000, 17, "2b", 40
    d = synth.make_clip_device(rig, T, J, dev, seed=7)
    R0, t0 = synth.perturb_cameras(d["R"], d["t"], seed=1)
    C = len(R0)
    X0 = api.triangulate_reproject(d["x2d"].permute(1, 0, 2, 3).contiguous(), d["K"], R0, t0, want=("X",)).X.double()
    g = torch.Generator(device=dev).manual_seed(11)
    R = torch.tensor(R0, device=dev)[None].expand(T, C, 3, 3).contiguous()
    t = (torch.tensor(t0, device=dev)[None].expand(T, C, 3) + torch.cumsum(0.002 * torch.randn(T, C, 3, generator=g, device=dev, dtype=torch.float64), 0)).contiguous()
    K = torch.tensor(d["K"], device=dev)
    a, b = ba.frame_shard(T, world, rank)
    for mode in ba.MODES:
        for graph in (False, True):
            Rs, ts, Xs, hs = ba.run_local_ba_first_order(K, R[a:b], t[a:b], X0[a:b], d["x2d"][a:b], d["conf"][a:b], num_iters=iters, lr=1e-2, device=dev,
                                                         mode=mode, graph=graph, group=dist.group.WORLD)
            if rank == 0:
                Rr, tr, Xr, hr = ba.run_local_ba_first_order(K, R, t, X0, d["x2d"], d["conf"], num_iters=iters, lr=1e-2, device=dev, mode=mode, graph=graph)
                dl = max(abs(x["loss"] - y["loss"]) / y["loss"] for x, y in zip(hs, hr))
                dterm = max(abs(x[k] - y[k]) / max(abs(y[k]), 1e-30) for x, y in zip(hs, hr) for k in ("smooth", "baseline", "bone_length", "pose_temporal"))
                dx = float((Xs - Xr[a:b]).abs().max())
                dtt = float((ts - tr[a:b]).abs().max())
                dR = float((Rs - Rr[a:b]).abs().max())
                print(f"{mode:10s} world={world} graph={graph}: loss dev {dl:.2e}, term dev {dterm:.2e}, max |dX| {dx:.2e} |dt| {dtt:.2e} |dR| {dR:.2e}; "
                      f"loss {hr[0]['loss']:.5f} -> {hr[-1]['loss']:.5f}", flush=True)
                ok = ok and dl < 1e-10 and dterm < 1e-8 and dx < 1e-9 and dtt < 1e-9 and dR < 1e-9
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_FIRST_ORDER_OK" if flag.item() == 1.0 else "MULTI_GPU_FIRST_ORDER_FAILED")
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
