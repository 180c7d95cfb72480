"""The exchange step over NVLink peer memory against NCCL (run under torchrun on N GPUs): correctness of the all-reduce /
all-gather, latency per exchange (eager and inside a CUDA graph), and what it does to the sharded LM trials.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/peer_bench.py
"""
import math
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from skiing_analysis_pytorch_b200 import api, ba, ba_reg, peer, synth  # noqa: E402


def timed(fn, n, dev):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) * 1e3  # us


def graphed(fn, reps, dev):
    fn()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
    torch.cuda.current_stream(dev).wait_stream(s)
    return g


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    px = peer.shared(None, dev)
    say = (lambda *a: print(*a, flush=True)) if rank == 0 else (lambda *a: None)
    if px is None:
        say("peer memory could not be set up on this box (IPC refused): NCCL stays")
        dist.destroy_process_group()
        return
    ok = True
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    for n in (1, 4, 40, 158, 1179, 2048):
        x = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
        a, b = x.clone(), x.clone()
        px.all_reduce(a)
        dist.all_reduce(b)
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x)
        ref = torch.stack(parts).sum(0) if world > 2 else parts[0] + parts[1]  # the kernel's order: rank 0 + rank 1 + ...
        exact = torch.zeros_like(x)
        for p in parts:
            exact = exact + p
        out = torch.empty((world, n), dtype=torch.float64, device=dev)
        px.all_gather(out, x)
        same_everywhere = a.clone()
        dist.broadcast(same_everywhere, 0)
        good = bool(torch.equal(a, exact)) and bool(torch.equal(out, torch.stack(parts))) and bool(torch.equal(a, same_everywhere)) \
            and float((a - b).abs().max()) <= 1e-12 * max(1.0, float(b.abs().max()))
        ok = ok and good
        del ref
    px.check()
    say(f"world {world}: all-reduce equals the rank-ordered sum bit for bit on every rank, all-gather exact: {ok}")
    for n in (4, 158, 1179):
        x = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
        t_peer = timed(lambda: px.all_reduce(x), 200, dev)
        t_nccl = timed(lambda: dist.all_reduce(x), 200, dev)
        gp, gn = graphed(lambda: px.all_reduce(x), 50, dev), graphed(lambda: dist.all_reduce(x), 50, dev)
        t_gp, t_gn = timed(gp.replay, 20, dev) / 50, timed(gn.replay, 20, dev) / 50
        del gp, gn  # a live graph holding captured NCCL kernels keeps destroy_process_group waiting
        say(f"  all-reduce of {n:5d} doubles: peer {t_peer:6.1f} us eager, {t_gp:6.1f} us in a graph | NCCL {t_nccl:6.1f} us eager, {t_gn:6.1f} us in a graph")
    px.check()
    # ---- the sharded LM trials
    iters = 20
    if "--quick" in sys.argv:
        say("PEER_EXCHANGE_OK" if ok else "PEER_EXCHANGE_FAILED")
        torch.cuda.synchronize()
        dist.barrier()
        peer.close_all()
        dist.destroy_process_group()
        return
    for name, rig, T, J in (("config 3", "2b", 100_000, 17), ("config 5", "8", 1_000_000, 70)):
        Tl = T // world
        d = synth.make_clip_device(rig, Tl, J, dev, seed=100, shardable=True, frame_offset=rank * Tl)
        R0, t0 = synth.perturb_cameras(d["R"], d["t"], seed=1)
        X0 = api.triangulate_reproject(d["x2d"].permute(1, 0, 2, 3).contiguous(), d["K"], R0, t0, want=("X",)).X
        res = {}
        for use in (True, "kernel", False):
            s = ba.BundleAdjuster(d["x2d"], d["conf"], d["K"], R0, t0, X0, max_iters=iters + 8, peer_exchange=bool(use), fuse_exchange=use is True)
            s.run(4, graph=True)
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s.run(iters, graph=True)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[use] = (float(t.item()), [h["cost"] for h in s.history], [h["accepted"] for h in s.history])
            del s
        dc = max(abs(x - y) / y for x, y in zip(res[True][1], res[False][1]))
        say(f"  {name} ({T} x {J} over {world} GPUs): {res[True][0]:.4f} ms per trial with the peer exchange fused into solve / control, "
            f"{res['kernel'][0]:.4f} with stand-alone exchange kernels, {res[False][0]:.4f} with NCCL; "
            f"cost trajectories agree to {dc:.1e}, decisions equal {res[True][2] == res[False][2]}")
        ok = ok and dc < 1e-9
        del d, X0
        torch.cuda.empty_cache()
    T, J, rig = 100_000, 17, "2b"
    Tl = T // world
    d = synth.make_clip_device(rig, Tl, J, dev, seed=100, shardable=True, frame_offset=rank * Tl)
    R0, t0 = synth.perturb_cameras(d["R"], d["t"], seed=1)
    C = len(R0)
    X0 = api.triangulate_reproject(d["x2d"].permute(1, 0, 2, 3).contiguous(), d["K"], R0, t0, want=("X",)).X.double()
    tg = torch.arange(rank * Tl, (rank + 1) * Tl, dtype=torch.float64, device=dev)
    drift = torch.stack([0.01 * torch.sin(2 * math.pi * tg / 200.0 + c) for c in range(C)], 1)[..., None] * torch.tensor([1.0, 0.5, 0.25], dtype=torch.float64, device=dev)
    R = torch.tensor(R0, device=dev)[None].expand(Tl, C, 3, 3).contiguous()
    t = (torch.tensor(t0, device=dev)[None] + drift).contiguous()
    for mode, cg in (("pose_only", 6), ("full", 48)):
        res = {}
        for use in (True, False):
            s = ba_reg.RegularisedBundleAdjuster(d["x2d"], d["conf"], d["K"], R, t, X0, mode=mode, max_iters=iters + 8, cg_iters=cg, peer_exchange=use)
            s.run(3, graph=True)
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s.run(10, graph=True)
            e1.record()
            torch.cuda.synchronize()
            tt = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            res[use] = (float(tt.item()), [h["trial_cost"] for h in s.history])
            del s
        dc = max(abs(x - y) / y for x, y in zip(res[True][1], res[False][1]))
        say(f"  regularised LM {mode} (100k x 17 x 2 per-frame cameras over {world} GPUs): {res[True][0]:.3f} ms per trial with the peer exchange, "
            f"{res[False][0]:.3f} with NCCL; trial costs agree to {dc:.1e}")
        ok = ok and dc < 1e-9
    px.check()
    say("PEER_EXCHANGE_OK" if ok else "PEER_EXCHANGE_FAILED")
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.broadcast(flag, 0)
    peer.close_all()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
