"""Host <-> device copy bandwidth per rank and in aggregate (pinned buffers, both directions at once), to separate what the
end-to-end path can reach from what the box's host side gives: run under torchrun with 1, 2, 4, 8 ranks.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_bw.py"""
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mb = 256
    h_in = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    d_out = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(h2d, d2h, n=8):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return n * mb / 1024 / float(t.item())  # GiB/s per direction per rank at the slowest rank's pace

    run(True, True, 2)
    a, b, c = run(True, False), run(False, True), run(True, True)
    if rank == 0:
        print(f"ranks {world}: per rank H2D alone {a:.1f} GiB/s, D2H alone {b:.1f} GiB/s, both at once {c:.1f} + {c:.1f} GiB/s; "
              f"aggregate both at once {2 * c * world:.0f} GiB/s  (cpus {len(os.sched_getaffinity(0))})", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
