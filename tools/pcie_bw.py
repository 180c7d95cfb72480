"""Host <-> device copy bandwidth per rank and in aggregate (pinned buffers, both directions at once), to separate what the
end-to-end path can reach from what the box's host side gives: run under torchrun with 1, 2, 4, 8 ranks.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_bw.py"""
import mmap
import os
import sys
import time

import torch
import torch.distributed as dist


_KEEP = []


def huge_pinned(nbytes):
    """Page-locked host buffer backed by 2 MiB transparent huge pages (mmap + MADV_HUGEPAGE + cudaHostRegister): 512x fewer
    IOMMU / page-table entries per byte than cudaHostAlloc's 4 KiB pages."""
    hp = 2 << 20
    n = (nbytes + hp - 1) & ~(hp - 1)
    m = mmap.mmap(-1, n + hp, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    whole = torch.frombuffer(m, dtype=torch.uint8)
    off = (-whole.data_ptr()) % hp
    m.madvise(mmap.MADV_HUGEPAGE, 0, n + hp) if off == 0 else m.madvise(mmap.MADV_HUGEPAGE)
    t = whole[off:off + n]
    t.fill_(1)  # touch: the pages are allocated here
    rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), n, 0)
    if int(rc) != 0:
        raise RuntimeError(f"cudaHostRegister failed: {rc}")
    _KEEP.append((m, whole))
    return t[:nbytes]


def anon_huge_kb():
    for line in open("/proc/self/smaps_rollup"):
        if line.startswith("AnonHugePages"):
            return int(line.split()[1])
    return -1


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mb = 256
    huge = "--huge" in sys.argv
    if huge:
        h_in, h_out = huge_pinned(mb << 20), huge_pinned(mb << 20)
    else:
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
        print(f"ranks {world} ({'2 MiB huge pages, AnonHugePages %d kB' % anon_huge_kb() if huge else 'cudaHostAlloc'}): per rank H2D alone {a:.1f} GiB/s, D2H alone {b:.1f} GiB/s, both at once {c:.1f} + {c:.1f} GiB/s; "
              f"aggregate both at once {2 * c * world:.0f} GiB/s  (cpus {len(os.sched_getaffinity(0))})", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
