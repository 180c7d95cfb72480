"""Host side of the host-buffer pipeline: where a rank's threads and page-locked buffers live, and what the link gives.

One process per GPU: before it allocates pinned memory a rank should sit on the CPU socket its GPU hangs off, or every copy
crosses the socket interconnect and all ranks contend for one memory controller.  `bind_to_device_numa` reads the GPU's
PCI address, the kernel's NUMA node for it, and binds the calling process (CPU affinity + preferred memory node) to that
node.  On a single-node host (or a VM that hides the topology: numa_node = -1) it changes nothing and says so.

`measure_link` times pinned copies alone and in both directions at once - the ceiling of api.triangulate_reproject_host.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import torch


def _parse_cpulist(s: str) -> set:
    cpus = set()
    for part in s.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def device_numa_node(device) -> int:
    """NUMA node of a CUDA device from sysfs (-1: unknown / not exposed)."""
    idx = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    p = torch.cuda.get_device_properties(idx)
    try:
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except AttributeError:
        return -1
    f = Path("/sys/bus/pci/devices") / bdf / "numa_node"
    try:
        return int(f.read_text().strip())
    except (OSError, ValueError):
        return -1


def bind_to_device_numa(device) -> dict:
    """Bind this process to the NUMA node of `device`.  Call BEFORE allocating pinned buffers.  Returns what was done."""
    node = device_numa_node(device)
    nodes = sorted(int(d.name[4:]) for d in Path("/sys/devices/system/node").glob("node[0-9]*")) if Path("/sys/devices/system/node").exists() else []
    info = {"gpu_numa_node": node, "host_numa_nodes": len(nodes), "bound": False, "cpus": len(os.sched_getaffinity(0))}
    if node < 0 or len(nodes) < 2:
        info["why"] = "single NUMA node visible" if len(nodes) < 2 else "the GPU's NUMA node is not exposed"
        return info
    try:
        cpus = _parse_cpulist((Path("/sys/devices/system/node") / f"node{node}" / "cpulist").read_text())
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            info["why"] = "no allowed CPU on the GPU's node"
            return info
        os.sched_setaffinity(0, cpus)
        info["cpus"] = len(cpus)
        # set_mempolicy(MPOL_PREFERRED, {node}): page-locked allocations made from now on come from the GPU's node
        libc = ctypes.CDLL(None, use_errno=True)
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        rc = libc.syscall(238, 1, mask, 16 * 64 + 1) if os.uname().machine == "x86_64" else -1
        info["mempolicy"] = "preferred" if rc == 0 else "default (first touch on the bound CPUs)"
        info["bound"] = True
    except OSError as e:
        info["why"] = repr(e)
    return info


def measure_link(device, mbytes: int = 256, repeats: int = 6, barrier=None) -> dict:
    """Pinned host <-> device copy rates of this rank in GB/s: H2D alone, D2H alone, both at once (per direction).
    `barrier` (callable) lines the ranks up so that every rank copies at the same time."""
    dev = torch.device(device)
    n = mbytes << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(h2d, d2h, k):
        torch.cuda.synchronize(dev)
        if barrier is not None:
            barrier()
        e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        e0.record(s1)
        e2.record(s2)
        for _ in range(k):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        e1.record(s1)
        e3.record(s2)
        torch.cuda.synchronize(dev)
        ms = max(e0.elapsed_time(e1) if h2d else 0.0, e2.elapsed_time(e3) if d2h else 0.0)
        return k * n / (ms * 1e-3) / 1e9

    run(True, True, 2)
    return {"h2d_alone": run(True, False, repeats), "d2h_alone": run(False, True, repeats), "both_each": run(True, True, repeats),
            "unit": "GB/s", "mbytes": mbytes}
