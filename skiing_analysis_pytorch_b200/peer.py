"""The sharded solvers' exchange step over NVLink peer memory (csrc/ska_peer.cu): all-reduce / all-gather of small fp64
payloads by one single-CTA kernel per exchange instead of an NCCL collective (~5 us against ~20-30 us on 8 GPUs; the
payloads are 1 .. 1 179 doubles, so the collectives' latency is all there is).  One process per GPU, one node.

Set-up uses torch.distributed once (the 64-byte CUDA IPC handles of every rank's receive region travel through
all_gather_object); afterwards no library collective runs on the data path.  Every rank must issue the same sequence of
exchanges.  `PeerExchange.create` returns None where peer memory cannot be set up (no process group, a single rank, IPC
refused by the platform): the callers then keep torch.distributed's NCCL collectives - both are GPU paths."""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi, _lib


_SHARED: dict = {}


def shared(group, device, slot_doubles: int = 2048):
    """One exchange per (process group, device), created on first use (collective: every rank must call it) - or None."""
    key = (id(group) if group is not None else 0, torch.device(device).index)
    if key not in _SHARED:
        _SHARED[key] = PeerExchange.create(group, device, slot_doubles)
    return _SHARED[key]


def close_all():
    for p in _SHARED.values():
        if p is not None:
            p.close()
    _SHARED.clear()


class PeerExchange:
    def __init__(self, group, device, slot_doubles: int = 2048):
        d = torch.distributed
        self.group, self.dev = group, torch.device(device)
        self.world, self.rank = d.get_world_size(group), d.get_rank(group)
        if not 2 <= self.world <= _cabi.MAX_PEERS:
            raise ValueError(f"peer exchange needs 2..{_cabi.MAX_PEERS} ranks")
        self.slot = int(slot_doubles)
        self.lib = lib = _lib.load()
        self._region = C.c_void_p()
        self._imported = []
        with torch.cuda.device(self.dev):
            nbytes = int(lib.ska_peer_region_bytes(self.world, self.slot))
            _lib.check(lib.ska_peer_alloc(nbytes, C.byref(self._region)))
            handle = C.create_string_buffer(64)
            _lib.check(lib.ska_peer_export(self._region, handle))
            everyone = [None] * self.world
            d.all_gather_object(everyone, (self.rank, bytes(handle.raw), torch.cuda.current_device()), group=group)
            ptrs = [None] * self.world
            for r, raw, _ in everyone:
                if r == self.rank:
                    ptrs[r] = self._region.value
                else:
                    p = C.c_void_p()
                    _lib.check(lib.ska_peer_import(raw, C.byref(p)))
                    self._imported.append(p)
                    ptrs[r] = p.value
        self.state = torch.zeros(2, dtype=torch.int64, device=self.dev)
        self.comm = _cabi.SkaPeerComm(world=self.world, rank=self.rank, slot_doubles=self.slot, poll_limit_log2=0,
                                      recv=(C.c_void_p * 8)(*ptrs), flags=(C.c_void_p * 8)(),
                                      d_state=self.state.data_ptr(), d_skip=None)
        self._skip_comms = {}
        d.barrier(group=group)  # every rank has mapped every region before the first push

    @classmethod
    def create(cls, group, device, slot_doubles: int = 2048):
        d = torch.distributed
        if not (d.is_available() and d.is_initialized()) or d.get_world_size(group) < 2 or d.get_world_size(group) > _cabi.MAX_PEERS:
            return None
        ok, obj = 1, None
        try:
            obj = cls(group, device, slot_doubles)
        except Exception:  # IPC refused (container policy, devices without peer access): keep NCCL
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=torch.device(device))
        d.all_reduce(flag, op=d.ReduceOp.MIN, group=group)  # all ranks or none
        if int(flag.item()) == 0:
            if obj is not None:
                obj.close()
            return None
        return obj

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def fits(self, t: torch.Tensor) -> bool:
        return t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.numel() <= self.slot

    def _comm(self, skip):
        """skip: None, or a one-element fp64 device tensor that is identical on every rank; non-zero = skip the exchange."""
        if skip is None:
            return self.comm
        key = skip.data_ptr()
        if key not in self._skip_comms:
            c = _cabi.SkaPeerComm.from_buffer_copy(self.comm)
            c.d_skip = key
            self._skip_comms[key] = c
        return self._skip_comms[key]

    def all_reduce(self, t: torch.Tensor, skip=None):
        """In-place sum over the ranks (fixed rank order: bit-identical everywhere)."""
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.ska_peer_allreduce_f64(C.byref(self._comm(skip)), C.c_void_p(t.data_ptr()), t.numel(), self._stream()))

    def all_gather(self, out: torch.Tensor, inp: torch.Tensor, skip=None):
        """out (world, n) <- every rank's inp (n,)."""
        if out.numel() != self.world * inp.numel():
            raise ValueError("out must hold world x inp.numel() elements")
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.ska_peer_allgather_f64(C.byref(self._comm(skip)), C.c_void_p(inp.data_ptr()), inp.numel(),
                                                       C.c_void_p(out.data_ptr()), self._stream()))

    def check(self):
        """Synchronises; raises if an exchange gave up waiting for a peer."""
        st = self.state.cpu()
        if int(st[1]) != 0:
            raise RuntimeError(f"peer exchange {int(st[1])} timed out waiting for a peer (rank {self.rank})")
        return int(st[0])

    def close(self):
        """Collective: unmap the peers' regions everywhere, THEN free the own one (a region is freed only after every
        importer has closed it)."""
        d = torch.distributed
        with torch.cuda.device(self.dev):
            torch.cuda.synchronize(self.dev)
            live = d.is_available() and d.is_initialized()
            if live:
                d.barrier(group=self.group)
            for p in self._imported:
                self.lib.ska_peer_close(p)
            self._imported = []
            if live:
                d.barrier(group=self.group)
            if self._region.value:
                self.lib.ska_peer_free(self._region)
                self._region = C.c_void_p()
