"""The peer-memory exchange kernels (csrc/ska_peer.cu) through the C ABI on ONE GPU: two "ranks" are two regions of the same
device and two streams - the kernels of both ranks are co-resident, push into each other's receive areas and wait for each
other's flags exactly as two processes on two GPUs do (tools/peer_bench.py runs that, on 2 and 8 B200s)."""
import ctypes as C

import pytest
import torch

from skiing_analysis_pytorch_b200 import _cabi, _lib

pytestmark = pytest.mark.gpu


def _comms(world, slot, dev, poll_limit_log2=0):
    lib = _lib.load()
    n = int(lib.ska_peer_region_bytes(world, slot))
    regions = [torch.zeros(n // 8, dtype=torch.float64, device=dev) for _ in range(world)]
    states = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    ptrs = [r.data_ptr() for r in regions]
    comms = [_cabi.SkaPeerComm(world=world, rank=r, slot_doubles=slot, poll_limit_log2=poll_limit_log2, recv=(C.c_void_p * 8)(*ptrs),
                               d_state=states[r].data_ptr()) for r in range(world)]
    return lib, comms, regions, states


@pytest.mark.parametrize("world", [2, 4])
def test_allreduce_and_allgather_between_co_resident_ranks(cuda, world):
    lib, comms, regions, states = _comms(world, 256, cuda)
    streams = [torch.cuda.Stream(cuda) for _ in range(world)]
    g = torch.Generator(device=cuda).manual_seed(0)
    for it, n in enumerate((1, 4, 40, 158, 256, 7, 256, 3)):  # consecutive exchanges: the receive areas alternate by parity
        xs = [torch.randn(n, dtype=torch.float64, device=cuda, generator=g) for _ in range(world)]
        exact = torch.zeros(n, dtype=torch.float64, device=cuda)
        for x in xs:
            exact = exact + x  # the kernel's order: rank 0 + rank 1 + ...
        torch.cuda.synchronize()
        gather = it % 2 == 1
        outs = [torch.empty((world, n), dtype=torch.float64, device=cuda) for _ in range(world)]
        bufs = [x.clone() for x in xs]
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                sp = C.c_void_p(streams[r].cuda_stream)
                if gather:
                    _lib.check(lib.ska_peer_allgather_f64(C.byref(comms[r]), C.c_void_p(xs[r].data_ptr()), n, C.c_void_p(outs[r].data_ptr()), sp))
                else:
                    _lib.check(lib.ska_peer_allreduce_f64(C.byref(comms[r]), C.c_void_p(bufs[r].data_ptr()), n, sp))
        torch.cuda.synchronize()
        for r in range(world):
            if gather:
                assert torch.equal(outs[r], torch.stack(xs))
            else:
                assert torch.equal(bufs[r], exact)  # bit-identical on every rank
            assert states[r].tolist() == [it + 1, 0]


def test_a_missing_peer_times_out_instead_of_hanging(cuda):
    lib, comms, regions, states = _comms(2, 16, cuda, poll_limit_log2=12)
    x = torch.ones(4, dtype=torch.float64, device=cuda)
    _lib.check(lib.ska_peer_allreduce_f64(C.byref(comms[0]), C.c_void_p(x.data_ptr()), 4, None))  # rank 1 never shows up
    torch.cuda.synchronize()
    assert states[0].tolist() == [1, 1]  # exchange 1 gave up


def test_argument_checks(cuda):
    lib, comms, regions, states = _comms(2, 16, cuda)
    x = torch.ones(32, dtype=torch.float64, device=cuda)
    assert lib.ska_peer_allreduce_f64(C.byref(comms[0]), C.c_void_p(x.data_ptr()), 17, None) == -1  # larger than the slot
    assert lib.ska_peer_allreduce_f64(None, C.c_void_p(x.data_ptr()), 4, None) == -1
    bad = _cabi.SkaPeerComm(world=9, rank=0, slot_doubles=16)
    assert lib.ska_peer_allreduce_f64(C.byref(bad), C.c_void_p(x.data_ptr()), 4, None) == -1
    assert lib.ska_peer_region_bytes(0, 16) == 0 and lib.ska_peer_region_bytes(2, 16) == 2 * 2 * 16 * 16
