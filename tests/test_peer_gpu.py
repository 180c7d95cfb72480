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


def test_skip_flag_leaves_the_counters_in_step(cuda):
    """d_skip: a device flag every rank holds identically (the CG convergence flag); non-zero = every rank skips the exchange."""
    lib, comms, regions, states = _comms(2, 16, cuda)
    skip = torch.ones(1, dtype=torch.float64, device=cuda)
    x = torch.full((4,), 3.0, dtype=torch.float64, device=cuda)
    c0 = _cabi.SkaPeerComm.from_buffer_copy(comms[0])
    c0.d_skip = skip.data_ptr()
    _lib.check(lib.ska_peer_allreduce_f64(C.byref(c0), C.c_void_p(x.data_ptr()), 4, None))  # alone: would time out if it did not skip
    torch.cuda.synchronize()
    assert states[0].tolist() == [0, 0] and torch.equal(x, torch.full((4,), 3.0, dtype=torch.float64, device=cuda))


def test_two_co_resident_ranks_run_the_sharded_schur_lm(cuda):
    """The frame-sharded Schur LM with the all-reduces FUSED into the solve / control kernels (SkaBaProblem.peer), on one GPU:
    two solvers hold the two halves of a clip, their trials are enqueued on two streams, and each solve / control kernel
    pushes its payload into the other's receive area and waits for the other's - exactly what two processes on two GPUs do
    (tools/ba_multi_gpu_check.py, bench.py's `parity` objects).  Both must follow the unsharded solve."""
    from oracle import lm
    from skiing_analysis_pytorch_b200 import ba

    clip, R0, t0, X0 = lm.make_problem("3", 64, 17)
    x = torch.from_numpy(clip.x_fm).to(cuda)
    c = torch.from_numpy(clip.conf_fm).to(cuda)
    X = torch.from_numpy(X0.astype("float32")).to(cuda)
    whole = ba.BundleAdjuster(x, c, clip.K, R0, t0, X, max_iters=8)
    whole.run(6)
    lib, comms, regions, states = _comms(2, 2048, cuda)
    halves, streams = [], [torch.cuda.Stream(cuda) for _ in range(2)]
    for r, (a, b) in enumerate(((0, 29), (29, 64))):
        s = ba.BundleAdjuster(x[a:b].contiguous(), c[a:b].contiguous(), clip.K, R0, t0, X[a:b].contiguous(), max_iters=8)
        s.ctrl[_cabi.BA_CTRL_SUMCONF] = whole.ctrl[_cabi.BA_CTRL_SUMCONF]  # the global sum of confidences (an all-reduce at set-up)
        s.prob.peer = C.addressof(comms[r])
        s.fused_exchange = True
        halves.append(s)
    torch.cuda.synchronize()
    for _ in range(6):
        for r in range(2):
            with torch.cuda.stream(streams[r]):
                halves[r].trial()
                halves[r].iters_done += 1
    torch.cuda.synchronize()
    assert [st.tolist()[1] for st in states] == [0, 0]  # no exchange timed out
    for s in halves:
        n_ok = 0
        for h, w in zip(s.history, whole.history):
            assert abs(h["cost"] - w["cost"]) <= 1e-5 * w["cost"] and abs(h["trial_cost"] - w["trial_cost"]) <= 1e-5 * w["trial_cost"]
            if abs(w["cost"] - w["trial_cost"]) < 1e-3 * w["cost"]:
                break  # converged: the decision is rounding noise of two fp32 summation orders, the damping sequences part ways
            assert h["accepted"] == w["accepted"]
            n_ok += 1
        assert n_ok >= 2
    assert halves[0].history == halves[1].history  # bit-identical sums on both ranks: identical decisions and cameras
    assert torch.equal(halves[0].cams, halves[1].cams)
