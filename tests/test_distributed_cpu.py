"""world_size-2 (and 3) gloo runs of the product's LM sequencer on the CPU-only box: each rank owns a
contiguous frame range, the packed reduced camera system and the trial scalars are all-reduced, and
every rank must reproduce the single-process fp64 oracle trajectory (oracle/lm.py) exactly to
rounding.  The compute engine here is the numpy oracle (tests/oracle_engine.py); the sequencing,
sharding and collectives are the product's (skiing_analysis_pytorch_b200/ba.py)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, rig, T, J, mode, iters, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from oracle import lm
    from skiing_analysis_pytorch_b200.ba import frame_shard
    from tests.oracle_engine import OracleBundleAdjuster

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        clip, R0, t0, X0 = lm.make_problem(rig, T, J)
        a, b = frame_shard(T, world, rank)
        ba = OracleBundleAdjuster(clip.x_fm[a:b], clip.conf_fm[a:b], clip.K, R0, t0, X0[a:b], mode=mode, max_iters=iters)
        ba.run(iters)
        np.savez(Path(out_dir) / f"rank{rank}.npz", R=ba.R, t=ba.t, X=ba.X, a=a, b=b,
                 hist=np.array([[h["cost"], h["trial_cost"], h["lam"], h["rho"], float(h["accepted"]), h["pred"]] for h in ba.history]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,rig,T,J,mode", [(2, "2b", 41, 17, "full"), (3, "4", 20, 17, "pose_cam_t"), (2, "8", 9, 70, "full")])
def test_sharded_lm_over_gloo_matches_single_process_oracle(tmp_path, world, rig, T, J, mode):
    from oracle import lm

    iters = 6
    mp.spawn(_worker, args=(world, _free_port(), rig, T, J, mode, iters, str(tmp_path)), nprocs=world, join=True)
    clip, R0, t0, X0 = lm.make_problem(rig, T, J)
    R, t, X, hist = lm.run_lm(X0, R0, t0, clip.K, clip.x_fm, clip.conf_fm, num_iters=iters, mode=mode)
    ref = np.array([[h["cost"], h["trial_cost"], h["lam"], h["rho"], float(h["accepted"]), h["pred"]] for h in hist])
    outs = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for o in outs:
        # LM cost trajectory: north-star tolerance 1e-4 relative per iteration; the fp64 engines agree far tighter
        np.testing.assert_allclose(o["hist"][:, :2], ref[:, :2], rtol=1e-9)
        # identical accept / reject decisions wherever the decision is not rounding noise (at the
        # optimum F_trial - F is ~1e-13 F and its sign is arbitrary); compare up to the first such trial
        decisive = np.abs(ref[:, 0] - ref[:, 1]) > 1e-9 * ref[:, 0]
        n_ok = len(decisive) if decisive.all() else int(np.argmin(decisive))
        assert n_ok >= 3
        np.testing.assert_array_equal(o["hist"][:n_ok, 4], ref[:n_ok, 4])
        np.testing.assert_allclose(o["hist"][:n_ok, 2], ref[:n_ok, 2], rtol=1e-6)  # damping trajectory
        if n_ok < len(decisive):
            continue
        np.testing.assert_allclose(o["R"], R, atol=1e-9)
        np.testing.assert_allclose(o["t"], t, atol=1e-8)
        np.testing.assert_allclose(o["X"], X.reshape(-1, 3)[int(o["a"]) * J: int(o["b"]) * J], atol=1e-7)
    # every rank took the same decisions and holds the same cameras
    for o in outs[1:]:
        np.testing.assert_array_equal(o["hist"], outs[0]["hist"])
        np.testing.assert_array_equal(o["R"], outs[0]["R"])
    assert sum(int(o["b"]) - int(o["a"]) for o in outs) == T


def _calib_worker(rank, world, port, T, J, iters, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from oracle import lm_calib as lc
    from skiing_analysis_pytorch_b200.ba import frame_shard
    from tests.oracle_engine import OracleCalibratingBundleAdjuster

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        clip, R0, t0, th, X0 = lc.make_problem("2b", T, J)
        a, b = frame_shard(T, world, rank)
        ba = OracleCalibratingBundleAdjuster(clip.x_fm[a:b], clip.conf_fm[a:b], th, R0, t0, X0[a:b], max_iters=iters,
                                             prior_theta=lc.intr_from_K(clip.K), prior_rho=lc.PRIOR_RHO)
        ba.run(iters)
        np.savez(Path(out_dir) / f"rank{rank}.npz", th=ba.th, R=ba.R,
                 hist=np.array([[h["cost"], h["trial_cost"], h["lam"], float(h["accepted"])] for h in ba.history]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_calibrating_lm_over_gloo_matches_single_process_oracle(tmp_path, world):
    """The calibrating BA's payload (n = 15 C - 6 reduced system + two 17 x 17 camera blocks, include/ska.h) is a plain
    sum over frames: frame-sharded ranks all-reducing it reproduce oracle/lm_calib.py's single-process trajectory."""
    from oracle import lm_calib as lc

    T, J, iters = 31, 17, 6
    mp.spawn(_calib_worker, args=(world, _free_port(), T, J, iters, str(tmp_path)), nprocs=world, join=True)
    clip, R0, t0, th, X0 = lc.make_problem("2b", T, J)
    R, t, th1, X, hist = lc.run_lm(X0, R0, t0, th, clip.x_fm, clip.conf_fm, num_iters=iters, prior_theta=lc.intr_from_K(clip.K),
                                   prior_rho=lc.PRIOR_RHO)
    ref = np.array([[h["cost"], h["trial_cost"], h["lam"], float(h["accepted"])] for h in hist])
    outs = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for o in outs:
        np.testing.assert_allclose(o["hist"][:, :2], ref[:, :2], rtol=1e-8)
        decisive = np.abs(ref[:, 0] - ref[:, 1]) > 1e-8 * ref[:, 0]
        n_ok = len(decisive) if decisive.all() else int(np.argmin(decisive))
        assert n_ok >= 3
        np.testing.assert_array_equal(o["hist"][:n_ok, 3], ref[:n_ok, 3])
        if n_ok == len(decisive):
            np.testing.assert_allclose(o["th"], th1, rtol=1e-6, atol=1e-6)
    for o in outs[1:]:
        np.testing.assert_array_equal(o["hist"], outs[0]["hist"])
        np.testing.assert_array_equal(o["th"], outs[0]["th"])


def test_frame_shard_partitions_the_clip():
    from skiing_analysis_pytorch_b200.ba import frame_shard

    for T in (0, 1, 7, 100, 1_000_003):
        for world in (1, 2, 3, 8):
            r = [frame_shard(T, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == T
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    with pytest.raises(ValueError):
        frame_shard(10, 2, 2)


def test_single_process_oracle_engine_equals_run_lm():
    """The engine adapter itself (payload pack / unpack) is faithful: world = 1, no process group."""
    from oracle import lm
    from tests.oracle_engine import OracleBundleAdjuster

    clip, R0, t0, X0 = lm.make_problem("3", 12, 17)
    ba = OracleBundleAdjuster(clip.x_fm, clip.conf_fm, clip.K, R0, t0, X0, max_iters=5).run(5)
    _, _, _, hist = lm.run_lm(X0, R0, t0, clip.K, clip.x_fm, clip.conf_fm, num_iters=5)
    np.testing.assert_allclose([h["cost"] for h in ba.history], [h["cost"] for h in hist], rtol=1e-10)
    np.testing.assert_allclose([h["trial_cost"] for h in ba.history], [h["trial_cost"] for h in hist], rtol=1e-10)
    assert [h["accepted"] for h in ba.history] == [h["accepted"] for h in hist]
    assert torch.is_tensor(ba.red) and ba.red.dtype == torch.float64


def _reg_worker(rank, world, port, rig, T, J, mode, iters, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from oracle import lm_reg
    from skiing_analysis_pytorch_b200.ba import frame_shard
    from tests.oracle_engine import OracleRegularisedBundleAdjuster

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        clip, R, t, X0 = lm_reg.make_problem(rig, T, J, cam_jitter=0.01)
        a, b = frame_shard(T, world, rank)
        s = OracleRegularisedBundleAdjuster(clip.x_fm[a:b], clip.conf_fm[a:b], clip.K, R[a:b], t[a:b], X0[a:b], (a, b), (clip.x_fm, clip.conf_fm),
                                            mode=mode, max_iters=iters)
        s.run(iters)
        np.savez(Path(out_dir) / f"rank{rank}.npz", X=s.X, a=a, b=b,
                 hist=np.array([[h["cost"], h["trial_cost"], h["lam"], float(h["accepted"]), h["cg_iters"]] for h in s.history]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,rig,T,J,mode", [(2, "2b", 21, 17, "pose_only"), (3, "2b", 14, 17, "full"), (2, "4", 9, 17, "pose_cam_t")])
def test_sharded_regularised_lm_over_gloo_matches_single_process_oracle(tmp_path, world, rig, T, J, mode):
    """Row e3: the regularised LM sharded by frame range - the product's RegLMSequencer moving the CG dot products, the
    cost sums (bone / baseline means are global) and the one-frame halos of the CG direction and of the trial point over
    gloo - reproduces the single-process exact-solve oracle (oracle/lm_reg.py)."""
    from oracle import lm_reg

    iters = 5
    mp.spawn(_reg_worker, args=(world, _free_port(), rig, T, J, mode, iters, str(tmp_path)), nprocs=world, join=True)
    clip, R, t, X0 = lm_reg.make_problem(rig, T, J, cam_jitter=0.01)
    _, _, X, hist = lm_reg.run_lm(X0, R, t, clip.K, clip.x_fm.astype(float), clip.conf_fm.astype(float), num_iters=iters, mode=mode)
    ref = np.array([[h["cost"], h["trial_cost"], h["lam"], float(h["accepted"])] for h in hist])
    outs = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for o in outs:
        np.testing.assert_allclose(o["hist"][:, :2], ref[:, :2], rtol=1e-7)
        decisive = np.abs(ref[:, 0] - ref[:, 1]) > 1e-9 * ref[:, 0]
        n_ok = len(decisive) if decisive.all() else int(np.argmin(decisive))
        assert n_ok >= 2  # pose_only converges quadratically: two decisive trials
        np.testing.assert_array_equal(o["hist"][:n_ok, 3], ref[:n_ok, 3])
        if n_ok == len(decisive):
            np.testing.assert_allclose(o["X"], X[int(o["a"]): int(o["b"])], atol=1e-6)
    for o in outs[1:]:
        np.testing.assert_array_equal(o["hist"], outs[0]["hist"])
    assert sum(int(o["b"]) - int(o["a"]) for o in outs) == T


def test_regularised_sequencer_needs_its_halo(tmp_path):
    """World 1 through the same engine equals the oracle; the single-process engine adapter itself is faithful."""
    from oracle import lm_reg
    from tests.oracle_engine import OracleRegularisedBundleAdjuster

    clip, R, t, X0 = lm_reg.make_problem("2b", 10, 17, cam_jitter=0.01)
    s = OracleRegularisedBundleAdjuster(clip.x_fm, clip.conf_fm, clip.K, R, t, X0, (0, 10), (clip.x_fm, clip.conf_fm), mode="full", max_iters=4)
    s.run(4)
    _, _, _, hist = lm_reg.run_lm(X0, R, t, clip.K, clip.x_fm.astype(float), clip.conf_fm.astype(float), num_iters=4, mode="full")
    np.testing.assert_allclose([h["trial_cost"] for h in s.history], [h["trial_cost"] for h in hist], rtol=1e-7)
    assert [h["accepted"] for h in s.history] == [h["accepted"] for h in hist]


def test_peer_exchange_is_optional_plumbing():
    """No process group (or a single rank): peer.shared returns None and the sequencers keep their collectives - the peer
    kernels are an accelerator of the exchange step, not a requirement (tests/test_peer_gpu.py exercises them on a GPU)."""
    from skiing_analysis_pytorch_b200 import peer
    from skiing_analysis_pytorch_b200.ba import LMSequencer
    from skiing_analysis_pytorch_b200.ba_reg import RegLMSequencer

    assert peer.PeerExchange.create(None, "cuda:0") is None
    assert LMSequencer.peer is None and RegLMSequencer.peer is None
