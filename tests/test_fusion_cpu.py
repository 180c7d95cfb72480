"""CPU-side checks of the fusion host logic (row N3): EMA chunk halo, per-joint alpha, dict <-> array conversion and the
C-ABI argument validation (which runs before any CUDA call)."""
import ctypes as C

import numpy as np
import pytest

from oracle import fusion as F
from skiing_analysis_pytorch_b200 import _cabi, _lib, fusion


def test_alpha_per_joint_is_the_references():
    ids = list(range(70))
    for kw in (dict(alpha=0.7, adaptive=True, alpha_min=0.45, alpha_max=0.92), dict(alpha=0.9, adaptive=True, alpha_min=0.5, alpha_max=0.95),
               dict(alpha=0.7, adaptive=False, alpha_min=0.45, alpha_max=0.92)):
        np.testing.assert_array_equal(fusion.alpha_per_joint(ids, **kw), F.alpha_per_joint(ids, **kw))
    assert (fusion.IDX_PELVIS, fusion.IDX_LHIP, fusion.IDX_RHIP, fusion.IDX_LSHO, fusion.IDX_RSHO) == (14, 11, 12, 5, 6)


def test_ema_halo_bounds_the_truncation_error():
    h = fusion.ema_halo(0.7, True, 0.45, 0.92)
    assert 0.55 ** h < 1e-18 <= 0.55 ** (h - 1)
    assert fusion.ema_halo(0.7, False, 0.45, 0.92) == int(np.ceil(np.log(1e-18) / np.log(0.3)))
    assert fusion.ema_halo(0.05, False, 0.0, 1.0) is None      # too slow a decay: sequential scan
    assert fusion.ema_halo(0.7, True, 0.05, 0.9) is None
    # numerical check of the contraction claim on the oracle: two different initial states converge at rate <= rho
    rng = np.random.default_rng(0)
    X = rng.normal(size=(80, 4, 3))
    Xa, Xb = X.copy(), X.copy()
    Xb[0] += 5.0
    Ya, Yb = F.temporal_smooth_ema(Xa), F.temporal_smooth_ema(Xb)
    gap = np.abs(Ya - Yb).max(axis=(1, 2))
    assert gap[h if h < 80 else 79] < 1e-15 and np.all(gap[1:40] <= 0.5500001 * gap[:39] + 1e-300)


def test_dict_array_round_trip():
    seq = [{0: [1.0, 2.0, 3.0], 2: [4.0, 5.0, 6.0]}, {1: [7.0, 8.0, np.nan]}]
    A = fusion.dicts_to_array(seq, [0, 1, 2], 3)
    assert A.shape == (2, 3, 3) and np.isnan(A[0, 1]).all() and np.isnan(A[1, 0]).all()
    back = fusion.array_to_dicts(A, [0, 1, 2])
    assert set(back[0]) == {0, 2} and back[1] == {}  # a row with a NaN is dropped (fuse/fuse.py:76-82)


def test_c_abi_argument_errors_without_gpu():
    lib = _lib.load()
    prm = _cabi.SkaFuseParams(12.0, 0.08, 0, 8, 14, 11, 12, 5, 6, 0)
    fake = C.c_void_p(4096)
    call = lambda J=70, T=1, p=prm, x=fake: lib.ska_fuse_frames_f64(x, fake, fake, fake, T, J, C.byref(p) if p else None, fake, None, None, None, None, None, 0, None)
    assert call(J=97) == -1 and call(J=0) == -1 and call(T=-1) == -1 and call(x=None) == -1 and call(p=None) == -1
    assert call(J=12) == -1  # key joint 14 outside a 12-joint skeleton
    bad = _cabi.SkaFuseParams(12.0, 0.08, 2, 8, 14, 11, 12, 5, 6, 0)
    assert call(p=bad) == -1
    assert call() == -4   # valid arguments but no workspace: SKA_EWORKSPACE before any CUDA call
    assert lib.ska_fuse_workspace_bytes(1000) == 1000 * 56 * 8 and lib.ska_fuse_workspace_bytes(0) == 0
    ema = lambda **k: lib.ska_ema_f64(k.get("x", fake), k.get("T", 4), k.get("J", 3), fake, 1, 0.7, k.get("amin", 0.45), 0.92, 0.25, 512, 70,
                                      k.get("y", C.c_void_p(8192)), None)
    assert ema(x=None) == -1 and ema(J=0) == -1 and ema(amin=0.95) == -1 and ema(y=fake) == -1
    with pytest.raises(RuntimeError):
        import torch
        fusion.temporal_smooth_ema(torch.zeros(3, 4, 3, dtype=torch.float64))
