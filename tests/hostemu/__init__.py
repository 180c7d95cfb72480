"""Host emulation of the per-point device arithmetic - TEST INFRASTRUCTURE for the CPU-only box.

Compiles skiing_analysis_pytorch_b200/csrc/ska_tri_point.cuh with g++ so the fp32 secular solver,
its fp64 Jacobi fallback and the differential reprojection scoring can be checked against the
fp64 oracle without a GPU.  The package never imports this; the product path is CUDA only.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SO = HERE / "_build" / "hostemu.so"
SRC = HERE / "hostemu.cpp"
CSRC = HERE.parent.parent / "skiing_analysis_pytorch_b200" / "csrc"


def build() -> Path:
    deps = [SRC, *CSRC.glob("*.cuh"), *CSRC.glob("*.h")]
    if SO.exists() and SO.stat().st_mtime >= max(d.stat().st_mtime for d in deps):
        return SO
    SO.parent.mkdir(exist_ok=True)
    subprocess.run(
        ["g++", "-O2", "-mfma", "-ffp-contract=fast", "-shared", "-fPIC", "-x", "c++", str(SRC), "-o", str(SO)], check=True
    )
    return SO


def triangulate(cams, V, kpts_vm, conf_vm=None, flags=0, centre=None):
    """kpts_vm (V,N,2) f32, conf_vm (V,N) f32|None -> X (N,3) f32, err (V,N) f32, status (N,) u8."""
    lib = C.CDLL(str(build()))
    k = np.ascontiguousarray(kpts_vm, np.float32)
    N = k.shape[1]
    X = np.zeros((N, 3), np.float32)
    err = np.zeros((V, N), np.float32)
    st = np.zeros(N, np.uint8)
    cf = None if conf_vm is None else np.ascontiguousarray(conf_vm, np.float32)
    cen = None if centre is None else np.ascontiguousarray(centre, np.float64)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    rc = lib.hostemu_triangulate(cams, C.c_int32(V), p(cen), p(k), p(cf), C.c_int64(N), C.c_uint32(flags), p(X), p(err), p(st))
    if rc != 0:
        raise RuntimeError(f"hostemu rc={rc}")
    return X, err, st


def calib_obs(cam24, X, uv):
    """One camera (24 doubles: R t theta pad), X (N,3) f32, uv (N,2) f32 -> au, av (N,3), bu, bv (N,17), clamped (N,),
    err2 (N,) exactly as the calibrating-BA kernels compute them (ska_ba_calib.cuh)."""
    lib = C.CDLL(str(build()))
    X = np.ascontiguousarray(X, np.float32)
    uv = np.ascontiguousarray(uv, np.float32)
    cam = np.ascontiguousarray(cam24, np.float64)
    N = X.shape[0]
    au, av = np.zeros((N, 3), np.float32), np.zeros((N, 3), np.float32)
    bu, bv = np.zeros((N, 17), np.float32), np.zeros((N, 17), np.float32)
    cl, e2 = np.zeros(N, np.uint8), np.zeros(N, np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.hostemu_calib_obs(p(cam), p(X), p(uv), C.c_int64(N), p(au), p(av), p(bu), p(bv), p(cl), p(e2))
    if rc != 0:
        raise RuntimeError(f"hostemu rc={rc}")
    return au, av, bu, bv, cl, e2


def calib_tri_maps():
    lib = C.CDLL(str(build()))
    return lib.hostemu_calib_tri, lib.hostemu_calib_tri_row, lib.hostemu_calib_tri_col
