"""CPU-side boundary checks: the C-ABI library builds for sm_100a, loads, and exports every symbol
include/ska.h declares.  No compute calls (there is no GPU on the authoring box)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _header_symbols():
    txt = (ROOT / "include" / "ska.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ska_[a-z0-9_]+)\s*\(", txt)))


def _cabi_version():
    """SKA_ABI_VERSION of include/ska.h (the ctypes side carries its own copy and refuses a mismatching library)."""
    from skiing_analysis_pytorch_b200 import _cabi

    hdr = int(re.search(r"#define SKA_ABI_VERSION (\d+)", (ROOT / "include" / "ska.h").read_text()).group(1))
    assert hdr == _cabi.ABI_VERSION
    return hdr


def test_library_builds_loads_and_exports_header_symbols():
    from skiing_analysis_pytorch_b200 import _lib, build

    build.build()
    lib = _lib.load()
    syms = _header_symbols()
    assert "ska_triangulate_reproject_f32" in syms
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ska.h but not exported by libska.so"
    assert sorted(_lib.exported_symbols()) == syms, "ctypes signature table and header disagree"
    assert lib.ska_abi_version() == _cabi_version()
    assert lib.ska_build_arch() == b"sm_100a"


def test_library_contains_sm100a_sass():
    import shutil
    import subprocess

    from skiing_analysis_pytorch_b200 import _lib

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_argument_errors_without_gpu():
    """Validation happens before any CUDA call, so it is testable on the CPU box."""
    from skiing_analysis_pytorch_b200 import _cabi, _lib, synth

    lib = _lib.load()
    R, t = synth.rig("2b")
    cams = _cabi.make_cameras(synth.K_CALIB, R, t)
    fake = C.c_void_p(256)
    args = lambda **kw: [kw.get("cams", cams), kw.get("V", 2), None, kw.get("kpts", fake), None, kw.get("T", 1),
                         kw.get("J", 17), kw.get("layout", 0), kw.get("flags", 0), kw.get("X", fake), None, None, None, None]
    assert lib.ska_triangulate_reproject_f32(*args(V=1)) == -1
    assert lib.ska_triangulate_reproject_f32(*args(V=9)) == -1
    assert lib.ska_triangulate_reproject_f32(*args(kpts=None)) == -1
    assert lib.ska_triangulate_reproject_f32(*args(layout=7)) == -1
    assert lib.ska_triangulate_reproject_f32(*args(flags=3)) == -1
    assert lib.ska_triangulate_reproject_f32(*args(kpts=C.c_void_p(260))) == -3
    assert lib.ska_triangulate_reproject_f32(*args(T=2**31, J=1)) == -1
    assert b"2^31" in lib.ska_last_error()
    assert lib.ska_triangulate_reproject_f32(*args(T=0)) == 0  # empty clip is a no-op
    tilted = _cabi.make_cameras(synth.K_CALIB, R, t, [0.1] * 14)
    assert lib.ska_triangulate_reproject_f32(*args(cams=tilted)) == -2
    with pytest.raises(ValueError):
        _cabi.make_cameras(synth.K_CALIB, R, t, [0.1] * 6)
    # per-frame extrinsics: workspace contract
    assert lib.ska_tri_frames_workspace_bytes(2, 10) > 0 and lib.ska_tri_frames_workspace_bytes(9, 10) == 0
    # (frame-major input takes the general two-kernel form, which needs the workspace; the fused kernel of the view-major
    #  layout needs none)
    fargs = [cams, 2, fake, fake, None, 4, 17, 1, 0, fake, None, None, None, fake, 16, None]
    assert lib.ska_triangulate_reproject_frames_f32(*fargs) == -4  # workspace too small
    fargs[2] = None
    assert lib.ska_triangulate_reproject_frames_f32(*fargs) == -1
    # standalone reprojection / statistics / losses
    assert lib.ska_reproject_points_f32(cams, 2, fake, None, 1, 17, 0, None, None, None) == -1  # nothing to compute
    assert lib.ska_reproject_points_f32(cams, 2, fake, None, 1, 17, 0, None, fake, None) == -1  # err needs kpts
    assert lib.ska_reproject_points_f32(cams, 2, fake, None, 0, 17, 0, fake, None, None) == 0
    assert lib.ska_frame_stats_f32(None, 0, 17, 2, 0, None, None) == 0
    assert lib.ska_project_points_f32(fake, 1, 17, 2, fake, 5, fake, 0, fake, 0, fake, None) == -1  # bad stride
    assert lib.ska_project_points_f64(fake, 0, 17, 2, fake, 0, fake, 0, fake, 0, fake, None) == 0
    assert lib.ska_reprojection_loss_f32(fake, 1, 17, 2, fake, 0, fake, 0, fake, 0, fake, fake, None, None, None, None, None,
                                         fake, 0, None) == -1
    assert lib.ska_loss_workspace_bytes(2) > 0 and lib.ska_reg_workspace_bytes() > 0
    # post-triangulation triage / smoothing
    assert lib.ska_post_triage_f32(cams, fake, fake, None, 1, 17, 4, 0.3, 2.0, fake, None, None, None) == -1  # unknown flag
    assert lib.ska_post_triage_f32(cams, fake, fake, None, 0, 17, 0, 0.3, 2.0, fake, None, None, None) == 0
    assert lib.ska_savgol_workspace_bytes(1000, 51) >= 1000 * 51 * 4
    assert lib.ska_savgol_f32(fake, 10, 51, 8, 2, C.c_void_p(512), fake, 1 << 30, None) == -2   # even window
    assert lib.ska_savgol_f32(fake, 10, 51, 9, 9, C.c_void_p(512), fake, 1 << 30, None) == -1   # poly >= window
    assert lib.ska_savgol_f32(fake, 10, 51, 9, 2, C.c_void_p(512), fake, 16, None) == -4        # workspace
    assert lib.ska_savgol_f32(fake, 10, 51, 9, 2, fake, fake, 1 << 30, None) == -1              # aliasing


def test_regularised_ba_argument_errors_without_gpu():
    """ska_ba_reg_*: validation happens before any CUDA call."""
    from skiing_analysis_pytorch_b200 import _cabi, _lib

    lib = _lib.load()
    fake = 256

    def prob(**kw):
        base = dict(C=2, J=17, n_bones=12, free_mask=0, T_local=10, has_prev=0, has_next=0, d_x2d=fake, d_conf=fake, d_K=fake, d_X=fake,
                    d_cams=fake, d_vec=fake, d_pinv=fake, d_lfac=None, d_sc=fake, d_sums=fake, d_hist=None, hist_rows=0, d_workspace=fake,
                    ws_bytes=1 << 30)
        base.update(kw)
        p = _cabi.SkaBaRegProblem(**base)
        for b in range(12):
            p.bone_i[b], p.bone_j[b] = b, b + 1
        return p

    assert lib.ska_ba_reg_workspace_bytes(0) == 0 and lib.ska_ba_reg_workspace_bytes(100) >= 100 * 42 * 8
    assert lib.ska_ba_reg_linearize_f64(None, None) == -1
    for bad in (dict(C=0), dict(C=9), dict(J=0), dict(J=97), dict(n_bones=17), dict(T_local=0), dict(d_x2d=None), dict(d_sc=None),
                dict(free_mask=0x3F)):  # free cameras need the Schur factor buffer
        assert lib.ska_ba_reg_linearize_f64(C.byref(prob(**bad)), None) == -1, bad
    assert lib.ska_ba_reg_cost_f64(C.byref(prob(ws_bytes=16)), 0, None) == -4
    p = prob()
    p.bone_j[3] = 17
    assert lib.ska_ba_reg_cost_f64(C.byref(p), 0, None) == -1 and b"bone" in lib.ska_last_error()


def test_api_refuses_cpu_tensors():
    import torch

    from skiing_analysis_pytorch_b200 import api, synth

    R, t = synth.rig("2b")
    with pytest.raises(RuntimeError, match="no CPU path"):
        api.triangulate_reproject(torch.zeros(2, 1, 17, 2), synth.K_CALIB, R, t)


def test_header_is_plain_c_and_links(tmp_path):
    """include/ska.h compiles as C11 (-Wall -Werror -pedantic) and a C program links against libska.so and calls it."""
    import shutil
    import subprocess

    from skiing_analysis_pytorch_b200 import _lib, build

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    build.build()
    exe = tmp_path / "link_check"
    libdir = _lib.LIB_PATH.parent
    cmd = [gcc, "-std=c11", "-Wall", "-Werror", "-pedantic", "-I", str(ROOT / "include"), str(ROOT / "tests" / "cabi" / "link_check.c"),
           "-L", str(libdir), "-lska", f"-Wl,-rpath,{libdir}", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok abi"), (r.returncode, r.stdout, r.stderr)
