"""Build libska.so (hand-written sm_100a CUDA behind the C ABI of include/ska.h) in-tree.

nvcc cross-compiles without a GPU; the resulting lib/libska.so travels with the repo snapshot to
the GPU box (it is git-ignored, not gpurun-ignored).  No torch linkage: the boundary is plain C,
loaded with ctypes (see _lib.py).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OBJ = CSRC / "_obj"
LIB = PKG / "lib" / "libska.so"
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xfatbin=-compress-all"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libska.so cannot be built (no CPU fallback exists)")


def sources():
    return sorted(CSRC.glob("*.cu"))


def _dep_files():
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [ROOT / "include" / "ska.h"])


def source_hash() -> str:
    """sha256 over the sources and flags the library is built from.  The library carries it in a side file
    (lib/libska.so.hash, written after a successful link): staleness is decided by CONTENT, not by mtimes - a snapshot
    copied to another machine keeps its library whatever the copy did to the timestamps."""
    import hashlib

    h = hashlib.sha256()
    h.update(" ".join(ARCH_FLAGS + NVCC_FLAGS).encode())
    for f in _dep_files():
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()


def _hash_file() -> Path:
    return LIB.with_name(LIB.name + ".hash")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    try:
        return _hash_file().read_text().strip() != source_hash()
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: Path | None = None) -> Path:
    """out/extra_flags: build an experimental variant next to the product library (tools/variants.py)."""
    global LIB, OBJ
    if out is not None:
        saved = (LIB, OBJ)
        LIB, OBJ = Path(out), CSRC / ("_obj_" + Path(out).stem)
        try:
            return build(force=True, verbose=verbose, extra_flags=extra_flags)
        finally:
            LIB, OBJ = saved
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    OBJ.mkdir(parents=True, exist_ok=True)
    LIB.parent.mkdir(parents=True, exist_ok=True)
    import fcntl

    with open(LIB.parent / ".build.lock", "w") as lock:  # one builder at a time (N ranks may import at once)
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not needs_build():  # somebody else built it while we waited
            return LIB
        return _build_locked(nvcc, force, verbose, extra_flags)


def _build_locked(nvcc, force, verbose, extra_flags) -> Path:
    import hashlib

    hdr = hashlib.sha256(" ".join([*ARCH_FLAGS, *NVCC_FLAGS, *extra_flags]).encode())
    for f in sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [ROOT / "include" / "ska.h"]):
        hdr.update(f.name.encode())
        hdr.update(f.read_bytes())
    hdr_digest = hdr.hexdigest()

    def compile_one(src: Path) -> Path:
        obj = OBJ / (src.stem + ".o")
        tag = obj.with_name(obj.name + ".hash")
        want = hashlib.sha256((hdr_digest + src.name).encode() + src.read_bytes()).hexdigest()
        try:  # an object is reused iff it was compiled from exactly these bytes with exactly these flags (content, not mtimes)
            fresh = obj.exists() and tag.read_text().strip() == want
        except OSError:
            fresh = False
        if not force and fresh:
            return obj
        cmd = [nvcc, *ARCH_FLAGS, *NVCC_FLAGS, *extra_flags, "-I", str(ROOT / "include"), "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose:
            sys.stderr.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n{r.stdout}\n{r.stderr}")
        tag.write_text(want + "\n")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    tmp = LIB.with_suffix(".so.tmp")
    cmd = [nvcc, *ARCH_FLAGS, "-shared", "-o", str(tmp), *map(str, objs), "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)
    if not extra_flags:
        _hash_file().write_text(source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
