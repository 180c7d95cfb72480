"""Drop-in replacements for the reference's hot-path modules.

The reference has no plugin layer: its pipelines import plain Python modules by absolute name
(``from triangulation.reproject import ...`` triangulation/triangulate.py:31; ``from vggt.triangulate
import ...`` vggt/multi_view_process.py:22) or relative name (``from .reproject import ...``
bundle_adjustment/run.py:43, front_side/side/run.py:28; ``from .loss import ...`` run.py:35), all of
which resolve through ``sys.modules``.  ``install()`` registers the shims under those names BEFORE
the pipeline's ``main`` is imported, so the front_side / fuse / vggt / triangulation pipelines run
unchanged while their arithmetic executes in libska.so on the GPU:

    import skiing_analysis_pytorch_b200.dropin as dropin
    dropin.install()
    runpy.run_module("triangulation.main", run_name="__main__")

Each shim keeps the reference's names, argument meaning, return types and error behaviour
(INTEGRATION.md lists them with file:line).  The overlay alternative - a one-line re-export file at
the reference's own path - is described there too.
"""
from __future__ import annotations

import importlib
import sys
import types

# reference module name -> shim module (relative to this package)
MODULE_MAP = {
    "triangulation.triangulate": "triangulate_two_view",
    "triangulation.reproject": "reproject_stereo",
    "triangulation.postprocess": "postprocess",
    "bundle_adjustment.loss": "ba_loss",
    "bundle_adjustment.reproject": "reproject_world",
    "vggt.triangulate": "triangulate_vggt",
    "vggt.reproject": "reproject_world",
    "front_side.side.reproject": "reproject_world",
    "fuse.side.reproject": "reproject_world",
    "bundle_adjustment.fuse.fuse": "rigid_fuse",
    "fuse.side.fuse.fuse": "rigid_fuse",
    "front_side.side.fuse.fuse": "rigid_fuse",
}


def shim(reference_name: str) -> types.ModuleType:
    """The shim module that stands in for `reference_name` (e.g. 'triangulation.reproject')."""
    return importlib.import_module(f"{__name__}.{MODULE_MAP[reference_name]}")


def install(names=None) -> dict:
    """Pre-register the shims in sys.modules under the reference's module names.

    Parent packages that cannot be imported (no reference checkout on sys.path) are created as
    namespace stubs so ``import triangulation.reproject`` still resolves.  Returns {name: module}.
    Also publishes ``run_local_ba`` - the optimiser vggt/multi_view_process.py:553 calls but the
    reference never defines - as ``bundle_adjustment.run_local_ba`` and ``vggt.run_local_ba``."""
    out = {}
    for name in names or MODULE_MAP:
        mod = shim(name)
        parts = name.split(".")
        for k in range(1, len(parts)):
            pkg = ".".join(parts[:k])
            if pkg not in sys.modules:
                try:
                    importlib.import_module(pkg)
                except Exception:
                    stub = types.ModuleType(pkg)
                    stub.__path__ = []  # mark as package
                    sys.modules[pkg] = stub
                    if k > 1:
                        setattr(sys.modules[".".join(parts[: k - 1])], parts[k - 1], stub)
        sys.modules[name] = mod
        setattr(sys.modules[".".join(parts[:-1])], parts[-1], mod)
        out[name] = mod
    from ..ba import run_local_ba

    for pkg in ("bundle_adjustment", "vggt"):
        if pkg in sys.modules:
            setattr(sys.modules[pkg], "run_local_ba", run_local_ba)
    return out


def uninstall():
    for name in MODULE_MAP:
        mod = sys.modules.get(name)
        if mod is not None and getattr(mod, "__name__", "").startswith(__name__):
            del sys.modules[name]
