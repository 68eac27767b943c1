"""Install compactfusion_b200 under the reference's module path `xfuser.compact`.

xDiT's hooks import the plugin by name:
    from xfuser.compact.main import compact_config, compact_get_step, ...   (attention_processor.py:43,
        hybrid/attn_layer.py:59,184, cache_manager.py:73, pipeline_flux.py:482)
    from xfuser.compact.ring import compact_fwd                            (hybrid/attn_layer.py:60)
    from xfuser.compact.utils import CompactConfig, COMPACT_COMPRESS_TYPE  (examples/*_example.py)
    from xfuser.compact.patchpara.df_utils import PatchConfig

`install()` registers this package's modules in `sys.modules` under those names, so an
unmodified xDiT checkout picks up the sm_100a kernels.  Call it BEFORE the first
`import xfuser.compact...` (i.e. at the top of the example script, before the pipeline is
built -- the same place `compact_init` is called, examples/flux_example.py:62-136).

    import compactfusion_b200.shim as shim; shim.install()

If `xfuser` itself is importable its other subpackages are untouched; if it is not (as in
this repo's tests) a namespace stub is created so `xfuser.compact.*` resolves on its own.
"""
from __future__ import annotations

import importlib
import sys
import types

_SUBMODULES = (
    "main", "ring", "utils", "fastpath", "slowpath", "compress_quantize", "compress_topk", "compress_lowrank",
    "stats", "prof", "attention", "patchpara", "patchpara.df_utils", "patchpara.df_cache", "patchpara.fwd",
)


def install(force: bool = True) -> types.ModuleType:
    """Alias compactfusion_b200[.sub] as xfuser.compact[.sub]; returns the package module.
    With force=False an already imported `xfuser.compact` is left alone."""
    pkg = importlib.import_module("compactfusion_b200")
    if not force and "xfuser.compact" in sys.modules:
        return sys.modules["xfuser.compact"]
    parent = sys.modules.get("xfuser")
    if parent is None:
        try:
            parent = importlib.import_module("xfuser")
        except Exception:  # xDiT (diffusers, yunchang ...) not importable: stand-alone stub
            parent = types.ModuleType("xfuser")
            parent.__path__ = []  # mark as package
            sys.modules["xfuser"] = parent
    sys.modules["xfuser.compact"] = pkg
    setattr(parent, "compact", pkg)
    for name in _SUBMODULES:
        mod = importlib.import_module(f"compactfusion_b200.{name}")
        sys.modules[f"xfuser.compact.{name}"] = mod
    # the reference keeps its profiler at xfuser/prof.py (imported as `from xfuser.prof import Profiler`)
    sys.modules.setdefault("xfuser.prof", importlib.import_module("compactfusion_b200.prof"))
    return pkg


def uninstall() -> None:
    for name in list(sys.modules):
        if name == "xfuser.compact" or name.startswith("xfuser.compact."):
            mod = sys.modules[name]
            if getattr(mod, "__name__", "").startswith("compactfusion_b200"):
                del sys.modules[name]
    prof = sys.modules.get("xfuser.prof")
    if prof is not None and getattr(prof, "__name__", "").startswith("compactfusion_b200"):
        del sys.modules["xfuser.prof"]
