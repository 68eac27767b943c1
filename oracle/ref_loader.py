"""Import the *reference* xfuser.compact package from /root/reference on CPU.

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_goldens.py (run in the build
container, where /root/reference exists) to validate the oracle restatement and
to generate the golden vectors committed under tests/golden/.  Nothing in the
product package, the -m gpu tests, smoke() or bench.py imports this module:
/root/reference does not exist on the GPU box.

Recipe: skip xfuser/__init__.py (it pulls in diffusers)
by pre-seeding a namespace module, stub matplotlib (imported by
xfuser/compact/stats.py:5), run eager (TORCHDYNAMO_DISABLE=1: eager is the
semantics the reference's own tests treat as ground truth), disable the CUDA
event profiler and initialise a disabled collector (CompactCache.put calls it,
xfuser/compact/utils.py:138-143).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "xfuser", "compact"))


def load_reference():
    """Returns the imported reference `xfuser` namespace (compact.* loaded)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
    if "xfuser" in sys.modules and getattr(sys.modules["xfuser"], "_cf_ref_stub", False):
        return sys.modules["xfuser"]
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                stub = types.ModuleType(name)
                sys.modules[name] = stub
    if "matplotlib.pyplot" in sys.modules and "matplotlib" in sys.modules:
        setattr(sys.modules["matplotlib"], "pyplot", sys.modules["matplotlib.pyplot"])
    pkg = types.ModuleType("xfuser")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "xfuser")]
    pkg._cf_ref_stub = True
    sys.modules["xfuser"] = pkg
    import importlib

    prof = importlib.import_module("xfuser.prof")
    prof.Profiler.instance().disable()
    for sub in (
        "compress_lowrank",
        "compress_quantize",
        "compress_topk",
        "utils",
        "slowpath",
        "fastpath",
        "main",
    ):
        importlib.import_module(f"xfuser.compact.{sub}")
    coll = importlib.import_module("xfuser.collector.collector")
    try:
        coll.init(coll.Collector(save_dir="/tmp/cf_collector_unused", enabled=False))
    except TypeError:
        coll.init(coll.Collector(enabled=False))
    return pkg
