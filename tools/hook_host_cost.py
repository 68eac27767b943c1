#!/usr/bin/env python
"""Host-side cost of one `compact_fwd` hook call (CPU only): bench.py --api dropin driven over the stand-in CUDA
library of tests/test_bench_dry_run.py under cProfile.  The C calls cost nothing here, so what is listed is the
Python the hooks add per layer (the part a GPU cannot hide once the kernels of a layer are shorter than it).
    python tools/hook_host_cost.py [--gpus 2] [--layers 57] [--steps 30]"""
import argparse
import cProfile
import io
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--layers", type=int, default=57)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--top", type=int, default=35)
    ap.add_argument("--plain", action="store_true", help="no cProfile: wall time per hook call (perf_counter)")
    a = ap.parse_args()
    import pytest
    import test_bench_dry_run as t

    class Cap:
        def readouterr(self):
            class R:
                out = "{}"
            return R()
    mp = pytest.MonkeyPatch()
    import bench
    real_print = print
    lines = []
    mp.setattr("builtins.print", lambda *x, **k: lines.append(" ".join(map(str, x))) if not k.get("file") else None)

    class Cap2:
        def readouterr(self):
            class R:
                out = "\n".join(lines)
            return R()
    argv = ["--gpus", str(a.gpus), "--layers", str(a.layers), "--steps", str(a.steps), "--no-cpu-baseline", "--no-gpu-reference",
            "--no-e2e", "--no-parity", "--api", "dropin"]
    pr = cProfile.Profile()
    orig = bench.DropinDriver.step if hasattr(bench, "DropinDriver") else None
    calls = {"n": 0}
    if orig is not None:
        def step(self, *x, **k):
            calls["n"] += 1
            if a.plain:
                import time
                t0 = time.perf_counter()
                try:
                    return orig(self, *x, **k)
                finally:
                    calls.setdefault("t", []).append(time.perf_counter() - t0)
            pr.enable()
            try:
                return orig(self, *x, **k)
            finally:
                pr.disable()
        mp.setattr(bench.DropinDriver, "step", step)
    # the hook modules reach torch.distributed through their own `dist`: forward to the stand-in bench.py gets
    import types
    from compactfusion_b200 import dropin
    import compactfusion_b200.patchpara.fwd as fwd_mod
    fwd = types.SimpleNamespace(**{n: (lambda *x, _n=n, **k: getattr(bench.dist, _n)(*x, **k))
                                   for n in ("get_world_size", "get_rank", "is_initialized", "all_gather_into_tensor", "barrier")})
    for mod in (dropin, fwd_mod):
        if hasattr(mod, "dist"):
            mp.setattr(mod, "dist", fwd)
    def usable(cfg, ctype, k):   # dropin.usable without `k.is_cuda` (the stand-in tensors live on the CPU)
        types_ = dropin._config_ok(cfg)
        if types_ is None or ctype not in types_:
            return False
        c = k.shape[-2] * k.shape[-1]
        return k.dtype == bench.torch.half and k.dim() == 4 and c % 128 == 0 and 64 <= c <= 8192
    mp.setattr(dropin, "usable", usable)
    real_hot = dropin.hot

    def hot(cfg, kind, group, k, mod_idx, ctype):   # the same probe without `k.is_cuda`
        ent = dropin._hot.get((kind, id(group), mod_idx, k.shape))
        if ent is not None and ent[3] == id(cfg) and k.dtype is bench.torch.half:
            return ent if ctype in ent[4] else None
        return real_hot(cfg, kind, group, k, mod_idx, ctype)
    mp.setattr(dropin, "hot", hot)
    try:
        line = t._run_bench(mp, Cap2(), argv, world=a.gpus)
    finally:
        mp.undo()
    if a.plain:
        ts = sorted(calls["t"])
        real_print(f"{len(ts)} steps x {a.layers} layers: median {ts[len(ts) // 2] / a.layers * 1e6:.1f} us, fastest "
                   f"{ts[0] / a.layers * 1e6:.1f} us of Python per hook call (stand-in C library: its calls cost ~1 us)")
        return
    s = io.StringIO()
    st = pstats.Stats(pr, stream=s).sort_stats("tottime")
    st.print_stats(a.top)
    total = sum(v[2] for v in st.stats.values())
    real_print(f"{calls['n']} steps x {a.layers} layers: {total / max(calls['n'] * a.layers, 1) * 1e6:.1f} us of Python per hook call "
               f"(cProfile inflates it about 2x)")
    real_print(s.getvalue()[:6000])


if __name__ == "__main__":
    main()
