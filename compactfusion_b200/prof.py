"""CUDA-event scope profiler with the reference's scope names (xfuser/prof.py:5-202).

Disabled by default (the reference's is enabled by default and creates CUDA events on every
scope); `Profiler.instance().enable()` turns it on.  Scope names on the hot path are kept
(`compact.compact_compress`, `compact.all_gather`, `compact.ring.wait`, ...) so latency
breakdowns are comparable.
"""
from __future__ import annotations

import contextlib
import functools
from collections import defaultdict

import torch


class Profiler:
    _instance = None

    def __init__(self):
        self.enabled = False
        self._open = {}
        self._pairs = defaultdict(list)
        self._totals = defaultdict(float)
        self._counts = defaultdict(int)

    @classmethod
    def instance(cls):
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    def enable(self):
        self.enabled = True

    def disable(self):
        self.enabled = False

    def reset(self):
        self._open.clear()
        self._pairs.clear()
        self._totals.clear()
        self._counts.clear()

    def start(self, name):
        if not self.enabled:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self._open[name] = ev

    def stop(self, name):
        if not self.enabled or name not in self._open:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self._pairs[name].append((self._open.pop(name), ev))

    def elapsed_time(self, name):
        """Accumulated milliseconds and call count of a scope (synchronises)."""
        for s, e in self._pairs.pop(name, []):
            e.synchronize()
            self._totals[name] += s.elapsed_time(e)
            self._counts[name] += 1
        return self._totals[name], self._counts[name]

    def summary(self):
        names = set(self._pairs) | set(self._totals)
        return {n: self.elapsed_time(n) for n in sorted(names)}

    @classmethod
    @contextlib.contextmanager
    def scope(cls, name):
        inst = cls.instance()
        inst.start(name)
        try:
            yield
        finally:
            inst.stop(name)

    @classmethod
    def prof_func(cls, name):
        def deco(fn):
            @functools.wraps(fn)
            def wrapper(*a, **kw):
                inst = cls.instance()
                if not inst.enabled:
                    return fn(*a, **kw)
                with cls.scope(name):
                    return fn(*a, **kw)
            return wrapper
        return deco


def prof_summary(profiler: Profiler | None = None, rank=None) -> str:
    p = profiler or Profiler.instance()
    rows = p.summary()
    total = rows.get("total", (0.0, 0))[0]
    lines = []
    for name, (ms, cnt) in rows.items():
        share = f" {100.0 * ms / total:5.1f}%" if total > 0 else ""
        lines.append(f"{name:48s} total {ms:10.3f} ms  calls {cnt:6d}  avg {ms / max(cnt, 1):8.4f} ms{share}")
    return "\n".join(lines)
