"""Scope profiler with the reference's API and scope names (mirror of xfuser/prof.py:5-203).

Same calls as the reference -- `Profiler.instance()`, `start/stop(name, stream=None, cpu=False)`,
`elapsed_time(name) -> (total_ms, avg_ms)`, `get_all_elapsed_times() -> (totals, avgs)`, `sync()`,
`reset()`, `Profiler.scope(...)`, `@Profiler.prof_func(...)`, `prof_summary(profiler, rank) -> lines`,
`set_torch_profiler` / `torch_profiler_step` -- so xDiT's hooks, the example scripts and latency
breakdowns keyed by scope name (`compact.compact_compress`, `compact.all_gather`, `compact.ring.wait`,
...) work unchanged.

Like the reference's, the profiler is ENABLED from the start (its example scripts rely on that:
`Profiler.instance().reset(); with Profiler.instance().scope("total"): ...; prof_summary(...)` without ever
calling `enable()`), and every scope records two CUDA events.  Two things differ, so that the same code also
runs where the reference's cannot: on a host without CUDA a section is timed with the wall clock instead of
raising, and while the current stream is being captured into a CUDA graph a section records nothing
(timing events cannot be read back from a capture).  `COMPACT_PROFILER=0` in the environment, or
`Profiler.instance().disable()`, switches the ~2 us per event off; the engines (engine.py) never use it.
"""
from __future__ import annotations

import functools
import os
import time

import torch


class Profiler:
    _instance = None

    def __init__(self):
        self.events = {}      # name -> {'start': [...], 'end': [...], 'elapsed': ms, 'count': n, 'cpu': bool}
        self.enabled = os.environ.get("COMPACT_PROFILER", "1") != "0"

    @staticmethod
    def instance() -> "Profiler":
        if Profiler._instance is None:
            Profiler._instance = Profiler()
        return Profiler._instance

    def enable(self):
        """All subsequent start / stop calls are recorded."""
        self.enabled = True

    def disable(self):
        """All subsequent start / stop calls are ignored."""
        self.enabled = False

    # -- recording ---------------------------------------------------------------------------
    @staticmethod
    def _mark(stream, cpu):
        """A time stamp: wall clock (cpu sections, or no CUDA on this host), a recorded CUDA event, or None
        while the stream is being captured into a graph."""
        if cpu or not torch.cuda.is_available():
            return time.time()
        if torch.cuda.is_current_stream_capturing():
            return None
        ev = torch.cuda.Event(enable_timing=True)
        if stream is not None:
            ev.record(stream)
        else:
            ev.record()
        return ev

    def start(self, name, stream=None, cpu=False):
        """Open section `name` (a section may be opened and closed many times; times accumulate)."""
        if not self.enabled:
            return
        rec = self.events.setdefault(name, {"start": [], "end": [], "elapsed": 0.0, "count": 0, "cpu": cpu})
        assert len(rec["start"]) == len(rec["end"]), f"Cannot start '{name}' as there are more starts than stops"
        rec["start"].append(self._mark(stream, cpu))

    def stop(self, name, stream=None, cpu=False):
        if not self.enabled:
            return
        assert name in self.events, f"No events recorded for '{name}'"
        rec = self.events[name]
        assert len(rec["start"]) - 1 == len(rec["end"]), f"Cannot stop '{name}' as there are more stops than starts"
        rec["end"].append(self._mark(stream, cpu))

    # -- read-out ----------------------------------------------------------------------------
    def elapsed_time(self, name):
        """(total_ms, avg_ms) of section `name`; folds the pending start / stop pairs into the total
        (synchronising the device for CUDA-event sections)."""
        if name not in self.events:
            raise ValueError(f"No events recorded for '{name}'")
        rec = self.events[name]
        pairs = list(zip(rec["start"], rec["end"]))
        if pairs:
            timed = [(s, e) for s, e in pairs if s is not None and e is not None]  # None: inside a graph capture
            if any(not isinstance(s, float) for s, _ in timed):
                torch.cuda.synchronize()
            for s, e in timed:
                rec["elapsed"] += (e - s) * 1000.0 if isinstance(s, float) else s.elapsed_time(e)
            rec["count"] += len(timed)
            # an open section (start without stop) stays open
            rec["start"], rec["end"] = rec["start"][len(pairs):], []
        return rec["elapsed"], (rec["elapsed"] / rec["count"] if rec["count"] else 0.0)

    def get_all_elapsed_times(self):
        totals, avgs = {}, {}
        for name in self.events:
            totals[name], avgs[name] = self.elapsed_time(name)
        return totals, avgs

    def sync(self):
        """Fold every pending event pair into the totals."""
        self.get_all_elapsed_times()

    def reset(self):
        self.events = {}

    def summary(self):
        """name -> (total_ms, calls)."""
        out = {}
        for name in sorted(self.events):
            total, _ = self.elapsed_time(name)
            out[name] = (total, self.events[name]["count"])
        return out

    # -- scopes ------------------------------------------------------------------------------
    class _Scope:
        def __init__(self, profiler, name, stream=None, cpu=False):
            self.profiler, self.name, self.stream, self.cpu = profiler, name, stream, cpu

        def __enter__(self):
            self.profiler.start(self.name, self.stream, self.cpu)
            return self

        def __exit__(self, exc_type, exc_val, exc_tb):
            self.profiler.stop(self.name, self.stream, self.cpu)
            return False

    @staticmethod
    def scope(name, stream=None, cpu=False):
        """`with Profiler.scope("compact.ring.wait"): ...`"""
        return Profiler._Scope(Profiler.instance(), name, stream, cpu)

    @staticmethod
    def prof_func(name, cpu=False):
        """Decorator: time every call of the function as section `name`."""
        def decorator(func):
            @functools.wraps(func)
            def wrapper(*args, **kwargs):
                inst = Profiler.instance()
                if not inst.enabled:
                    return func(*args, **kwargs)
                inst.start(name, cpu=cpu)
                try:
                    return func(*args, **kwargs)
                finally:
                    inst.stop(name, cpu=cpu)
            return wrapper
        return decorator


def prof_summary(profiler: Profiler | None = None, rank=None):
    """Breakdown as a list of text lines, largest section first; shares are relative to the section
    named 'total' when one was recorded (xfuser/prof.py:172-189)."""
    p = profiler or Profiler.instance()
    rank = "N/A" if rank is None else rank
    totals, avgs = p.get_all_elapsed_times()
    whole = totals.get("total", 0.0)  # (the reference raises without a 'total' section; here shares are omitted)
    split = "-" * 20
    lines = [split, f"Profiling Summary for Rank {rank}"]
    for name, ms in sorted(totals.items(), key=lambda kv: kv[1], reverse=True):
        share = f" {ms / whole:.2%}" if whole > 0 else ""
        lines.append(f"[Rank {rank}] [{name}] {ms / 1000:.2f}s{share} avg={avgs[name]:.2f}ms")
    lines.append(split)
    return lines


_torch_profiler = None


def torch_profiler_step():
    """Advance the torch.profiler schedule installed with `set_torch_profiler`, if any."""
    if _torch_profiler is not None:
        _torch_profiler.step()


def set_torch_profiler(profiler):
    global _torch_profiler
    _torch_profiler = profiler
