"""Compression statistics of the compact plugin (mirror of xfuser/compact/stats.py: same module
functions, same per-record field names, same dump file formats).

The reference's `StatsLogger.log` (stats.py:107-328) computes every figure with eager torch
reductions followed by `.item()` -- a device synchronisation per figure, several per call -- and
parks the previous activation on the CPU.  Here a call only ENQUEUES one-pass reduction kernels
(`cf_error_stats`, csrc/cf_consumer.cu) that write 4 floats into a preallocated device table; nothing
synchronises inside the denoising loop, the previous activation stays in HBM, and the table is read
back once when a summary or dump is requested.

Per (key, step) record, reference field names:
  error                      ||x - recv||_2                      (stats.py:150)
  activation_norm            ||x||_2                             (stats.py:221)
  delta_norm                 ||x - base||_2         residual >= 1 (stats.py:205,222)
  delta_delta_norm           ||x - base - dbase||_2 residual == 2 (stats.py:212,223)
  delta_before_feedback_norm ||x - x_prev||_2                    (stats.py:193-194)
  activation_similarity      cos(x, x_prev), from ||x||, ||x_prev||, ||x - x_prev|| (stats.py:233-239)
  residual, original_size_bytes, compressed_size_bytes           (stats.py:164-169)
plus max_abs_error, rel_l2 and psnr_db (BASELINE.json north_star's per-step parity figures).
Not computed (always None): total_error (needs activation dumps of an uncompressed run),
delta / transmitted-delta / low-rank similarities; eigenvalue plots need matplotlib, which this
image does not have -- `plot_eigenvalues` / `save_eigenvalues` are accepted and do nothing.
"""
from __future__ import annotations

import math
import os

import torch

_CHUNK_ROWS = 8192
_NONE_FIELDS = ("total_error", "delta_similarity", "delta_before_feedback_similarity",
                "delta_before_feedback_lowrank_similarity", "transmitted_delta_similarity")


def stats_hello():
    print("compact stats: per-call error / norm figures are reduced on the GPU (cf_error_stats) and read back "
          "once per summary; eigenvalue plots are unavailable (no matplotlib)")


def _mean(vals):
    vals = [v for v in vals if v is not None]
    return sum(vals) / len(vals) if vals else None


class StatsLogger:
    """Statistics logger for compression metrics (API of the reference's class of the same name)."""

    def __init__(self):
        self._records = {}        # key -> list of per-step dicts; row handles until `_flush`
        self._tables = []         # device tables (_CHUNK_ROWS, 4) fp32
        self._used = 0            # rows handed out in the last table
        self._dirty = False
        self.prev_activations = {}
        self.step_counts = {}
        self.total_original_volume = 0
        self.total_compressed_volume = 0
        self.eigenvalues = {}

    # -- device side -----------------------------------------------------------------------
    def _pair(self, a: torch.Tensor, b: torch.Tensor):
        """Enqueue [sum (a-b)^2, sum b^2, max |a-b|, max |b|] into the next table row; returns its handle."""
        from .quality import error_stats_raw
        if not self._tables or self._used == _CHUNK_ROWS:
            self._tables.append(torch.zeros((_CHUNK_ROWS, 4), dtype=torch.float32, device=a.device))
            self._used = 0
        handle = (len(self._tables) - 1, self._used)
        error_stats_raw(a.contiguous().view(-1), b.contiguous().view(-1), out=self._tables[-1][self._used])
        self._used += 1
        return handle

    def _flush(self):
        """One read-back of all tables; row handles in the pending records become numbers."""
        if not self._dirty:
            return
        host = [t.tolist() for t in self._tables]  # synchronises once per table
        for recs in self._records.values():
            for r in recs:
                h = r.pop("_rows", None)
                if h is None:
                    continue
                get = lambda name: host[h[name][0]][h[name][1]] if name in h else None  # noqa: E731
                numel = r.pop("_numel")
                act = get("act")
                if act is None:                      # recv was None: only ||x|| via the prev / base pass, if any
                    act = get("base") or get("prev")
                r["activation_norm"] = math.sqrt(act[1]) if act is not None else None
                err = get("act")
                if err is not None:
                    sse, ssr, max_err, max_ref = err
                    r["error"] = math.sqrt(sse)
                    r["max_abs_error"] = max_err
                    r["rel_l2"] = math.sqrt(sse / ssr) if ssr > 0 else (0.0 if sse == 0 else math.inf)
                    mse = sse / numel
                    r["psnr_db"] = math.inf if mse == 0 else (10.0 * math.log10(max_ref * max_ref / mse) if max_ref > 0 else -math.inf)
                base = get("base")
                if base is not None:
                    r["delta_norm"] = math.sqrt(base[0])
                dd = get("dd")
                if dd is not None:
                    r["delta_delta_norm"] = math.sqrt(dd[0])
                prev = get("prev")
                if prev is not None:
                    # a = x_prev, b = x:  [||x_prev - x||^2, ||x||^2];  ||x_prev||^2 is the previous record's ||x||^2
                    d2, x2 = prev[0], prev[1]
                    p2 = r.pop("_prev_norm2_from")
                    p2 = p2["activation_norm"] ** 2 if p2 is not None and p2.get("activation_norm") is not None else None
                    r["delta_before_feedback_norm"] = math.sqrt(d2)
                    if p2 is not None and x2 > 0 and p2 > 0:
                        r["activation_similarity"] = (x2 + p2 - d2) / (2.0 * math.sqrt(x2) * math.sqrt(p2))
                r.pop("_prev_norm2_from", None)
        self._dirty = False

    @property
    def stats(self):
        """key -> list of per-step dicts (the reference's attribute of the same name)."""
        self._flush()
        return self._records

    # -- logging ---------------------------------------------------------------------------
    def log(self, key, base, delta_base, before_comp_activation, recv_activation, compressed_tensor,
            compress_residual):
        """Record one compress call (argument meaning as stats.py:107-127).  Enqueues kernels only."""
        if compress_residual not in (0, 1, 2):
            raise ValueError("invalid residual")
        x = before_comp_activation
        step = self.step_counts.get(key, 0)
        self.step_counts[key] = step + 1
        recs = self._records.setdefault(key, [])
        orig_bytes = x.numel() * x.element_size()
        comp_bytes = compressed_tensor.numel() * compressed_tensor.element_size() if compressed_tensor is not None else 0
        self.total_original_volume += orig_bytes
        self.total_compressed_volume += comp_bytes
        rows = {}
        if recv_activation is not None:
            rows["act"] = self._pair(recv_activation, x)          # error + ||x||
        if compress_residual >= 1 and base is not None:
            rows["base"] = self._pair(base, x)                    # ||x - base||
        if compress_residual == 2 and base is not None and delta_base is not None:
            rows["dd"] = self._pair(base + delta_base, x)         # ||x - base - delta_base||
        prev = self.prev_activations.get(key)
        if prev is not None and prev.shape == x.shape:
            rows["prev"] = self._pair(prev, x)                    # ||x - x_prev||, cos(x, x_prev)
            prev.copy_(x)                                         # stream-ordered behind the kernel above
        else:
            self.prev_activations[key] = x.detach().clone()
        rec = {"error": None, "activation_norm": None, "delta_norm": None, "delta_delta_norm": None,
               "delta_before_feedback_norm": None, "activation_similarity": None, "max_abs_error": None,
               "rel_l2": None, "psnr_db": None, "residual": compress_residual, "original_size_bytes": orig_bytes,
               "compressed_size_bytes": comp_bytes, "_rows": rows, "_numel": x.numel(),
               "_prev_norm2_from": recs[-1] if recs else None}
        for f in _NONE_FIELDS:
            rec[f] = None
        recs.append(rec)
        self._dirty = True

    def clear(self):
        self.__init__()

    # -- summaries (plain text; same content as stats.py:373-608) ----------------------------
    def summary_over_steps(self, steps=None, keys=None):
        stats = self.stats
        if not stats:
            print("No statistics logged yet.")
            return
        if keys is not None and not isinstance(keys, (list, tuple)):
            keys = [keys]
        avail = [k for k in (keys if keys is not None else stats.keys()) if k in stats]
        max_steps = max((len(stats[k]) for k in avail), default=0)
        for step in (steps if steps is not None else range(max_steps)):
            if step >= max_steps:
                print(f"Step {step} is out of range")
                continue
            print(f"=== Step {step} ===")
            for k in ([None] if keys is None else keys):
                self.summary_over_keys(step_range=(step, step + 1), key=k)

    def summary_over_keys(self, step_range=None, key=None):
        stats = self.stats
        if not stats:
            print("No statistics logged yet.")
            return
        for k in ([key] if key is not None else sorted(stats.keys())):
            if k not in stats:
                print(f"No statistics for key {k}")
                continue
            recs = stats[k]
            lo, hi = step_range if step_range is not None else (0, len(recs))
            sel = recs[lo:hi]
            if not sel:
                continue
            f = lambda name: _mean([r[name] for r in sel])  # noqa: E731
            fmt = lambda v, spec=".3f": "n/a" if v is None else format(v, spec)  # noqa: E731
            comp = sum(r["compressed_size_bytes"] for r in sel)
            ratio = sum(r["original_size_bytes"] for r in sel) / comp if comp else float("nan")
            print(f"[{k}] steps {lo}-{min(hi, len(recs)) - 1}: act {fmt(f('activation_norm'))}, delta {fmt(f('delta_norm'))}, "
                  f"dd {fmt(f('delta_delta_norm'))}, dbf {fmt(f('delta_before_feedback_norm'))}, err {fmt(f('error'))}, "
                  f"rel-l2 {fmt(f('rel_l2'), '.3e')}, max-abs {fmt(f('max_abs_error'), '.3e')}, psnr {fmt(f('psnr_db'), '.1f')} dB, "
                  f"act_sim {fmt(f('activation_similarity'))}, ratio {ratio:.2f}x")

    def summary_compression_volume(self):
        if self.total_original_volume == 0:
            print("No volume data logged yet.")
            return
        line = (f"Vol: Orig {self.total_original_volume / 2**20:.2f} MB, "
                f"Comp {self.total_compressed_volume / 2**20:.2f} MB")
        line += (f", Ratio {self.total_original_volume / self.total_compressed_volume:.2f}x"
                 if self.total_compressed_volume > 0 else ", Ratio N/A")
        print(line)

    def totals(self):
        """The averages `summary_total_avg` prints, as a dict (mean over keys of the per-key mean for
        activation norm and error, flat means elsewhere: stats.py:531-596)."""
        stats = self.stats
        per_key = lambda name: _mean([_mean([r[name] for r in recs]) for recs in stats.values()])  # noqa: E731
        flat = lambda name: _mean([r[name] for recs in stats.values() for r in recs])  # noqa: E731
        act, err = per_key("activation_norm"), per_key("error")
        return {"activation_norm": act, "delta_norm": flat("delta_norm"),
                "delta_before_feedback_norm": flat("delta_before_feedback_norm"),
                "delta_delta_norm": flat("delta_delta_norm"), "activation_similarity": flat("activation_similarity"),
                "error": err, "rel_error": (err / act if act and act > 1e-8 else math.inf) if err is not None else None,
                "rel_l2": flat("rel_l2"), "max_abs_error": max((r["max_abs_error"] for recs in stats.values() for r in recs
                                                               if r["max_abs_error"] is not None), default=None),
                "psnr_db": flat("psnr_db")}

    def summary_total_avg(self):
        if not self.stats:
            print("No statistics logged yet.")
            return
        t = self.totals()
        fmt = lambda v, spec=".3f": "n/a" if v is None else format(v, spec)  # noqa: E731
        print(f"avg activation: {fmt(t['activation_norm'])}, avg delta: {fmt(t['delta_norm'])}, avg dbf: "
              f"{fmt(t['delta_before_feedback_norm'])}, avg delta-delta: {fmt(t['delta_delta_norm'])}")
        if t["activation_similarity"] is not None:
            print(f"avg similarities: act_sim: {t['activation_similarity']:.3f}")
        print(f"avg comp error: {fmt(t['error'])}, avg rel err: {fmt(t['rel_error'], '.1%')}, avg rel-l2: "
              f"{fmt(t['rel_l2'], '.3e')}, worst max-abs: {fmt(t['max_abs_error'], '.3e')}, avg psnr: "
              f"{fmt(t['psnr_db'], '.1f')} dB [total err not logged]")

    # -- dumps (file names and dict keys of plot.py:413-560) ----------------------------------
    def _per_step(self, name):
        stats = self.stats
        max_steps = max((len(v) for v in stats.values()), default=0)
        return [_mean([recs[s][name] for recs in stats.values() if s < len(recs)]) for s in range(max_steps)]

    def dump_average_error_vs_steps(self, save_dir: str):
        assert self.stats, "No statistics logged. Cannot dump data."
        errs = self._per_step("error")
        data = {"steps": list(range(len(errs))), "avg_comp_errors": errs, "avg_total_errors": [None] * len(errs)}
        os.makedirs(save_dir, exist_ok=True)
        path = os.path.join(save_dir, "average_error_vs_steps.pt")
        torch.save(data, path)
        print(f"Saved average error data to {path}")
        return data

    def dump_average_norms_and_similarity_vs_steps(self, save_dir: str):
        assert self.stats, "No statistics logged. Cannot dump data."
        act = self._per_step("activation_norm")
        data = {"steps": list(range(len(act))), "avg_act_norms": act, "avg_delta_norms": self._per_step("delta_norm"),
                "avg_act_similarities": self._per_step("activation_similarity")}
        os.makedirs(save_dir, exist_ok=True)
        path = os.path.join(save_dir, "average_norms_and_similarity_vs_steps.pt")
        torch.save(data, path)
        print(f"Saved average norms and similarity data to {path}")
        return data

    def save_eigenvalues(self, save_dir="eigenvalues"):
        return None


_stats: StatsLogger | None = None


def stats_log() -> StatsLogger:
    global _stats
    if _stats is None:
        _stats = StatsLogger()
    return _stats


def stats_clear():
    global _stats
    _stats = None


def log(key, base, delta_base, real_activation, recv_activation, compressed_tensor, compress_residual):
    stats_log().log(key, base, delta_base, real_activation, recv_activation, compressed_tensor, compress_residual)


def stats_verbose(step_range=None, key=None, summary_keys=True):
    if _stats is None:
        print("No statistics logged.")
        return
    if summary_keys:
        _stats.summary_over_keys(step_range, key)
    _stats.summary_compression_volume()
    _stats.summary_total_avg()


def stats_verbose_steps(steps=None, keys=None):
    if _stats is None:
        print("No statistics logged.")
        return
    _stats.summary_over_steps(steps, keys)


def plot_eigenvalues(key=None, step=None, data_type="activation", save_dir=None, log_scale=True, top_k=None,
                     cum_sum=False):
    """Accepted for API compatibility; eigenvalue plots need matplotlib (absent from this image)."""
    return None


def save_eigenvalues(save_dir="eigenvalues"):
    return None


def dump_err_vs_steps(save_dir: str):
    if _stats is None:
        print("No statistics logged. Cannot dump data.")
        return None
    return _stats.dump_average_error_vs_steps(save_dir)


def dump_norms_sim_vs_steps(save_dir: str):
    if _stats is None:
        print("No statistics logged. Cannot dump data.")
        return None
    return _stats.dump_average_norms_and_similarity_vs_steps(save_dir)
