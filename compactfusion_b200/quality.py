"""Per-step fidelity figures of the compressed path, computed on the GPU in one pass
(`cf_error_stats`, csrc/cf_consumer.cu): max-abs error, relative L2 and PSNR of a tensor against
its reference -- the numbers BASELINE.json's north_star asks every parity report to carry, and
what the reference's StatsLogger derives from eager torch reductions (stats.py:44-120).

LPIPS needs pretrained networks that are not in this image (no network access): PSNR and
relative L2 on activations / latents only.
"""
from __future__ import annotations

import math

import torch

from . import _native as nv

_ws: dict = {}


def _workspace(device: torch.device) -> torch.Tensor:
    key = (device.index if device.index is not None else torch.cuda.current_device(), nv.stream_ptr())
    ws = _ws.get(key)
    if ws is None:
        ws = torch.zeros(int(nv.lib().cf_error_stats_workspace_bytes()), dtype=torch.uint8, device=device)
        _ws[key] = ws
    return ws


def error_stats_raw(test: torch.Tensor, ref: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Enqueue one `k_error_stats` pass; returns a 4-element fp32 DEVICE tensor
    [sum (test-ref)^2, sum ref^2, max |test-ref|, max |ref|] (no synchronisation)."""
    nv.require_cuda_half(test, "test")
    nv.require_cuda_half(ref, "ref")
    assert test.shape == ref.shape and test.is_contiguous() and ref.is_contiguous()
    assert test.numel() % 8 == 0 and test.numel() > 0, "numel must be a positive multiple of 8"
    if out is None:
        out = torch.empty(4, dtype=torch.float32, device=test.device)
    ws = _workspace(test.device)
    rc = nv.lib().cf_error_stats(test.data_ptr(), ref.data_ptr(), test.numel(), out.data_ptr(), ws.data_ptr(),
                                 ws.numel(), nv.stream_ptr())
    nv.check(rc, "cf_error_stats")
    return out


def summarize(raw, numel: int) -> dict:
    """[sse, ssr, max_err, max_ref] -> {'max_abs', 'rel_l2', 'psnr_db', 'mse'} (PSNR against the
    reference's peak magnitude)."""
    sse, ssr, max_err, max_ref = (float(v) for v in raw)
    mse = sse / numel
    rel = math.sqrt(sse / ssr) if ssr > 0 else (0.0 if sse == 0 else math.inf)
    if mse == 0:
        psnr = math.inf
    elif max_ref == 0:
        psnr = -math.inf
    else:
        psnr = 10.0 * math.log10(max_ref * max_ref / mse)
    return {"max_abs": max_err, "rel_l2": rel, "psnr_db": psnr, "mse": mse}


def error_stats(test: torch.Tensor, ref: torch.Tensor) -> dict:
    """Synchronising convenience wrapper: one kernel, one 16-byte D2H read."""
    return summarize(error_stats_raw(test, ref).tolist(), test.numel())


class QualityTrace:
    """Collects the per-step figures of a run without synchronising inside the denoising loop:
    `record` enqueues one kernel writing into a preallocated row; `rows()` reads them all back once."""

    def __init__(self, max_records: int, device: torch.device):
        self._buf = torch.zeros((max_records, 4), dtype=torch.float32, device=device)
        self._meta = []

    def record(self, tag, test: torch.Tensor, ref: torch.Tensor):
        i = len(self._meta)
        assert i < self._buf.shape[0], "QualityTrace is full"
        error_stats_raw(test, ref, out=self._buf[i])
        self._meta.append((tag, test.numel()))

    def rows(self):
        vals = self._buf[:len(self._meta)].tolist()
        return [dict(tag=tag, **summarize(v, n)) for (tag, n), v in zip(self._meta, vals)]
