"""Oracle verdict on ONE compressed exchange of the fastpath codecs (TEST INFRASTRUCTURE ONLY, see
oracle/__init__.py): used by the multi-rank parity tests and by `bench.py`'s parity leg (outside the
timed region), never by the product.

A receiver holds, for every origin r of a layer: the base it had cached before the step, the wire payload
[codes | U | V] that arrived from r (main.py:149-152), and the reconstruction it wrote.  The reference's
invariant (main.py:398-419, ring.py:184-200; utils.py:164-196 is its own check) is that every rank ends
with the same `new_base` for origin r, namely what `compact_decompress` computes from that payload.
Checked here, per origin:

  * sign bits (BINARY) of the payload == (x - base >= 0) computed by the oracle: bit-exact
    (fastpath.py:58-72);
  * V (a mean over N) within 1 fp16 ulp of the oracle's, U within 2: mean reductions leave the fp32 summation
    order free (SURVEY.md section 7 hard part 2), so a row mean and the token mean it is divided by can each
    sit one fp16 step away from the oracle's, and their quotient U = rowmean / mean(rowmean)
    (fastpath.py:164-165) two;
  * INT2 code bytes == the oracle's codes GIVEN the payload's scales: bit-exact (fastpath.py:529-549);
  * reconstruction == oracle dequant of THAT payload against THAT base: bit-exact
    (fastpath.py:328-363 / :672-741).
"""
from __future__ import annotations

import numpy as np
import torch

from . import codecs


def _ulp_diff(a: torch.Tensor, b: torch.Tensor) -> int:
    """Largest distance in fp16 representation steps (same-sign positive scales)."""
    return int((a.contiguous().view(torch.int16).int() - b.contiguous().view(torch.int16).int()).abs().max())


def split_payload(payload_u8: np.ndarray, n: int, c: int, ctype: str):
    """[codes (n, c/per_byte) u8 | U (n) fp16 | V (c) fp16] -> (codes ndarray, U (n,1), V (c,1))."""
    per_byte = 8 if ctype == "binary" else 4
    code_b = n * c // per_byte
    assert payload_u8.dtype == np.uint8 and payload_u8.size >= code_b + 2 * (n + c)
    codes = payload_u8[:code_b].reshape(n, c // per_byte)
    u = torch.from_numpy(payload_u8[code_b:code_b + 2 * n].copy().view(np.int16)).view(torch.half).view(n, 1)
    v = torch.from_numpy(payload_u8[code_b + 2 * n:code_b + 2 * (n + c)].copy().view(np.int16)).view(torch.half).view(c, 1)
    return codes, u, v


def check_origin(ctype: str, x: torch.Tensor | None, base: torch.Tensor, payload_u8: np.ndarray,
                 recon: torch.Tensor) -> dict:
    """One origin's tensor.  x (the origin's raw shard) may be None when it is not available on the checking
    rank: the sender-side checks are then skipped.  All tensors CPU fp16 (n, c)."""
    assert ctype in ("binary", "int2")
    n, c = base.shape
    codes, u, v = split_payload(payload_u8, n, c, ctype)
    out = {"n": n, "c": c}
    dequant = codecs.binary_dequant if ctype == "binary" else codecs.int2_dequant
    want = dequant(codes, u, v, base)
    bad = int((want.view(torch.int16) != recon.view(torch.int16)).sum())
    out["recon_mismatch"] = bad
    out["scales_finite"] = bool(torch.isfinite(u.float()).all() and torch.isfinite(v.float()).all())
    if x is not None:
        quant = codecs.binary_quant if ctype == "binary" else codecs.int2_quant
        o_codes, o_u, o_v, _ = quant(x, base, False)
        out["u_ulp"], out["v_ulp"] = _ulp_diff(u, o_u), _ulp_diff(v, o_v)
        if ctype == "binary":
            out["code_mismatch"] = int((o_codes != codes).sum())
        else:
            # magnitude bits depend on the thresholds: bit-exact given the payload's own scales
            g_codes, _, _, _ = quant(x, base, False, scales=(u, v))
            out["code_mismatch"] = int((g_codes != codes).sum())
            out["code_mismatch_end_to_end_frac"] = float((o_codes != codes).mean())
    out["ok"] = (bad == 0 and out["scales_finite"] and out.get("code_mismatch", 0) == 0
                 and out.get("u_ulp", 0) <= 2 and out.get("v_ulp", 0) <= 1)
    return out


def check_exchange(ctype: str, xs, bases, payloads, recons) -> dict:
    """All origins of one tensor (lists indexed by origin; xs entries may be None).  Returns the merged verdict."""
    per = [check_origin(ctype, x, b, p, r) for x, b, p, r in zip(xs, bases, payloads, recons)]
    return {
        "ok": all(p["ok"] for p in per),
        "origins": len(per),
        "recon_mismatch": sum(p["recon_mismatch"] for p in per),
        "code_mismatch": sum(p.get("code_mismatch", 0) for p in per),
        "max_scale_ulp": max([max(p.get("u_ulp", 0), p.get("v_ulp", 0)) for p in per] or [0]),
        "max_u_ulp": max([p.get("u_ulp", 0) for p in per] or [0]),
        "max_v_ulp": max([p.get("v_ulp", 0) for p in per] or [0]),
        "elements": sum(p["n"] * p["c"] for p in per),
    }
