"""Stand-alone quantisation codecs (mirror of xfuser/compact/compress_quantize.py).

Same function names, argument meaning and tensor layouts as the reference; the tensor work
runs in the sm_100a kernels of libcompactb200 (csrc/cf_sign_codecs.cu, cf_minmax_codecs.cu).
The `sim_*` functions return the dequantised tensor directly, like the reference's.
"""
from __future__ import annotations

import torch

from . import _native as nv
from .prof import Profiler


def _check2d(t: torch.Tensor, name="input_tensor"):
    nv.require_cuda_half(t, name)
    assert t.dim() == 2, f"{name} must be 2D"
    return t.contiguous()


def _sign_compress(codec, x, base, update_cache, packed=None, u=None, v=None, new_base=None):
    """Fused residual compress for BINARY / INT2.  Output tensors may be supplied (views into
    a wire payload); returns (packed, U (N,1), V (C,1), new_base|None)."""
    nv.require_cuda_half(x, "x")
    n, c = x.shape
    per_byte = 8 if codec == nv.CODEC_BINARY else 4
    assert c % 8 == 0, "C must be divisible by 8"
    dev = x.device
    if packed is None:
        packed = torch.empty((n, c // per_byte), dtype=torch.uint8, device=dev)
    if u is None:
        u = torch.empty((n, 1), dtype=torch.half, device=dev)
    if v is None:
        v = torch.empty((c, 1), dtype=torch.half, device=dev)
    if update_cache and new_base is None:
        new_base = torch.empty_like(x)
    ws_bytes = nv.workspace_bytes(codec, n, c)
    ws = nv.workspace(ws_bytes, dev)
    fn = nv.lib().cf_binary_compress if codec == nv.CODEC_BINARY else nv.lib().cf_int2_compress
    rc = fn(nv.ptr(x), nv.ptr(base), nv.ptr(new_base) if update_cache else None, nv.ptr(packed), nv.ptr(u),
            nv.ptr(v), n, c, nv.ptr(ws), ws.numel(), nv.stream_ptr())
    nv.check(rc, "cf_binary_compress" if codec == nv.CODEC_BINARY else "cf_int2_compress")
    return packed, u, v, (new_base if update_cache else None)


# ------------------------------------------------------------------------------------ 1-bit
def quantize_1bit(input_tensor: torch.Tensor, rank):
    """(N,C) fp16 -> packed (N,C/8) u8, scale_u (N,K), scale_v (K,C).  compress_quantize.py:7-90."""
    assert rank >= 1 or rank == -1, "Rank must be >= 1 or -1"
    x = _check2d(input_tensor)
    if rank == -1:
        packed, u, v, _ = _sign_compress(nv.CODEC_BINARY, x, None, False)
        return packed, u, v.view(1, -1)
    from .compress_lowrank import subspace_iter
    with Profiler.scope(f"compact.quant.scale_rank{rank}_approx"):
        su, svt, _ = subspace_iter(torch.abs(x), rank=rank, num_iters=2)
    packed, _, _, _ = _sign_compress(nv.CODEC_BINARY, x, None, False)
    return packed, su.contiguous().half(), svt.contiguous().half()


def dequantize_1bit(packed_tensor: torch.Tensor, scale_u: torch.Tensor, scale_v: torch.Tensor) -> torch.Tensor:
    """packed (N,C/8), U (N,K), V (K,C) -> (N,C) fp16.  compress_quantize.py:154-225."""
    assert packed_tensor.dtype == torch.uint8, "Packed tensor must be UINT8"
    assert scale_u.dtype == torch.half and scale_v.dtype == torch.half
    assert packed_tensor.ndim == 2 and scale_u.ndim == 2 and scale_v.ndim == 2
    assert scale_u.shape[1] == scale_v.shape[0], "Rank K mismatch"
    n, c8 = packed_tensor.shape
    c, k = c8 * 8, scale_u.shape[1]
    assert scale_u.shape[0] == n and scale_v.shape[1] == c
    packed_tensor, scale_u = packed_tensor.contiguous(), scale_u.contiguous()
    v_ck = scale_v.contiguous() if k == 1 else scale_v.t().contiguous()  # kernel wants (C,K)
    out = torch.empty((n, c), dtype=torch.half, device=packed_tensor.device)
    rc = nv.lib().cf_binary_decompress(nv.ptr(packed_tensor), nv.ptr(scale_u), nv.ptr(v_ck), k, None, nv.ptr(out),
                                       n, c, nv.stream_ptr())
    nv.check(rc, "cf_binary_decompress")
    return out


def sim_binary(input_tensor: torch.Tensor, rank: int | None = None) -> torch.Tensor:
    """Quantise-dequantise in one call.  compress_quantize.py:300-335."""
    assert rank is not None, "Rank must be provided"
    assert rank >= 1 or rank == -1, "Rank must be >= 1 or -1"
    x = _check2d(input_tensor)
    if rank == -1:
        # base = 0 and "new_base" = 0 + (+-scale): the bare dequantised tensor
        return _sign_compress(nv.CODEC_BINARY, x, None, True)[3]
    packed, u, v = quantize_1bit(x, rank)
    return dequantize_1bit(packed, u, v)


# ------------------------------------------------------------------------------------ INT2
def quantize_int2(input_tensor: torch.Tensor):
    """-> packed (N,C/4) u8, chan_scale (1,C), tok_scale (N,1).  compress_quantize.py:643-704."""
    x = _check2d(input_tensor)
    assert x.shape[1] % 4 == 0
    packed, tok, chan, _ = _sign_compress(nv.CODEC_INT2, x, None, False)
    return packed, chan.view(1, -1), tok


def dequantize_int2(packed_indices: torch.Tensor, chan_scale: torch.Tensor, tok_scale: torch.Tensor) -> torch.Tensor:
    """compress_quantize.py:707-753."""
    assert packed_indices.dtype == torch.uint8, "Packed tensor must be UINT8"
    assert chan_scale.dtype == torch.half and tok_scale.dtype == torch.half
    assert chan_scale.dim() == 2 and chan_scale.shape[0] == 1, "Chan scale shape error"
    assert tok_scale.dim() == 2 and tok_scale.shape[1] == 1, "Tok scale shape error"
    n, c = tok_scale.shape[0], chan_scale.shape[1]
    assert packed_indices.shape == (n, c // 4)
    out = torch.empty((n, c), dtype=torch.half, device=packed_indices.device)
    rc = nv.lib().cf_int2_decompress(nv.ptr(packed_indices.contiguous()), nv.ptr(tok_scale.contiguous()),
                                     nv.ptr(chan_scale.contiguous()), None, nv.ptr(out), n, c, nv.stream_ptr())
    nv.check(rc, "cf_int2_decompress")
    return out


def sim_int2(input_tensor: torch.Tensor) -> torch.Tensor:
    """compress_quantize.py:339-384."""
    x = _check2d(input_tensor)
    return _sign_compress(nv.CODEC_INT2, x, None, True)[3]


def sim_int2_minmax(input_tensor: torch.Tensor) -> torch.Tensor:
    """Channel-wise 4-level min/max quantise-dequantise (simulation only).  compress_quantize.py:386-426.
    As for sim_int4, a constant column (zero scale) reconstructs to its value instead of NaN."""
    x = _check2d(input_tensor)
    n = x.shape[0]
    if n % 2:  # the nibble packing needs an even N; pad with a copy of the last row (min/max unchanged)
        xp = torch.cat([x, x[-1:]], dim=0).contiguous()
        return _minmax_compress(nv.CODEC_INT4, xp, None, want_recon=True, levels=3)[3][:n].contiguous()
    return _minmax_compress(nv.CODEC_INT4, x, None, want_recon=True, levels=3)[3]


# ------------------------------------------------------------------------------------ INT4 / INT8
def _minmax_compress(codec, x, base, want_codes=True, want_recon=False, levels=15):
    nv.require_cuda_half(x, "x")
    n, c = x.shape
    dev = x.device
    if codec == nv.CODEC_INT4:
        assert n % 2 == 0, f"Dimension N (0) size must be even for INT4 packing, got {n}"
        codes = torch.empty((n // 2, c), dtype=torch.uint8, device=dev)
        second = torch.empty((1, c), dtype=torch.half, device=dev)
        fn, name = nv.lib().cf_int4_compress, "cf_int4_compress"
        if levels == 3:  # the 4-level simulation-only variant (sim_int2_minmax)
            fn, name = nv.lib().cf_int2mm_compress, "cf_int2mm_compress"
    else:
        codes = torch.empty((n, c), dtype=torch.int8, device=dev)
        second = torch.empty((1, c), dtype=torch.int16, device=dev)
        fn, name = nv.lib().cf_int8_compress, "cf_int8_compress"
    scale = torch.empty((1, c), dtype=torch.half, device=dev)
    recon = torch.empty_like(x) if want_recon else None
    ws_bytes = nv.workspace_bytes(codec, n, c)
    ws = nv.workspace(ws_bytes, dev)
    rc = fn(nv.ptr(x), nv.ptr(base), nv.ptr(recon), nv.ptr(codes), nv.ptr(scale), nv.ptr(second), n, c, nv.ptr(ws),
            ws.numel(), nv.stream_ptr())
    nv.check(rc, name)
    return codes, scale, second, recon


def quantize_int4(input_tensor: torch.Tensor):
    """-> packed (N/2,C) u8, scale (1,C), min (1,C).  compress_quantize.py:522-583."""
    with Profiler.scope("compact.quantize_int4"):
        x = _check2d(input_tensor)
        codes, scale, mn, _ = _minmax_compress(nv.CODEC_INT4, x, None)
        return codes, scale, mn


def dequantize_int4(packed_tensor: torch.Tensor, scale: torch.Tensor, min_val: torch.Tensor) -> torch.Tensor:
    """compress_quantize.py:585-640."""
    with Profiler.scope("compact.dequantize_int4"):
        assert packed_tensor.dtype == torch.uint8, "Packed tensor must be UINT8"
        assert scale.dtype == torch.half and min_val.dtype == torch.half
        n2, c = packed_tensor.shape
        assert scale.shape == (1, c) and min_val.shape == (1, c)
        out = torch.empty((n2 * 2, c), dtype=torch.half, device=packed_tensor.device)
        rc = nv.lib().cf_int4_decompress(nv.ptr(packed_tensor.contiguous()), nv.ptr(scale.contiguous()),
                                         nv.ptr(min_val.contiguous()), None, nv.ptr(out), n2 * 2, c, nv.stream_ptr())
        nv.check(rc, "cf_int4_decompress")
        return out


def sim_int4(input_tensor: torch.Tensor, dim) -> torch.Tensor:
    """Quantise-dequantise with min/max taken along `dim`.  compress_quantize.py:487-520.
    Differs from the reference only for constant columns (zero scale): the reference's
    simulation yields NaN there, this yields the column value (codes NaN -> 0)."""
    assert input_tensor.dim() == 2
    if dim in (1, -1):
        return sim_int4(input_tensor.t().contiguous(), 0).t().contiguous()
    x = _check2d(input_tensor)
    n = x.shape[0]
    if n % 2:  # the packed codec needs an even N; pad with a copy of the last row (min/max unchanged)
        xp = torch.cat([x, x[-1:]], dim=0).contiguous()
        return _minmax_compress(nv.CODEC_INT4, xp, None, want_recon=True)[3][:n].contiguous()
    return _minmax_compress(nv.CODEC_INT4, x, None, want_recon=True)[3]


def quantize_int8(input_tensor: torch.Tensor):
    """-> q (N,C) int8, scale (1,C) fp16, zero_point (1,C) int16.  compress_quantize.py:428-471."""
    x = _check2d(input_tensor)
    codes, scale, zp, _ = _minmax_compress(nv.CODEC_INT8, x, None)
    return codes, scale, zp


def dequantize_int8(q_tensor, scale, zero_point):
    """compress_quantize.py:473-484."""
    assert q_tensor.dtype == torch.int8 and scale.dtype == torch.half and zero_point.dtype == torch.int16
    n, c = q_tensor.shape
    out = torch.empty((n, c), dtype=torch.half, device=q_tensor.device)
    rc = nv.lib().cf_int8_decompress(nv.ptr(q_tensor.contiguous()), nv.ptr(scale.contiguous()),
                                     nv.ptr(zero_point.contiguous()), None, nv.ptr(out), n, c, nv.stream_ptr())
    nv.check(rc, "cf_int8_decompress")
    return out
