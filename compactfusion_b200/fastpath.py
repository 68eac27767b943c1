"""Fused residual codecs ("fastpath"; mirror of xfuser/compact/fastpath.py).

`binary_quant_fastpath` / `int2_quant_fastpath` compute delta = x - base, the rank-1
token x channel scales, the packed codes and (optionally) the error-feedback base in fused
sm_100a kernels; `*_dequant_fastpath` reconstruct base + dequant(codes).  Signatures, return
tuples and tensor layouts are the reference's (fastpath.py:124, :371, :584, :745).  The
`sim_*` twins are built from the stand-alone codecs exactly like the reference's
(fastpath.py:233, :441, :814, :853).
"""
from __future__ import annotations

import torch

from . import _native as nv
from .compress_quantize import (_sign_compress, dequantize_1bit, dequantize_int2, quantize_1bit, quantize_int2)
from .prof import Profiler


def _check_pair(x, base):
    nv.require_cuda_half(x, "x_tensor_nc")
    nv.require_cuda_half(base, "base_tensor_nc")
    assert x.ndim == 2 and base.ndim == 2
    assert x.shape == base.shape
    return x.contiguous(), base.contiguous()


@Profiler.prof_func("compact.binary_quant_fastpath")
def binary_quant_fastpath(x_tensor_nc: torch.Tensor, base_tensor_nc: torch.Tensor, rank: int, update_cache: bool):
    """-> packed (N,C/8) u8, scale_u (N,K), scale_v (C,K), new_base (N,C) | None."""
    assert rank >= 1 or rank == -1, "Rank must be >= 1 or -1"
    x, base = _check_pair(x_tensor_nc, base_tensor_nc)
    assert x.shape[1] % 8 == 0, "C_COLS must be divisible by 8 for packing output alignment"
    if rank == -1:
        return _sign_compress(nv.CODEC_BINARY, x, base, update_cache)
    # rank-K scales (deprecated in the reference, main.py:188-189): subspace iteration on |delta|
    from .compress_lowrank import subspace_iter
    delta_abs = torch.abs(x - base)
    with Profiler.scope(f"compact.quant.scale_rank{rank}_approx"):
        su, svt, _ = subspace_iter(delta_abs, rank=rank, num_iters=2)
    u = su.contiguous().half()
    v = svt.t().contiguous().half()
    packed, _, _, _ = _sign_compress(nv.CODEC_BINARY, x, base, False)
    new_base = binary_dequant_fastpath(packed, u, v, base) if update_cache else None
    return packed, u, v, new_base


@Profiler.prof_func("compact.binary_dequant_fastpath")
def binary_dequant_fastpath(packed: torch.Tensor, scale_u_nk: torch.Tensor, scale_v_ck: torch.Tensor,
                            base_nc: torch.Tensor, out: torch.Tensor | None = None):
    """recon (N,C) = base + (2 bit - 1) * fp16(U V^T).  `out` may alias base_nc (in-place)."""
    assert packed.dtype == torch.uint8
    assert scale_u_nk.dtype == torch.half and scale_v_ck.dtype == torch.half
    nv.require_cuda_half(base_nc, "base_nc")
    assert packed.ndim == 2 and scale_u_nk.ndim == 2 and scale_v_ck.ndim == 2 and base_nc.ndim == 2
    n, c8 = packed.shape
    c, k = c8 * 8, scale_u_nk.shape[1]
    assert k >= 1 and scale_v_ck.shape == (c, k), "Scale V shape mismatch"
    assert base_nc.shape == (n, c) and scale_u_nk.shape == (n, k)
    packed, scale_u_nk, scale_v_ck, base_nc = (packed.contiguous(), scale_u_nk.contiguous(),
                                               scale_v_ck.contiguous(), base_nc.contiguous())
    if out is None:
        out = torch.empty_like(base_nc)
    rc = nv.lib().cf_binary_decompress(nv.ptr(packed), nv.ptr(scale_u_nk), nv.ptr(scale_v_ck), k, nv.ptr(base_nc),
                                       nv.ptr(out), n, c, nv.stream_ptr())
    nv.check(rc, "cf_binary_decompress")
    return out


@Profiler.prof_func("compact.int2_quant_fastpath")
def int2_quant_fastpath(x_tensor_nc: torch.Tensor, base_tensor_nc: torch.Tensor, update_cache: bool, rank: int = -1):
    """-> packed (N,C/4) u8, scale_u = tok (N,1), scale_v = chan (C,1), new_base | None."""
    assert rank == -1, "INT2 fastpath only supports channel/token scales (rank=-1 equivalent)"
    x, base = _check_pair(x_tensor_nc, base_tensor_nc)
    assert x.shape[1] % 4 == 0, "C_COLS must be divisible by 4 for packing output alignment"
    return _sign_compress(nv.CODEC_INT2, x, base, update_cache)


@Profiler.prof_func("compact.int2_dequant_fastpath")
def int2_dequant_fastpath(packed: torch.Tensor, scale_u_nk: torch.Tensor, scale_v_ck: torch.Tensor,
                          base_nc: torch.Tensor, out: torch.Tensor | None = None):
    assert packed.dtype == torch.uint8
    assert scale_u_nk.dtype == torch.half and scale_v_ck.dtype == torch.half
    nv.require_cuda_half(base_nc, "base_nc")
    n, c4 = packed.shape
    c = c4 * 4
    assert scale_u_nk.shape == (n, 1), f"INT2 expects scale_u (N,1), got {tuple(scale_u_nk.shape)}"
    assert scale_v_ck.shape == (c, 1), f"INT2 expects scale_v (C,1), got {tuple(scale_v_ck.shape)}"
    assert base_nc.shape == (n, c)
    packed, scale_u_nk, scale_v_ck, base_nc = (packed.contiguous(), scale_u_nk.contiguous(),
                                               scale_v_ck.contiguous(), base_nc.contiguous())
    if out is None:
        out = torch.empty_like(base_nc)
    rc = nv.lib().cf_int2_decompress(nv.ptr(packed), nv.ptr(scale_u_nk), nv.ptr(scale_v_ck), nv.ptr(base_nc),
                                     nv.ptr(out), n, c, nv.stream_ptr())
    nv.check(rc, "cf_int2_decompress")
    return out


# ---- simulation twins: same composition as the reference's (slowpath codec + base add) ------
def sim_binary_quant_fastpath(x_tensor_nc, base_tensor_nc, rank: int, update_cache: bool):
    assert rank >= 1 or rank == -1
    delta = x_tensor_nc - base_tensor_nc
    packed, u_nk, v_kc = quantize_1bit(delta, rank=rank)
    new_base = base_tensor_nc + dequantize_1bit(packed, u_nk, v_kc) if update_cache else None
    return packed, u_nk, v_kc.transpose(0, 1).contiguous(), new_base


def sim_binary_dequant_fastpath(packed_in_nc8, scale_u_nk, scale_v_ck, base_nc):
    return base_nc + dequantize_1bit(packed_in_nc8, scale_u_nk, scale_v_ck.transpose(0, 1).contiguous())


def sim_int2_quant_fastpath(x_tensor_nc, base_tensor_nc, update_cache: bool, rank: int = -1):
    assert rank == -1
    delta = x_tensor_nc - base_tensor_nc
    packed, chan_1c, tok_n1 = quantize_int2(delta)
    new_base = base_tensor_nc + dequantize_int2(packed, chan_1c, tok_n1) if update_cache else None
    return packed, tok_n1, chan_1c.t().contiguous(), new_base


def sim_int2_dequant_fastpath(packed_in_nc4, scale_u_nk, scale_v_ck, base_nc):
    return base_nc + dequantize_int2(packed_in_nc4, scale_v_ck.t().contiguous(), scale_u_nk)
