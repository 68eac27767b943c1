"""1:m block-argmax sparsifier ("top-k"; mirror of xfuser/compact/compress_topk.py)."""
from __future__ import annotations

import torch

from . import _native as nv

SPARSE_LAST_DIM_SIZE = 1024
VALID_COMPRESS_LEVELS = [2, 4, 8, 16]


def _topk_compress(x, base, m, want_new_base=False):
    nv.require_cuda_half(x, "input_tensor")
    x = x.contiguous()
    numel = x.numel()
    assert m in VALID_COMPRESS_LEVELS, f"sparse ratio {m} not in {VALID_COMPRESS_LEVELS}"
    assert numel % SPARSE_LAST_DIM_SIZE == 0
    val = torch.empty(numel // m, dtype=torch.half, device=x.device)
    idx = torch.empty(numel // (2 * m), dtype=torch.uint8, device=x.device)
    nb = torch.empty_like(x) if want_new_base else None
    rc = nv.lib().cf_topk_compress(nv.ptr(x), nv.ptr(base), nv.ptr(nb), nv.ptr(val), nv.ptr(idx), numel, m,
                                   nv.stream_ptr())
    nv.check(rc, "cf_topk_compress")
    return val, idx, nb


def topk_compress(input_tensor: torch.Tensor, m: int):
    """(A, 2mB) -> val (A, 2B) fp16, idx (A, B) u8 (high nibble = first block).  compress_topk.py:11-41."""
    a, w = input_tensor.shape
    assert w % (2 * m) == 0, "The number of columns must be divisible by 2*m."
    val, idx, _ = _topk_compress(input_tensor, None, m)
    return val.view(a, w // m), idx.view(a, w // (2 * m))


def topk_decompress(compressed_val_tensor: torch.Tensor, compressed_idx_tensor: torch.Tensor, m: int,
                    base: torch.Tensor | None = None):
    """compress_topk.py:108-126 (+ optional fused residual add)."""
    a, b = compressed_idx_tensor.shape
    numel = a * b * 2 * m
    out = torch.empty((a, 2 * m * b), dtype=torch.half, device=compressed_val_tensor.device)
    rc = nv.lib().cf_topk_decompress(nv.ptr(compressed_val_tensor.contiguous()),
                                     nv.ptr(compressed_idx_tensor.contiguous()), nv.ptr(base), nv.ptr(out), numel, m,
                                     nv.stream_ptr())
    nv.check(rc, "cf_topk_decompress")
    return out


def topk_sparsify(input_tensor: torch.Tensor, m: int):
    """Keep the arg-max-|x| element of every m-block, zero the rest.  compress_topk.py:165-194."""
    assert input_tensor.is_cuda, "Input tensor must be on CUDA."
    assert input_tensor.dim() == 2, "Input tensor must be 2D."
    assert input_tensor.shape[1] % m == 0
    return _topk_compress(input_tensor, None, m, want_new_base=True)[2]


def sim_topk(x: torch.Tensor, m: int):
    """compress_topk.py:221-236 (ties: lowest index)."""
    assert x.shape[-1] % m == 0
    return _topk_compress(x, None, m, want_new_base=True)[2].view(x.shape)
