"""Pure compress / decompress functions and their payload (de)serialisation
(mirror of xfuser/compact/slowpath.py).  Payload = one flat fp16 tensor, layouts of
SURVEY.md App-A; INT4 and INT2 are wire codecs here (the reference only simulates them,
slowpath.py:80-81 vs :201-206).
"""
from __future__ import annotations

import torch

from .compress_lowrank import lowrank_q_pack, lowrank_q_reconstruct, lowrank_reconstruct, subspace_iter
from .compress_quantize import (dequantize_1bit, dequantize_int2, dequantize_int4, quantize_1bit, quantize_int2,
                                quantize_int4, sim_binary, sim_int2, sim_int2_minmax, sim_int4)
from .compress_topk import SPARSE_LAST_DIM_SIZE, sim_topk, topk_compress, topk_decompress
from .utils import COMPACT_COMPRESS_TYPE

T = COMPACT_COMPRESS_TYPE


def _flat_half(t: torch.Tensor) -> torch.Tensor:
    t = t.contiguous()
    return (t.view(torch.half) if t.dtype != torch.half else t).reshape(-1)


def slowpath_compress(x: torch.Tensor, compress_type: COMPACT_COMPRESS_TYPE, rank: int = None, sparse_ratio: int = None):
    """(N,C) fp16 -> flat fp16 payload.  slowpath.py:26-84."""
    assert x.dtype == torch.half, f"x.dtype: {x.dtype}"
    assert x.dim() == 2
    n, c = x.shape
    if compress_type == T.BINARY:
        assert rank is not None and (rank >= 1 or rank == -1), "Rank must be >= 1 or -1 for BINARY compression"
        q, su, sv = quantize_1bit(x, rank=rank)
        parts = [q, su, sv]
    elif compress_type == T.LOW_RANK:
        assert rank is not None and rank >= 1, "Rank must be provided for LOW_RANK compression"
        u, v, _ = subspace_iter(x, rank, 2)
        parts = [u, v]
    elif compress_type == T.LOW_RANK_Q:
        assert rank is not None and rank >= 1, "Rank must be provided for LOW_RANK_Q compression"
        u, v, _ = subspace_iter(x, rank, 2)
        if n % 2 == 0 and c % 2 == 0 and rank <= 64 and (n * rank) % 4 == 0 and (c * rank) % 4 == 0:
            return lowrank_q_pack(u, v)   # both int4 encodes and the concatenation in two launches
        qu, su, mu = quantize_int4(u)
        qv, sv, mv = quantize_int4(v.t().contiguous())
        parts = [qu, su, mu, qv, sv, mv]
    elif compress_type == T.SPARSE:
        assert sparse_ratio is not None, "sparse_ratio must be provided for SPARSE compression"
        val, idx = topk_compress(x.view(-1, SPARSE_LAST_DIM_SIZE), sparse_ratio)
        parts = [val, idx]
    elif compress_type == T.INT4:
        parts = list(quantize_int4(x))
    elif compress_type == T.INT2:
        q, chan, tok = quantize_int2(x)
        parts = [q, tok, chan]
    else:
        raise ValueError(f"Invalid compress_type value: {compress_type}")
    return torch.cat([_flat_half(p) for p in parts], dim=0)


def slowpath_decompress(x: torch.Tensor, shape: tuple, compress_type: COMPACT_COMPRESS_TYPE, rank: int = None,
                        sparse_ratio: int = None):
    """flat fp16 payload -> (N,C) fp16.  slowpath.py:86-175."""
    assert x.dim() == 1 and x.dtype == torch.half
    assert len(shape) == 2
    n, c = shape
    numel = n * c

    def u8(t, rows, cols):
        return t.contiguous().view(torch.uint8).view(rows, cols)

    if compress_type == T.BINARY:
        assert rank is not None and (rank >= 1 or rank == -1)
        k = 1 if rank == -1 else rank
        sizes = [numel // 16, n * k, k * c]
        assert sum(sizes) == x.numel(), f"Binary split error: {sum(sizes)} != {x.numel()}"
        q, su, sv = torch.split(x, sizes)
        return dequantize_1bit(u8(q, n, c // 8), su.view(n, k), sv.view(k, c))
    if compress_type == T.LOW_RANK:
        assert rank is not None and rank >= 1
        u, v = torch.split(x, [n * rank, rank * c])
        return lowrank_reconstruct(u.view(n, rank), v.view(rank, c))
    if compress_type == T.LOW_RANK_Q:
        assert rank is not None and rank >= 1
        assert (n * rank) % 4 == 0 and (c * rank) % 4 == 0
        if n % 2 == 0 and c % 8 == 0 and rank <= 64 and x.is_contiguous():
            return lowrank_q_reconstruct(x, n, c, rank)   # decode + transpose + product in one launch
        qu, su, mu, qv, sv, mv = torch.split(x, [n * rank // 4, rank, rank, c * rank // 4, rank, rank])
        u = dequantize_int4(u8(qu, n // 2, rank), su.view(1, rank), mu.view(1, rank))
        v = dequantize_int4(u8(qv, c // 2, rank), sv.view(1, rank), mv.view(1, rank))
        return lowrank_reconstruct(u, v.t().contiguous())
    if compress_type == T.SPARSE:
        val, idx = torch.split(x, [numel // sparse_ratio, numel // sparse_ratio // 4])
        a = numel // SPARSE_LAST_DIM_SIZE
        return topk_decompress(val.view(a, -1), u8(idx, a, SPARSE_LAST_DIM_SIZE // sparse_ratio // 2),
                               sparse_ratio).view(shape)
    if compress_type == T.INT4:
        q, s, mn = torch.split(x, [numel // 4, c, c])
        return dequantize_int4(u8(q, n // 2, c), s.view(1, c), mn.view(1, c))
    if compress_type == T.INT2:
        q, tok, chan = torch.split(x, [numel // 8, n, c])
        return dequantize_int2(u8(q, n, c // 4), chan.view(1, c), tok.view(n, 1))
    raise ValueError(f"Invalid compress_type value: {compress_type}")


def sim_compress(x: torch.Tensor, compress_type: COMPACT_COMPRESS_TYPE, sparse_ratio: int = None, rank: int = None):
    """compress o decompress without size reduction.  slowpath.py:185-239."""
    if compress_type == T.IDENTITY:
        return x
    if compress_type == T.SPARSE:
        assert sparse_ratio is not None
        return sim_topk(x, sparse_ratio)
    if compress_type == T.BINARY:
        assert rank is not None
        return sim_binary(x.half(), rank=rank).half()
    if compress_type == T.INT2:
        return sim_int2(x)
    if compress_type == T.INT4:
        return sim_int4(x, dim=0)
    if compress_type == T.LOW_RANK:
        assert rank is not None
        u, v, _ = subspace_iter(x, rank, 2)
        return lowrank_reconstruct(u, v)
    if compress_type == T.LOW_RANK_Q:
        assert rank is not None
        u, v, _ = subspace_iter(x, rank, 2)
        return lowrank_reconstruct(sim_int4(u, dim=0), sim_int4(v, dim=1))
    if compress_type == T.INT2_MINMAX:
        return sim_int2_minmax(x)
    if compress_type == T.LOW_RANK_AWL:
        raise ValueError(f"{compress_type} is a deprecated experiment of the reference and is not provided")
    raise ValueError("Invalid compress_type value")
