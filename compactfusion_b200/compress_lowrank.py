"""Low-rank projector (mirror of xfuser/compact/compress_lowrank.py)."""
from __future__ import annotations

import os

import torch

from . import _native as nv
from .prof import Profiler

MAX_RANK = 64


def svd(input_tensor: torch.Tensor, rank: int):
    """Exact truncated SVD reference (library call, test helper).  compress_lowrank.py:5-12."""
    u, s, vh = torch.linalg.svd(input_tensor.float(), full_matrices=False)
    return (u[:, :rank] @ torch.diag(s[:rank])).to(input_tensor.dtype), vh[:rank, :].to(input_tensor.dtype)


def _init_q(n: int, rank: int, device, orthonormalise: bool) -> torch.Tensor:
    """The reference's random start, randn(n, rank) in fp32 from the global torch RNG (same draw, same RNG
    consumption, compress_lowrank.py:40-42).  The reference then orthonormalises it with a library QR; that
    step only changes the BASIS of span(Q0): Z = A^T A Q0 spans the same subspace for Q0 and for Q0 R^-1, and
    every later step re-orthonormalises, so U V (the only thing the codec ships or anyone compares) is the
    same.  The library QR of a (C, r) matrix costs ~1 ms on the GPU -- twice the whole projector -- so it is
    only run when no iteration follows (num_iters == 0 returns Q0 itself) or with CF_LR_ORTHO_INIT=1."""
    q = torch.randn(n, rank, device=device, dtype=torch.float)
    if orthonormalise or os.environ.get("CF_LR_ORTHO_INIT", "0") == "1":
        q, _ = torch.linalg.qr(q)
    return q.contiguous()


def lowrank_project(x: torch.Tensor, base: torch.Tensor | None, rank: int, num_iters: int,
                    init_q: torch.Tensor | None = None, u_out=None, v_out=None, want_q=False):
    """U (N,r), V (r,C) fp16 with A = x - base ~= U V, fused residual subtract."""
    nv.require_cuda_half(x, "A")
    assert x.dim() == 2
    assert 1 <= rank <= MAX_RANK, f"rank must be in [1, {MAX_RANK}]"
    x = x.contiguous()
    n, c = x.shape
    q0 = _init_q(c, rank, x.device, num_iters == 0) if init_q is None else init_q.float().contiguous()
    assert q0.shape == (c, rank)
    u = torch.empty((n, rank), dtype=torch.half, device=x.device) if u_out is None else u_out
    v = torch.empty((rank, c), dtype=torch.half, device=x.device) if v_out is None else v_out
    q = torch.empty((c, rank), dtype=torch.float, device=x.device) if want_q else None
    ws_bytes = nv.workspace_bytes(nv.CODEC_LOWRANK, n, c, rank)
    ws = nv.workspace(ws_bytes, x.device)
    rc = nv.lib().cf_lowrank_project(nv.ptr(x), nv.ptr(base), nv.ptr(q0), nv.ptr(u), nv.ptr(v), nv.ptr(q), n, c, rank,
                                     num_iters, nv.ptr(ws), ws.numel(), nv.stream_ptr())
    nv.check(rc, "cf_lowrank_project")
    return u, v, q


@Profiler.prof_func("compact.subspace_iter")
def subspace_iter(A: torch.Tensor, rank: int, num_iters: int = 10, init_q: torch.Tensor | None = None):
    """A (m,n) ~= U (m,rank) @ V (rank,n); returns U, V, Q in A.dtype.  compress_lowrank.py:16-62.
    fp16 input only (the reference's callers pass fp16 activations / residuals)."""
    dtype = A.dtype
    a16 = A if A.dtype == torch.half else A.half()
    u, v, q = lowrank_project(a16, None, rank, num_iters, init_q=init_q, want_q=True)
    return u.to(dtype), v.to(dtype), q.to(dtype)


def lowrank_reconstruct(u: torch.Tensor, v: torch.Tensor, base: torch.Tensor | None = None,
                        out: torch.Tensor | None = None) -> torch.Tensor:
    """base + fp16(U V), fused (replaces torch.matmul(u, v) + add, slowpath.py:152-154)."""
    n, r = u.shape
    c = v.shape[1]
    assert v.shape[0] == r
    if out is None:
        out = torch.empty((n, c), dtype=torch.half, device=u.device)
    v = v.contiguous()
    if v.data_ptr() % 16:  # a view into a wire payload may be only 2-byte aligned
        v = v.clone()
    rc = nv.lib().cf_lowrank_reconstruct(nv.ptr(u.contiguous()), nv.ptr(v), nv.ptr(base), nv.ptr(out), n, c,
                                         r, nv.stream_ptr())
    nv.check(rc, "cf_lowrank_reconstruct")
    return out


def lowrank_q_reconstruct(payload: torch.Tensor, n: int, c: int, rank: int, base: torch.Tensor | None = None,
                          out: torch.Tensor | None = None) -> torch.Tensor:
    """base + fp16(deq(qU) deq(qV^T)^T) straight from a LOW_RANK_Q wire payload (slowpath.py:69-75): the int4
    decode of both factors, the transpose and the product of slowpath_decompress (slowpath.py:156-164) plus the
    residual add, in one launch."""
    want = (n * rank + c * rank) // 4 + 4 * rank
    assert payload.dtype == torch.half and payload.is_contiguous() and payload.numel() >= want, "not a LOW_RANK_Q payload"
    if out is None:
        out = torch.empty((n, c), dtype=torch.half, device=payload.device)
    rc = nv.lib().cf_lowrank_q_reconstruct(nv.ptr(payload), nv.ptr(base), nv.ptr(out), n, c, rank, nv.stream_ptr())
    nv.check(rc, "cf_lowrank_q_reconstruct")
    return out


def lowrank_q_pack(u: torch.Tensor, v: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """U (N, r), V (r, C) fp16 -> the LOW_RANK_Q payload [qU, sU, mU, qV^T, sV, mV] (slowpath.py:62-75) as flat fp16:
    quantize_int4 of U and of V^T (never formed) and the concatenation, in two launches."""
    n, r = u.shape
    c = v.shape[1]
    assert v.shape[0] == r and u.dtype == torch.half and v.dtype == torch.half
    numel = (n * r + c * r) // 4 + 4 * r
    if out is None:
        out = torch.empty(numel, dtype=torch.half, device=u.device)
    assert out.dtype == torch.half and out.is_contiguous() and out.numel() >= numel
    rc = nv.lib().cf_lowrank_q_pack(nv.ptr(u.contiguous()), nv.ptr(v.contiguous()), nv.ptr(out), n, c, r, nv.stream_ptr())
    nv.check(rc, "cf_lowrank_q_pack")
    return out[:numel]
