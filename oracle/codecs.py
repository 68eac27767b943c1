"""CPU restatement of the reference codecs (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Every function cites the reference file:line it follows (paths relative to
/root/reference).  Arithmetic is done with torch CPU fp16 tensors because the
reference *is* eager torch arithmetic: each fp16 op is an fp32 op followed by
one rounding to fp16, which is what the Triton kernels and the CUDA kernels
under test also produce (double rounding fp32->fp16 is innocuous for + - * /).
Bit packing is plain numpy integer work.

All tensors are (N, C) row-major fp16 on CPU unless stated otherwise.
"""
from __future__ import annotations

import numpy as np
import torch

SPARSE_LAST_DIM_SIZE = 1024  # xfuser/compact/compress_topk.py:8


def _h(x: torch.Tensor) -> torch.Tensor:
    assert x.dtype == torch.half and x.device.type == "cpu"
    return x.contiguous()


# ----------------------------------------------------------------------------
# bit packing helpers
# ----------------------------------------------------------------------------
def pack_bits_lsb_first(bits_nc: np.ndarray) -> np.ndarray:
    """(N, C) of 0/1 -> (N, C/8) uint8; element c is bit c%8 of byte c//8.
    xfuser/compact/fastpath.py:62-72, compress_quantize.py:123-145."""
    n, c = bits_nc.shape
    assert c % 8 == 0
    return np.packbits(bits_nc.astype(np.uint8), axis=1, bitorder="little")


def unpack_bits_lsb_first(packed: np.ndarray, c: int) -> np.ndarray:
    """fastpath.py:339-352."""
    return np.unpackbits(packed, axis=1, bitorder="little")[:, :c]


def pack_int2(idx_nc: np.ndarray) -> np.ndarray:
    """(N, C) of 0..3 -> (N, C/4) uint8; element c at bits 2(c%4)..2(c%4)+1.
    fastpath.py:546-549, compress_quantize.py:690-695."""
    n, c = idx_nc.shape
    assert c % 4 == 0
    v = idx_nc.astype(np.uint8).reshape(n, c // 4, 4)
    return (v[..., 0] | (v[..., 1] << 2) | (v[..., 2] << 4) | (v[..., 3] << 6)).astype(np.uint8)


def unpack_int2(packed: np.ndarray) -> np.ndarray:
    """compress_quantize.py:739-743, fastpath.py:717-724."""
    n, c4 = packed.shape
    out = np.empty((n, c4, 4), dtype=np.uint8)
    for j in range(4):
        out[..., j] = (packed >> (2 * j)) & 3
    return out.reshape(n, c4 * 4)


# ----------------------------------------------------------------------------
# BINARY (1-bit sign + rank-1 token x channel mean-|delta| scale)
# ----------------------------------------------------------------------------
def binary_scales(delta: torch.Tensor):
    """U (N,1), V (C,1) for rank=-1.  fastpath.py:154-166 == compress_quantize.py:35-48."""
    a = torch.abs(_h(delta))
    v = torch.mean(a, dim=0)  # (C,) fp16
    u = torch.mean(a, dim=1, keepdim=True)  # (N,1) fp16
    u = u / u.mean(dim=0, keepdim=True)
    return u.contiguous(), v.unsqueeze(1).contiguous()


def binary_sign_bits(delta: torch.Tensor) -> np.ndarray:
    """1 <=> delta >= 0 (zero is positive, NaN -> 0).  fastpath.py:62."""
    return (_h(delta) >= 0).numpy().astype(np.uint8)


def scale_matrix(u_nk: torch.Tensor, v_ck: torch.Tensor) -> torch.Tensor:
    """fp16 scale[n,c] = sum_k U[n,k] V[c,k].  K=1: one fp16 product
    (fastpath.py:109).  K>1: the reference sums fp16 products in an
    implementation-defined tree order (Triton) or by GEMM (dequantize_1bit,
    compress_quantize.py:196); we define fp32 accumulation in k order and one
    rounding, and compare with tolerance there."""
    k = u_nk.shape[1]
    if k == 1:
        return (u_nk * v_ck.t()).to(torch.half)
    return (u_nk.float() @ v_ck.float().t()).to(torch.half)


def binary_quant(x: torch.Tensor, base: torch.Tensor, update_cache: bool, scales=None):
    """== binary_quant_fastpath(rank=-1) (fastpath.py:124-228).
    Returns packed (N,C/8) uint8 ndarray, U (N,1), V (C,1), new_base|None.
    `scales=(U,V)` injects externally computed scales (used to test the
    elementwise stage bit-exactly given identical scale tensors)."""
    x, base = _h(x), _h(base)
    delta = x - base  # fastpath.py:151 / :58
    u, v = binary_scales(delta) if scales is None else scales
    bits = binary_sign_bits(delta)
    packed = pack_bits_lsb_first(bits)
    new_base = None
    if update_cache:
        new_base = binary_dequant(packed, u, v, base)  # same expression, fastpath.py:109-116 vs :328-363
    return packed, u, v, new_base


def binary_dequant(packed: np.ndarray, u_nk: torch.Tensor, v_ck: torch.Tensor, base: torch.Tensor | None):
    """recon = base + (2*bit-1) * fp16(U V^T).  fastpath.py:277-367.
    base=None gives the bare dequantised delta (dequantize_1bit, compress_quantize.py:154-225)."""
    c = v_ck.shape[0]
    bits = torch.from_numpy(unpack_bits_lsb_first(packed, c).astype(np.int8))
    scale = scale_matrix(_h(u_nk), _h(v_ck))
    sign = (2 * bits - 1).to(torch.half)
    recv = sign * scale
    if base is None:
        return recv
    return _h(base) + recv


def sim_binary(delta: torch.Tensor) -> torch.Tensor:
    """compress_quantize.py:300-335 with rank=-1."""
    delta = _h(delta)
    a = torch.abs(delta)
    chan = torch.mean(a, dim=0, keepdim=True)
    tok = torch.mean(a, dim=1, keepdim=True)
    tok = tok / tok.mean()
    scale = chan * tok
    q = torch.sign(delta)
    q = torch.where(q == 0, torch.ones_like(q), q)
    return q * scale


# ----------------------------------------------------------------------------
# INT2 (sign + magnitude bit, levels +-0.5 thr / +-2 thr)
# ----------------------------------------------------------------------------
def int2_scales(delta: torch.Tensor):
    """tok (N,1), chan (C,1).  fastpath.py:614-625 == compress_quantize.py:671-683."""
    a = torch.abs(_h(delta))
    chan = torch.mean(a, dim=0, keepdim=True)  # (1,C)
    tok = torch.mean(a, dim=1, keepdim=True)  # (N,1)
    tok_mean = tok.mean()
    tok = tok / (tok_mean + 1e-6)
    return tok.contiguous(), chan.t().contiguous()


def int2_levels(idx: torch.Tensor, thr: torch.Tensor) -> torch.Tensor:
    """fastpath.py:565-572 / :727-733, compress_quantize.py:744-750."""
    sign_bit = idx >> 1
    mag_bit = idx & 1
    small = 0.5 * thr
    large = 2.0 * thr
    level = torch.where(mag_bit == 0, small, large)
    mult = (sign_bit.to(torch.half) * 2.0) - 1.0
    return mult * level


def int2_quant(x: torch.Tensor, base: torch.Tensor, update_cache: bool, scales=None):
    """== int2_quant_fastpath (fastpath.py:584-669).  Returns packed (N,C/4) ndarray,
    U=tok (N,1), V=chan (C,1), new_base|None."""
    x, base = _h(x), _h(base)
    delta = x - base
    tok, chan = int2_scales(delta) if scales is None else scales
    thr = (chan.t() * tok).to(torch.half)  # fastpath.py:536
    sign_bit = (delta >= 0).to(torch.uint8)
    mag_bit = (torch.abs(delta) > thr).to(torch.uint8)  # strict, fastpath.py:540
    idx = (sign_bit << 1) | mag_bit
    packed = pack_int2(idx.numpy())
    new_base = None
    if update_cache:
        new_base = base + int2_levels(idx, thr).to(torch.half)
    return packed, tok, chan, new_base


def int2_dequant(packed: np.ndarray, tok_n1: torch.Tensor, chan_c1: torch.Tensor, base: torch.Tensor | None):
    """fastpath.py:672-741 (base given) / dequantize_int2 compress_quantize.py:707-753 (base None)."""
    idx = torch.from_numpy(unpack_int2(packed))
    thr = (_h(chan_c1).t() * _h(tok_n1)).to(torch.half)
    recv = int2_levels(idx, thr).to(torch.half)
    if base is None:
        return recv
    return _h(base) + recv


def sim_int2(x: torch.Tensor) -> torch.Tensor:
    """compress_quantize.py:339-384 (eager)."""
    x = _h(x)
    a = torch.abs(x)
    chan = torch.mean(a, dim=0, keepdim=True)
    tok = torch.mean(a, dim=1, keepdim=True)
    tok = tok / (tok.mean() + 1e-6)
    thr = (chan * tok).to(torch.half)
    out = torch.zeros_like(x)
    out = torch.where(x < -thr, -2.0 * thr, out)
    out = torch.where((x >= -thr) & (x < 0), -0.5 * thr, out)
    out = torch.where((x >= 0) & (x <= thr), 0.5 * thr, out)
    out = torch.where(x > thr, 2.0 * thr, out)
    return out


def sim_int2_minmax(x: torch.Tensor) -> torch.Tensor:
    """compress_quantize.py:386-426."""
    x = _h(x)
    mn = torch.min(x, dim=0, keepdim=True).values
    mx = torch.max(x, dim=0, keepdim=True).values
    scale = ((mx - mn) / (3 - 0 + 1e-6)).to(torch.half)
    q = torch.clamp(torch.round((x - mn) / scale), 0, 3)
    return q.to(torch.half) * scale + mn


# ----------------------------------------------------------------------------
# INT4 (per-channel min/max over N, two rows per byte along N)
# ----------------------------------------------------------------------------
def int4_params(x: torch.Tensor):
    """scale (1,C), min (1,C).  compress_quantize.py:551-558."""
    x = _h(x)
    mn = torch.min(x, dim=0, keepdim=True).values
    mx = torch.max(x, dim=0, keepdim=True).values
    scale = ((mx - mn) / (15 - 0 + 1e-6)).to(torch.half)
    return scale, mn.to(torch.half)


def int4_codes(x: torch.Tensor, scale: torch.Tensor, mn: torch.Tensor) -> np.ndarray:
    """(N,C) uint8 codes 0..15.  compress_quantize.py:561-564.  A zero scale
    (constant column) makes 0/0 = NaN; the uint8 cast of NaN is undefined in
    the reference -- we define NaN -> code 0."""
    q = torch.round((_h(x) - mn) / scale)
    q = torch.clamp(q, 0, 15)
    q = torch.nan_to_num(q.float(), nan=0.0)
    return q.to(torch.uint8).numpy()


def int4_pack(codes: np.ndarray) -> np.ndarray:
    """rows (2i, 2i+1) -> byte (i, c), low nibble = even row.  compress_quantize.py:566-573."""
    n, c = codes.shape
    assert n % 2 == 0
    v = codes.reshape(n // 2, 2, c)
    return ((v[:, 0, :] & 0x0F) | ((v[:, 1, :] & 0x0F) << 4)).astype(np.uint8)


def int4_unpack(packed: np.ndarray) -> np.ndarray:
    """compress_quantize.py:626-632."""
    n2, c = packed.shape
    out = np.empty((n2 * 2, c), dtype=np.uint8)
    out[0::2] = packed & 0x0F
    out[1::2] = (packed >> 4) & 0x0F
    return out


def int4_quantize(x: torch.Tensor):
    """== quantize_int4 eager (compress_quantize.py:527-583)."""
    scale, mn = int4_params(x)
    return int4_pack(int4_codes(x, scale, mn)), scale, mn


def int4_dequantize(packed: np.ndarray, scale: torch.Tensor, mn: torch.Tensor) -> torch.Tensor:
    """== dequantize_int4 eager (compress_quantize.py:594-640)."""
    q = torch.from_numpy(int4_unpack(packed))
    return q.to(torch.half) * scale + mn


def sim_int4(x: torch.Tensor, dim: int = 0) -> torch.Tensor:
    """compress_quantize.py:487-520.  NaN from a zero scale propagates here
    (that is what the reference's simulation does)."""
    x = _h(x)
    mx = torch.max(x, dim=dim, keepdim=True)[0]
    mn = torch.min(x, dim=dim, keepdim=True)[0]
    scale = ((mx - mn) / (15 + 1e-6)).to(torch.half)
    q = torch.clamp(torch.round((x - mn) / scale), min=0, max=15)
    return q.to(torch.half) * scale + mn


# ----------------------------------------------------------------------------
# INT8 (per-channel affine with int16 zero point; cache quantiser in the reference)
# ----------------------------------------------------------------------------
def int8_quantize(x: torch.Tensor):
    """== quantize_int8 eager (compress_quantize.py:428-471).
    Returns q (N,C) int8 ndarray, scale (1,C) fp16, zero_point (1,C) int16."""
    x = _h(x)
    qmin, qmax = -128, 127
    mn = torch.min(x, dim=0, keepdim=True).values
    mx = torch.max(x, dim=0, keepdim=True).values
    scale = ((mx - mn) / (qmax - qmin + 1e-6)).to(torch.half)
    zp = qmin - torch.round(mn / scale)
    zp = torch.nan_to_num(torch.clamp(zp, qmin, qmax).float(), nan=0.0).to(torch.int16)
    q = torch.round(x / scale + zp)
    q = torch.nan_to_num(torch.clamp(q, qmin, qmax).float(), nan=0.0).to(torch.int8)
    return q.numpy(), scale, zp


def int8_dequantize(q: np.ndarray, scale: torch.Tensor, zp: torch.Tensor) -> torch.Tensor:
    """== dequantize_int8 (compress_quantize.py:473-484)."""
    qt = torch.from_numpy(q)
    return (qt.half() - zp.half()) * scale


# ----------------------------------------------------------------------------
# SPARSE 1:m ("top-k": per-m-block argmax |x|, lowest index on ties)
# ----------------------------------------------------------------------------
def topk_compress(x_rows: torch.Tensor, m: int):
    """x_rows (A, 1024) -> val (A, 1024/m) fp16, idx (A, 512/m) uint8 ndarray;
    byte = idx_block1 << 4 | idx_block2 for consecutive m-blocks.
    compress_topk.py:11-104 (Triton argmax: lowest index wins ties)."""
    x_rows = _h(x_rows)
    a, w = x_rows.shape
    assert w % (2 * m) == 0 and m <= 16
    blocks = x_rows.view(a, w // m, m)
    mag = torch.abs(blocks).float()
    mag = torch.nan_to_num(mag, nan=-1.0)
    # lowest index among maxima
    mx = mag.max(dim=2, keepdim=True).values
    is_max = mag == mx
    idx = torch.argmax(is_max.to(torch.uint8), dim=2)  # first True
    val = torch.gather(blocks, 2, idx.unsqueeze(2)).squeeze(2)  # (a, w/m)
    idx_np = idx.numpy().astype(np.uint8).reshape(a, w // (2 * m), 2)
    packed = ((idx_np[..., 0] << 4) | idx_np[..., 1]).astype(np.uint8)
    return val.contiguous(), packed


def topk_decompress(val: torch.Tensor, packed_idx: np.ndarray, m: int) -> torch.Tensor:
    """compress_topk.py:108-163."""
    a, b = packed_idx.shape
    idx = np.empty((a, b, 2), dtype=np.int64)
    idx[..., 0] = (packed_idx >> 4) & 0xF
    idx[..., 1] = packed_idx & 0xF
    idx_t = torch.from_numpy(idx.reshape(a, 2 * b, 1))
    out = torch.zeros((a, 2 * b, m), dtype=torch.half)
    out.scatter_(2, idx_t, _h(val).view(a, 2 * b, 1))
    return out.view(a, 2 * b * m)


def sim_topk(x: torch.Tensor, m: int) -> torch.Tensor:
    """compress_topk.py:221-236 with the tie rule fixed to lowest index."""
    shape = x.shape
    rows = _h(x).view(-1, m)
    mag = torch.abs(rows).float()
    mx = mag.max(dim=1, keepdim=True).values
    idx = torch.argmax((mag == mx).to(torch.uint8), dim=1, keepdim=True)
    out = torch.zeros_like(rows)
    out.scatter_(1, idx, rows.gather(1, idx))
    return out.view(shape)


# ----------------------------------------------------------------------------
# low-rank subspace iteration
# ----------------------------------------------------------------------------
def subspace_iter(a: torch.Tensor, rank: int, num_iters: int = 2, init_q: torch.Tensor | None = None):
    """compress_lowrank.py:16-62.  Returns U (m,r), V (r,n), Q (n,r) in a.dtype.
    Without init_q it consumes the global torch CPU RNG exactly like the reference."""
    m, n = a.shape
    dtype = a.dtype
    af = a.float()
    if init_q is None:
        q = torch.randn(n, rank, dtype=torch.float)
        q, _ = torch.linalg.qr(q)
    else:
        q = init_q.float()
    for _ in range(num_iters):
        z = af.t() @ (af @ q)
        q, _ = torch.linalg.qr(z)
    u, _ = torch.linalg.qr(af @ q)
    v = u.t() @ af
    return u.to(dtype), v.to(dtype), q.to(dtype)


# ----------------------------------------------------------------------------
# slowpath payloads (one flat fp16 tensor)
# ----------------------------------------------------------------------------
def _as_half_view(u8: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(u8).reshape(-1)).view(torch.half)


def fastpath_payload(packed: np.ndarray, u: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """cat[q.view(half), U(N,K), V(C,K)]  (main.py:149-152)."""
    return torch.cat([_as_half_view(packed), u.reshape(-1), v.reshape(-1)])


def fastpath_split(payload: torch.Tensor, n: int, c: int, k: int, items_per_byte: int):
    """main.py:283-304."""
    qh = n * (c // items_per_byte) // 2
    q, u, v = torch.split(payload, [qh, n * k, c * k])
    packed = q.contiguous().view(torch.uint8).numpy().reshape(n, c // items_per_byte)
    return packed, u.reshape(n, k), v.reshape(c, k)


def slowpath_compress(x: torch.Tensor, ctype: str, rank=None, sparse_ratio=None) -> torch.Tensor:
    """slowpath.py:26-84.  ctype in {"binary","low-rank","low-rank-int4","sparse","int4"}.
    "int4" is our wire extension (the reference only simulates INT4, slowpath.py:205-206):
    [packed(N/2,C).view(half), scale(1,C), min(1,C)]."""
    x = _h(x)
    n, c = x.shape
    if ctype == "binary":
        assert rank == -1
        packed, u, v, _ = binary_quant(x, torch.zeros_like(x), False)
        parts = [_as_half_view(packed), u.reshape(-1), v.reshape(-1)]  # V (K,C)==(C,K) bytes for K=1
    elif ctype == "low-rank":
        u, v, _ = subspace_iter(x, rank, 2)
        parts = [u.reshape(-1), v.reshape(-1)]
    elif ctype == "low-rank-int4":
        u, v, _ = subspace_iter(x, rank, 2)
        qu, su, mu = int4_quantize(u)
        qv, sv, mv = int4_quantize(v.t().contiguous())
        parts = [_as_half_view(qu), su.reshape(-1), mu.reshape(-1), _as_half_view(qv), sv.reshape(-1), mv.reshape(-1)]
    elif ctype == "sparse":
        val, idx = topk_compress(x.view(-1, SPARSE_LAST_DIM_SIZE), sparse_ratio)
        parts = [val.reshape(-1), _as_half_view(idx)]
    elif ctype == "int4":
        q, s, mn = int4_quantize(x)
        parts = [_as_half_view(q), s.reshape(-1), mn.reshape(-1)]
    else:
        raise ValueError(f"Invalid compress_type value: {ctype}")
    return torch.cat(parts)


def slowpath_decompress(p: torch.Tensor, shape, ctype: str, rank=None, sparse_ratio=None) -> torch.Tensor:
    """slowpath.py:86-175."""
    n, c = shape
    numel = n * c
    if ctype == "binary":
        packed, u, v = fastpath_split(p, n, c, 1, 8)
        return binary_dequant(packed, u, v, None)
    if ctype == "low-rank":
        u, v = torch.split(p, [n * rank, rank * c])
        return torch.matmul(u.view(n, rank), v.view(rank, c))
    if ctype == "low-rank-int4":
        sizes = [n * rank // 4, rank, rank, c * rank // 4, rank, rank]
        qu, su, mu, qv, sv, mv = torch.split(p, sizes)
        u = int4_dequantize(qu.contiguous().view(torch.uint8).numpy().reshape(n // 2, rank), su.view(1, rank), mu.view(1, rank))
        v = int4_dequantize(qv.contiguous().view(torch.uint8).numpy().reshape(c // 2, rank), sv.view(1, rank), mv.view(1, rank))
        return torch.matmul(u, v.t())
    if ctype == "sparse":
        val, idx = torch.split(p, [numel // sparse_ratio, numel // sparse_ratio // 4])
        a = numel // SPARSE_LAST_DIM_SIZE
        idx_np = idx.contiguous().view(torch.uint8).numpy().reshape(a, SPARSE_LAST_DIM_SIZE // sparse_ratio // 2)
        return topk_decompress(val.view(a, -1), idx_np, sparse_ratio).view(shape)
    if ctype == "int4":
        q, s, mn = torch.split(p, [numel // 4, c, c])
        return int4_dequantize(q.contiguous().view(torch.uint8).numpy().reshape(n // 2, c), s.view(1, c), mn.view(1, c))
    raise ValueError(f"Invalid compress_type value: {ctype}")


def sim_compress(x: torch.Tensor, ctype: str, sparse_ratio=None, rank=None) -> torch.Tensor:
    """slowpath.py:185-239 (compress then decompress, no size reduction)."""
    if ctype == "identity":
        return x
    if ctype == "sparse":
        return sim_topk(x, sparse_ratio)
    if ctype == "binary":
        assert rank == -1
        return sim_binary(x.half()).half()
    if ctype == "int2":
        return sim_int2(x)
    if ctype == "int2-minmax":
        return sim_int2_minmax(x)
    if ctype == "int4":
        return sim_int4(x, dim=0)
    if ctype == "low-rank":
        u, v, _ = subspace_iter(x, rank, 2)
        return torch.matmul(u, v)
    if ctype == "low-rank-int4":
        u, v, _ = subspace_iter(x, rank, 2)
        return torch.matmul(sim_int4(u, dim=0), sim_int4(v, dim=1))
    raise ValueError("Invalid compress_type value")
