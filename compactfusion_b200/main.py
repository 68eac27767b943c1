"""Plugin state + residual / error-feedback state machine + compressed all-gather
(mirror of xfuser/compact/main.py; same function names, signatures and wire formats).

What differs from the reference, by design:
  * fastpath payloads are written by the kernels straight into one flat buffer
    (no torch.cat, main.py:149-152), and split by views (no torch.split copies);
  * `compact_all_gather` gathers into one buffer and decompresses all W peers in ONE batched
    launch (reference: W launches, main.py:410-419);
  * residual-1 slowpath codecs fuse the residual subtract / add into their kernels where a
    fused entry point exists (INT4, SPARSE, LOW_RANK);
  * optional in-place cache update (`compact_set_inplace(True)`) reuses the cached base
    buffer for the new base when the cache owns it (never a caller's tensor).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _native as nv
from .compress_lowrank import lowrank_project, lowrank_reconstruct
from .compress_quantize import _minmax_compress, _sign_compress
from .compress_topk import _topk_compress
from .fastpath import binary_dequant_fastpath, int2_dequant_fastpath
from .patchpara.df_cache import AllGatherCache
from .prof import Profiler
from .slowpath import sim_compress, slowpath_compress, slowpath_decompress
from .stats import stats_clear, stats_hello, stats_log
from .utils import ALLOW_DEPRECATED, COMPACT_COMPRESS_TYPE, CompactCache, CompactConfig

T = COMPACT_COMPRESS_TYPE

_config: CompactConfig | None = None
_cache: CompactCache | None = None
_step = None
_allgather_cache: AllGatherCache | None = None
_current_cache_key = None
_inplace = False
_owned: set = set()  # cache keys whose base tensor was allocated by this module


# ------------------------------------------------------------------------------ state API
def compact_init(config: CompactConfig):
    """main.py:37-52.  Must precede pipeline construction."""
    global _config, _cache, _step, _allgather_cache, _current_cache_key
    from . import dropin
    dropin.shutdown()  # engines behind the hooks are built for one configuration
    _config = config
    _cache = CompactCache(quantize=config.quantized_cache)
    _owned.clear()
    _step = None
    if config.override_with_patch_gather_fwd:
        _allgather_cache = AllGatherCache()
    _current_cache_key = None


def compact_hello():
    """main.py:54-71 (plain-text banner on rank 0)."""
    if dist.is_initialized() and dist.get_rank() != 0:
        return
    c = _config
    print("--- compactfusion_b200 (sm_100a) initialized ---")
    print("compact enabled" if c.enabled else "compact disabled")
    if c.enabled:
        if not c.override_with_patch_gather_fwd:
            print(f"fastpath={c.fastpath} simulate={c.simulate_compress} residual={c.compress_residual} "
                  f"ef={c.error_feedback} check_consistency={c.check_cache_consistency}")
        else:
            pc = c.patch_gather_fwd_config
            print(f"patch-parallel all-gather: async(DistriFusion)={pc.async_comm} compact={pc.use_compact}")
    if c.log_compress_stats and c.enabled:
        stats_hello()


def compact_config():
    return _config


def compact_set_step(step):
    global _step
    _step = step


def compact_get_step():
    return _step


def compact_cache():
    return _cache


def allgather_cache():
    return _allgather_cache


def compact_reset():
    """Drop all cached bases (per image).  main.py:93-106."""
    global _cache, _step, _allgather_cache, _current_cache_key
    _cache = CompactCache(quantize=_config.quantized_cache)
    _owned.clear()
    stats_clear()
    _step = None
    if _config.override_with_patch_gather_fwd:
        _allgather_cache = AllGatherCache()
    _current_cache_key = None


def compact_get_current_cache_key():
    return _current_cache_key


def compact_set_inplace(flag: bool):
    """Opt in to updating cache-owned base buffers in place (halves allocator traffic; the
    tensor returned by compact_decompress for a key is then overwritten by the next call
    for that key, which the reference's callers never observe: they consume it at once)."""
    global _inplace
    _inplace = bool(flag)


# ------------------------------------------------------------------------------ helpers
def _to_2d_shape(shape):
    """(..., h, d) -> (prod(...), h*d); (b, s, c) -> (b*s, c).  main.py:180-185, :333-342."""
    if len(shape) >= 4:
        rows = 1
        for s in shape[:-2]:
            rows *= s
        return (rows, shape[-2] * shape[-1])
    if len(shape) == 3:
        return (shape[0] * shape[1], shape[2])
    assert len(shape) == 2
    return tuple(shape)


def _effective_rank():
    return 1 if _config.comp_rank == -1 else _config.comp_rank


def fastpath_payload_numel(n: int, c: int, compress_type, k: int = 1) -> int:
    per_byte = 8 if compress_type == T.BINARY else 4
    return n * (c // per_byte) // 2 + n * k + c * k


def _payload_views(payload: torch.Tensor, n: int, c: int, compress_type, k: int = 1):
    """Views (packed u8 (N,C/per_byte), U (N,K), V (C,K)) into a flat fp16 payload (main.py:283-304)."""
    per_byte = 8 if compress_type == T.BINARY else 4
    qh = n * (c // per_byte) // 2
    assert payload.numel() == qh + n * k + c * k, \
        f"Mismatch in compressed tensor size: expected {qh + n * k + c * k}, got {payload.numel()}, (N,C)=({n},{c}), K={k}"
    packed = payload[:qh].view(torch.uint8).view(n, c // per_byte)
    return packed, payload[qh:qh + n * k].view(n, k), payload[qh + n * k:].view(c, k)


def _new_base_buffer(key, base):
    """Where the updated base goes: in place when allowed and the cache owns `base`."""
    # (with log_stats the old base is still needed after the update: never alias it)
    if _inplace and key in _owned and base is not None and not _config.log_compress_stats:
        return base
    return torch.empty_like(base)


def _put(key, val, delta=None, owned=True):
    _cache.put(key, val, delta)
    if owned:
        _owned.add(key)
    else:
        _owned.discard(key)


# ------------------------------------------------------------------------------ compress
def _compact_compress_fastpath(cache_key, x, compress_type, update_cache: bool, rank: int):
    """main.py:130-166."""
    assert compress_type in (T.BINARY, T.INT2)
    assert _config.compress_residual == 1
    base = _cache.get_base(cache_key)
    assert base is not None, f"no cached base for key {cache_key}: run a WARMUP step first"
    n, c = x.shape
    if compress_type == T.BINARY and rank != -1:
        from .fastpath import binary_quant_fastpath
        q, u, v, new_base = binary_quant_fastpath(x, base, rank, update_cache)
        payload = torch.cat([q.view(torch.half).flatten(), u.flatten(), v.flatten()])
    else:
        payload = torch.empty(fastpath_payload_numel(n, c, compress_type), dtype=torch.half, device=x.device)
        packed, u, v = _payload_views(payload, n, c, compress_type)
        codec = nv.CODEC_BINARY if compress_type == T.BINARY else nv.CODEC_INT2
        nb = _new_base_buffer(cache_key, base) if update_cache else None
        _, _, _, new_base = _sign_compress(codec, x, base, update_cache, packed=packed, u=u, v=v, new_base=nb)
    if update_cache:
        _put(cache_key, new_base)
    if _config.log_compress_stats:
        stats_log().log(cache_key, base, None, x, new_base, payload, 1)
    return payload


def _compress_fn(x, compress_type, rank):
    if _config.simulate_compress:
        return sim_compress(x, compress_type, _config.sparse_ratio, rank)
    return slowpath_compress(x, compress_type, rank=rank, sparse_ratio=_config.sparse_ratio)


def _decompress_fn(x, compress_type, shape, rank):
    if _config.simulate_compress:
        return x.view(shape)
    return slowpath_decompress(x, shape, compress_type, rank=rank, sparse_ratio=_config.sparse_ratio)


def _residual1_fused(x, base, compress_type, rank):
    """Residual-1 compress with the subtract / add fused into the codec kernels.
    Returns (payload, reconstructed) or None if this (type, mode) has no fused path."""
    sim = _config.simulate_compress
    n, c = x.shape
    if compress_type == T.INT4 and n % 2 == 0 and not sim:
        codes, scale, mn, recon = _minmax_compress(nv.CODEC_INT4, x, base, want_recon=True)
        return torch.cat([codes.view(torch.half).flatten(), scale.flatten(), mn.flatten()]), recon
    if compress_type == T.SPARSE and not sim:
        val, idx, recon = _topk_compress(x, base, _config.sparse_ratio, want_new_base=True)
        return torch.cat([val, idx.view(torch.half)]), recon
    if compress_type == T.LOW_RANK and not sim:
        payload = torch.empty(rank * (n + c), dtype=torch.half, device=x.device)
        u, v = payload[:n * rank].view(n, rank), payload[n * rank:].view(rank, c)
        lowrank_project(x, base, rank, 2, u_out=u, v_out=v)
        return payload, lowrank_reconstruct(u, v, base)
    return None


@Profiler.prof_func("compact.compact_compress")
def compact_compress(cache_key, x: torch.Tensor, compress_type: COMPACT_COMPRESS_TYPE, update_cache: bool = False):
    """main.py:169-270."""
    global _current_cache_key
    _current_cache_key = cache_key
    assert x.is_contiguous()
    assert _config.enabled
    original_shape = x.shape
    x = x.view(_to_2d_shape(x.shape))
    rank = _config.comp_rank
    if compress_type == T.BINARY and rank != -1:
        assert ALLOW_DEPRECATED, "Binary compression with rank != -1 is deprecated"
    residual = _config.compress_residual

    if compress_type == T.WARMUP:
        if update_cache:
            if _config.fastpath or residual == 1:
                _put(cache_key, x, None, owned=False)  # the caller's tensor itself (main.py:199)
            elif residual == 2:
                base = _cache.get_base(cache_key)
                _put(cache_key, x, None if base is None else x - base, owned=False)
        return x.view(original_shape)

    if _config.fastpath:
        return _compact_compress_fastpath(cache_key, x, compress_type, update_cache, rank)

    if residual == 0:
        compressed = _compress_fn(x, compress_type, rank)
        if _config.log_compress_stats:  # main.py:216-228
            stats_log().log(cache_key, None, None, x, _decompress_fn(compressed, compress_type, x.shape, rank),
                            compressed, 0)
        return compressed
    if residual == 1:
        base = _cache.get_base(cache_key)
        fused = _residual1_fused(x, base, compress_type, rank)
        if fused is not None:
            compressed, reconstructed = fused
        else:
            delta = x - base
            compressed = _compress_fn(delta, compress_type, rank)
            reconstructed = base + _decompress_fn(compressed, compress_type, x.shape, rank)
        if update_cache:
            if _config.error_feedback:
                _put(cache_key, reconstructed)
            else:
                _put(cache_key, x, owned=False)  # main.py:233
        if _config.log_compress_stats:  # main.py:234-245
            stats_log().log(cache_key, base, None, x, reconstructed, compressed, 1)
        return compressed
    if residual == 2:
        base = _cache.get_base(cache_key)
        delta_base = _cache.get_delta_base(cache_key)
        delta_delta = x - base - delta_base
        compressed = _compress_fn(delta_delta, compress_type, rank)
        recv_dd = _decompress_fn(compressed, compress_type, x.shape, rank)
        new_base = base + delta_base + recv_dd
        if update_cache:
            _put(cache_key, new_base, _decay_delta_base(delta_base + recv_dd))
        if _config.log_compress_stats:  # main.py:257-268
            stats_log().log(cache_key, base, delta_base, x, new_base, compressed, 2)
        return compressed
    raise ValueError("Invalid compress_residual value")


def _decay_delta_base(delta_base):
    return delta_base * _config.delta_decay_factor


# ------------------------------------------------------------------------------ decompress
def _compact_decompress_fastpath(cache_key, compressed, compress_type, shape, update_cache: bool, rank: int):
    """main.py:276-319."""
    assert compress_type in (T.BINARY, T.INT2)
    assert _config.compress_residual == 1
    n, c = shape
    packed, u, v = _payload_views(compressed, n, c, compress_type, rank)
    base = _cache.get_base(cache_key)
    assert base is not None, f"no cached base for key {cache_key}: run a WARMUP step first"
    fn = binary_dequant_fastpath if compress_type == T.BINARY else int2_dequant_fastpath
    out = _new_base_buffer(cache_key, base) if update_cache else None
    recon = fn(packed, u, v, base, out=out)
    if update_cache:
        _put(cache_key, recon)
    return recon


@Profiler.prof_func("compact.compact_decompress")
def compact_decompress(cache_key, compressed: torch.Tensor, compress_type: COMPACT_COMPRESS_TYPE, shape: tuple,
                       update_cache: bool = False):
    """main.py:322-388."""
    global _current_cache_key
    _current_cache_key = cache_key
    assert _config.enabled
    original_shape = tuple(shape)
    shape = _to_2d_shape(original_shape)
    rank = _effective_rank()
    residual = _config.compress_residual

    if compress_type == T.WARMUP:
        val = compressed.view(shape)
        if update_cache:
            if _config.fastpath or residual == 1:
                _put(cache_key, val, None, owned=False)
            elif residual == 2:
                base = _cache.get_base(cache_key)
                _put(cache_key, val, None if base is None else val - base, owned=False)
        return val.view(original_shape)

    if _config.fastpath:
        return _compact_decompress_fastpath(cache_key, compressed, compress_type, shape, update_cache, rank).view(
            original_shape)
    if residual == 0:
        return _decompress_fn(compressed, compress_type, shape, rank).view(original_shape)
    if residual == 1:
        base = _cache.get_base(cache_key)
        reconstructed = base + _decompress_fn(compressed, compress_type, shape, rank)
        if update_cache:
            _put(cache_key, reconstructed)
        return reconstructed.view(original_shape)
    if residual == 2:
        base = _cache.get_base(cache_key)
        delta_base = _cache.get_delta_base(cache_key)
        recv_dd = _decompress_fn(compressed, compress_type, shape, rank)
        reconstructed = base + delta_base + recv_dd
        if update_cache:
            _put(cache_key, reconstructed, _decay_delta_base(delta_base + recv_dd))
        return reconstructed.view(original_shape)
    raise ValueError("Invalid compress_residual value")


# ------------------------------------------------------------------------------ all-gather
def _decompress_peers_batched(tags, payloads, compress_type, shape2d):
    """One launch reconstructs every peer's tensor of a fastpath all-gather and updates the
    caches (replaces the W sequential compact_decompress calls of main.py:410-419)."""
    n, c = shape2d
    views = [_payload_views(p, n, c, compress_type) for p in payloads]
    bases = [_cache.get_base(t) for t in tags]
    for t, b in zip(tags, bases):
        assert b is not None, f"no cached base for key {t}: run a WARMUP step first"
    outs = [_new_base_buffer(t, b) for t, b in zip(tags, bases)]
    fn = nv.lib().cf_binary_decompress_batched if compress_type == T.BINARY else nv.lib().cf_int2_decompress_batched
    for s in range(0, len(tags), nv.CF_MAX_BATCH):
        e = min(len(tags), s + nv.CF_MAX_BATCH)
        rc = fn(e - s, nv.ptr_array([v[0] for v in views[s:e]]), nv.ptr_array([v[1] for v in views[s:e]]),
                nv.ptr_array([v[2] for v in views[s:e]]), nv.ptr_array(bases[s:e]), nv.ptr_array(outs[s:e]), n, c,
                nv.stream_ptr())
        nv.check(rc, "cf_*_decompress_batched")
    for t, o in zip(tags, outs):
        _put(t, o)
    return outs


def compact_all_gather(tag, x: torch.Tensor, comp_type: COMPACT_COMPRESS_TYPE, group=None):
    """compress (no cache update) -> all-gather payloads -> decompress every origin (own shard
    included) with cache update.  main.py:390-420.  Returns W tensors of x.shape."""
    assert _config.enabled
    rank = dist.get_rank(group)
    world_size = dist.get_world_size(group)
    to_send = compact_compress(f"{tag}-{rank}", x, comp_type, update_cache=False)
    flat = to_send.reshape(-1)
    gathered = torch.empty((world_size, flat.numel()), dtype=flat.dtype, device=flat.device)
    with Profiler.scope("compact.all_gather"):
        dist.all_gather_into_tensor(gathered.view(-1), flat, group=group)
    bufs = [gathered[i] for i in range(world_size)]
    tags = [f"{tag}-{i}" for i in range(world_size)]
    if _config.fastpath and comp_type in (T.BINARY, T.INT2) and _config.comp_rank == -1:
        outs = _decompress_peers_batched(tags, bufs, comp_type, _to_2d_shape(x.shape))
        return [o.view(x.shape) for o in outs]
    return [compact_decompress(t, b.view(to_send.shape) if comp_type == T.WARMUP else b, comp_type, x.shape,
                               update_cache=True) for t, b in zip(tags, bufs)]
