"""Patch-parallel attention: all-gather K/V, then one attention over the full sequence
(mirror of xfuser/compact/patchpara/fwd.py:20-237).  Three modes: compact (compressed
all-gather, `compact_all_gather`), synchronous uncompressed, stale-async (DistriFusion)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from .. import dropin
from ..attention import attn_forward
from ..prof import Profiler
from .df_cache import DummyHandle
from .df_utils import PatchConfig

_buffers = {}
_main = None
_JOINT_ALLOWED = ["front", "rear", "none"]


def _joint_flags(joint_tensor_key, joint_tensor_value, joint_strategy, allowed):
    """fwd.py:60-76 of the reference (same checks as ring.py's)."""
    if joint_tensor_key is None and joint_tensor_value is None:
        return False
    if joint_tensor_key is not None and joint_tensor_value is not None:
        if joint_strategy not in allowed:
            raise ValueError(f"joint_strategy: {joint_strategy} not supprted. supported joint strategy: {allowed}")
        return joint_strategy != "none"
    raise ValueError("joint_tensor_key and joint_tensor_value should be None or not None simultaneously.")


def _plugin():
    global _main
    if _main is None:
        from .. import main as m   # (main imports this package: resolved at first call)
        _main = m
    return _main


@Profiler.prof_func("patch_gather_fwd.gather_patch_fwd")
def patch_gather_fwd(q, k, v, dropout_p=0, softmax_scale=None, causal=True, window_size=(-1, -1),
                     alibi_slopes=None, return_attn_probs=None, deterministic=False, attn_layer=None, group=None,
                     joint_tensor_key=None, joint_tensor_value=None, joint_strategy="none", mod_idx=None,
                     current_iter=None):
    m = _plugin()
    cfg = m.compact_config()
    assert alibi_slopes is None, "Alibi slopes not supported in this basic gather impl."
    if softmax_scale is None:
        softmax_scale = q.shape[-1] ** (-0.5)
    assert cfg.override_with_patch_gather_fwd, "Patch gather fwd is not enabled"
    config: PatchConfig = cfg.patch_gather_fwd_config
    assert mod_idx is not None, "mod_idx is required for caching"
    assert current_iter is not None, "current_iter is required for async logic"
    is_joint = _joint_flags(joint_tensor_key, joint_tensor_value, joint_strategy, _JOINT_ALLOWED)

    q, k, v = q.contiguous(), k.contiguous(), v.contiguous()

    key_to_use = value_to_use = None
    if config.use_compact:
        ctype = cfg.compress_func(mod_idx, current_iter)
        ent = dropin.hot(cfg, "patch", group, k, mod_idx, ctype)
        if ent is not None:
            # K and V of the layer through the persistent-buffer engine: one compress(+put) launch pair, one
            # reconstruct launch for all W origins, straight into the buffer attention reads (no cat for bs == 1)
            eng, layer, views = ent[0], ent[1], ent[2]
            gk, gv = eng.exchange(layer, k, v, ctype)
            if views is not None:
                key_to_use, value_to_use = views
            else:
                key_to_use = dropin.as_sequence(gk, eng.world, k.shape)
                value_to_use = dropin.as_sequence(gv, eng.world, v.shape)
        else:
            k_list = m.compact_all_gather(f"{mod_idx}-k", k, comp_type=ctype, group=group)
            v_list = m.compact_all_gather(f"{mod_idx}-v", v, comp_type=ctype, group=group)
    elif not config.async_comm:
        world_size, rank = dropin.group_info(group)
        k_list = [torch.empty_like(k) for _ in range(world_size)]
        v_list = [torch.empty_like(v) for _ in range(world_size)]
        with Profiler.scope("compact.gather.all_gather_sync"):
            dist.all_gather(k_list, k, group=group)
            dist.all_gather(v_list, v, group=group)
    else:
        world_size, rank = dropin.group_info(group)
        cache = m.allgather_cache()
        kk, vk = f"{mod_idx}-k", f"{mod_idx}-v"
        with Profiler.scope("df.all_gather"):
            if current_iter < config.async_warmup:
                if _buffers.get(kk) is None or _buffers[kk][0].shape != k.shape:
                    _buffers[kk] = [torch.empty_like(k) for _ in range(world_size)]
                    _buffers[vk] = [torch.empty_like(v) for _ in range(world_size)]
                k_list, v_list = _buffers[kk], _buffers[vk]
                dist.all_gather(k_list, k, group=group)
                dist.all_gather(v_list, v, group=group)
                cache.put(kk, DummyHandle(), k_list, k)
                cache.put(vk, DummyHandle(), v_list, v)
            else:
                if not cache.contains(kk) or not cache.contains(vk):
                    raise RuntimeError(f"DistriFusion cache miss for key {kk} or {vk} at iter {current_iter}. "
                                       "Check async_warmup steps.")
                hk, prev_k, _ = cache.get(kk)
                hv, prev_v, _ = cache.get(vk)
                hk.wait()
                hv.wait()
                k_list = [b.clone() for b in prev_k]  # stale K/V of the other ranks
                v_list = [b.clone() for b in prev_v]
                k_list[rank], v_list[rank] = k.clone(), v.clone()  # fresh local shard
                nk, nv_ = _buffers[kk], _buffers[vk]
                cache.put(kk, dist.all_gather(nk, k, group=group, async_op=True), nk, k)
                cache.put(vk, dist.all_gather(nv_, v, group=group, async_op=True), nv_, v)

    if key_to_use is None:
        key_to_use = torch.cat(k_list, dim=1)
        value_to_use = torch.cat(v_list, dim=1)
    if is_joint and joint_strategy == "front":
        key_to_use = torch.cat([joint_tensor_key, key_to_use], dim=1)
        value_to_use = torch.cat([joint_tensor_value, value_to_use], dim=1)
    elif is_joint and joint_strategy == "rear":
        key_to_use = torch.cat([key_to_use, joint_tensor_key], dim=1)
        value_to_use = torch.cat([value_to_use, joint_tensor_value], dim=1)

    out, lse = attn_forward(q, key_to_use, value_to_use, dropout_p, softmax_scale, causal=causal,
                            window_size=window_size)
    # (b, h, s) like the flash-attn LSE the reference post-processes (fwd.py:234-235 is a no-op reshape
    # chain on that layout); callers of the patch path only use `out`
    return (out if out.dtype == q.dtype else out.to(q.dtype)), lse, None
