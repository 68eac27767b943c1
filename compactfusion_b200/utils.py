"""Config, enum and cache of the compact plugin (mirror of xfuser/compact/utils.py).

Same names, constructor arguments, validity rules and error behaviour as the reference
(`COMPACT_COMPRESS_TYPE` utils.py:10-28, `CompactConfig` :31-117, `CompactCache` :123-196) so
that reference-side code (examples/configs.py presets, the xDiT hooks) works unchanged.
"""
from __future__ import annotations

import os
from enum import Enum

import torch
import torch.distributed as dist

from .patchpara.df_utils import PatchConfig

ALLOW_DEPRECATED = os.environ.get("COMPACT_ALLOW_DEPRECATED", "0") == "1"


class COMPACT_COMPRESS_TYPE(Enum):
    """Compression kinds; values are the reference's strings (utils.py:19-28)."""

    WARMUP = "warmup"
    SPARSE = "sparse"
    BINARY = "binary"
    INT2 = "int2"
    INT2_MINMAX = "int2-minmax"
    INT4 = "int4"
    IDENTITY = "identity"
    LOW_RANK = "low-rank"
    LOW_RANK_Q = "low-rank-int4"
    LOW_RANK_AWL = "low-rank-awl"


class CompactConfig:
    """Settings of the plugin; argument meaning as in the reference (utils.py:33-62)."""

    def __init__(
        self,
        enabled: bool = False,
        override_with_patch_gather_fwd: bool = False,
        patch_gather_fwd_config: PatchConfig = None,
        compress_func: callable = None,
        sparse_ratio=None,
        comp_rank=None,
        residual: int = 0,
        ef: bool = False,
        simulate: bool = False,
        log_stats: bool = False,
        check_consist: bool = False,
        fastpath: bool = False,
        quantized_cache: bool = False,
        delta_decay_factor: float | None = None,
    ):
        assert residual in [0, 1, 2]
        self.enabled = enabled
        self.compress_func = compress_func
        self.sparse_ratio = sparse_ratio
        self.comp_rank = comp_rank
        self.compress_residual = residual
        self.error_feedback = ef
        self.simulate_compress = simulate
        self.log_compress_stats = log_stats
        self.check_cache_consistency = check_consist
        self.fastpath = fastpath
        self.quantized_cache = quantized_cache
        self.delta_decay_factor = delta_decay_factor
        self.override_with_patch_gather_fwd = override_with_patch_gather_fwd
        self.patch_gather_fwd_config = patch_gather_fwd_config

        # cross-field rules, utils.py:83-106
        if residual == 0:
            assert not ef, "No residual does not support error feedback."
        if residual == 2:
            assert ef, "2nd order compression requires error feedback enabled."
        if fastpath:
            assert ef, "Fastpath requires error feedback enabled."
            assert not simulate, "Fastpath does not support simulation."
            assert residual == 1, "Fastpath requires 1st order residual."
        if quantized_cache:
            # int8 cache storage: deprecated in the reference, gated the same way (utils.py:128-129)
            assert ALLOW_DEPRECATED, "quantized_cache is deprecated"
        if override_with_patch_gather_fwd:
            assert enabled, "Compact must be enabled if override_with_patch_gather_fwd is True"
            assert patch_gather_fwd_config is not None, \
                "patch_gather_fwd_config must be set if override_with_patch_gather_fwd is True"
            if patch_gather_fwd_config.use_compact:
                assert not patch_gather_fwd_config.async_comm, "Compact does not support async communication"
        else:
            assert patch_gather_fwd_config is None, \
                "patch_gather_fwd_config must be None if override_with_patch_gather_fwd is False"

    def get_compress_type(self):
        """Name used for result files (utils.py:108-117)."""
        if self.compress_func is None or not self.enabled:
            return "NO_COMPACT"
        t = self.compress_func(0, 4)
        return t.name if isinstance(t, COMPACT_COMPRESS_TYPE) else str(t)


class CompactCache:
    """key -> base / delta_base tensors (utils.py:123-160).

    Keys are `"{layer}-{origin_rank}-k|v"` (ring.py:184-185) or `"{layer}-k|v-{rank}"`
    (main.py:399,412).  Unlike the reference, `put` does not call the activation collector
    unless one was installed.
    """

    def __init__(self, quantize=False):
        self.quantize = quantize
        self.base = {}
        self.delta_base = {}
        if quantize:
            assert ALLOW_DEPRECATED  # utils.py:128-129
        self.passed_count = 0

    def put(self, key, base, delta_base):
        """With `quantize` the base is stored as per-channel int8 (q, scale, zero_point) and dequantised on
        every read (utils.py:134-136, :151-155: quantize_int8 / dequantize_int8, here the sm_100a codec)."""
        if self.quantize:
            from .compress_quantize import quantize_int8
            base = quantize_int8(base.reshape(-1, base.shape[-1]) if base.dim() != 2 else base) + (base.shape,)
        self.base[key] = base
        self.delta_base[key] = delta_base

    def get_base(self, key):
        base = self.base.get(key, None)
        if self.quantize and base is not None:
            from .compress_quantize import dequantize_int8
            q, scale, zp, shape = base
            base = dequantize_int8(q, scale, zp).view(shape)
        return base

    def get_delta_base(self, key):
        return self.delta_base.get(key, None)

    def check_consistency(self, group=None):
        """All ranks hold the same cache: all-reduce mean vs local, atol 1e-2 (utils.py:164-196)."""
        if group is None:
            group = dist.group.WORLD
        world_size = dist.get_world_size(group)
        if world_size <= 1:
            return
        for key in sorted(self.base.keys()):
            parts = [t.flatten() for t in (self.get_base(key), self.get_delta_base(key)) if t is not None]
            if not parts:
                continue
            local = torch.cat(parts).float()
            summed = local.clone()
            dist.all_reduce(summed, op=dist.ReduceOp.SUM, group=group)
            mean = summed / world_size
            assert torch.allclose(local, mean, atol=1e-2), \
                f"Inconsistent cache at key {key}, max diff: {torch.max(torch.abs(local - mean)):.6f}"
        self.passed_count += 1
