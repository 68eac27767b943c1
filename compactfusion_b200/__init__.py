"""compactfusion_b200 -- B200-native (sm_100a) implementation of CompactFusion's
residual-compression communication hot path, behind the reference's `xfuser.compact` API.

Importing the package does not load the CUDA library; the first codec call does, and raises
if `libcompactb200.so` is missing (there is no CPU fallback).
"""
from .utils import ALLOW_DEPRECATED, COMPACT_COMPRESS_TYPE, CompactCache, CompactConfig  # noqa: F401
from .patchpara.df_utils import PatchConfig  # noqa: F401
from .main import (  # noqa: F401
    allgather_cache, compact_all_gather, compact_cache, compact_compress, compact_config, compact_decompress,
    compact_get_step, compact_hello, compact_init, compact_reset, compact_set_inplace, compact_set_step,
)
from .ring import compact_fwd  # noqa: F401

__version__ = "0.1.0"
