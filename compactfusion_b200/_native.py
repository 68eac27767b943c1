"""ctypes binding of libcompactb200.so (the C ABI declared in include/compactb200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is
raised.  Tensors are passed as raw device pointers; all work is enqueued on torch's
current CUDA stream.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_int, c_int64, c_size_t, c_void_p, POINTER

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcompactb200.so")

CF_MAX_BATCH = 16
CODEC_BINARY, CODEC_INT2, CODEC_INT4, CODEC_INT8, CODEC_TOPK, CODEC_LOWRANK = 1, 2, 4, 8, 16, 32
PASS_STATS, PASS_FINALIZE, PASS_ENCODE, PASS_ALL = 1, 2, 4, 7
FLAG_INPUTS_STABLE = 0x100  # include/compactb200.h: enum cf_flag

# every symbol include/compactb200.h declares: (restype, argtypes)
_VPP = POINTER(c_void_p)
SYMBOLS = {
    "cf_abi_version": (c_int, []),
    "cf_last_error": (c_char_p, []),
    "cf_sm_count": (c_int, []),
    "cf_workspace_bytes": (c_size_t, [c_int, c_int64, c_int64, c_int, c_int]),
    "cf_binary_compress": (c_int, [c_void_p] * 6 + [c_int64, c_int64, c_void_p, c_size_t, c_void_p]),
    "cf_binary_compress_batched": (c_int, [c_int] + [_VPP] * 6 + [c_int64, c_int64, c_void_p, c_size_t, c_void_p]),
    "cf_binary_decompress": (c_int, [c_void_p] * 3 + [c_int, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "cf_binary_decompress_batched": (c_int, [c_int] + [_VPP] * 5 + [c_int64, c_int64, c_void_p]),
    "cf_int2_compress": (c_int, [c_void_p] * 6 + [c_int64, c_int64, c_void_p, c_size_t, c_void_p]),
    "cf_int2_compress_batched": (c_int, [c_int] + [_VPP] * 6 + [c_int64, c_int64, c_void_p, c_size_t, c_void_p]),
    "cf_int2_decompress": (c_int, [c_void_p] * 5 + [c_int64, c_int64, c_void_p]),
    "cf_int2_decompress_batched": (c_int, [c_int] + [_VPP] * 5 + [c_int64, c_int64, c_void_p]),
    "cf_int2_encode_with_scales": (c_int, [c_void_p] * 6 + [c_int64, c_int64, c_void_p]),
    "cf_sign_compress_passes": (c_int, [c_int, c_int, c_int] + [_VPP] * 6 + [c_int64, c_int64, c_void_p, c_size_t, c_void_p]),
    "cf_int4_compress": (c_int, [c_void_p] * 6 + [c_int64, c_int64, c_void_p, c_size_t, c_void_p]),
    "cf_int2mm_compress": (c_int, [c_void_p] * 6 + [c_int64, c_int64, c_void_p, c_size_t, c_void_p]),
    "cf_int4_decompress": (c_int, [c_void_p] * 5 + [c_int64, c_int64, c_void_p]),
    "cf_int8_compress": (c_int, [c_void_p] * 6 + [c_int64, c_int64, c_void_p, c_size_t, c_void_p]),
    "cf_int8_decompress": (c_int, [c_void_p] * 5 + [c_int64, c_int64, c_void_p]),
    "cf_topk_compress": (c_int, [c_void_p] * 5 + [c_int64, c_int, c_void_p]),
    "cf_topk_decompress": (c_int, [c_void_p] * 4 + [c_int64, c_int, c_void_p]),
    "cf_lowrank_project": (c_int, [c_void_p] * 6 + [c_int64, c_int64, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "cf_lowrank_reconstruct": (c_int, [c_void_p] * 4 + [c_int64, c_int64, c_int, c_void_p]),
    "cf_lowrank_q_pack": (c_int, [c_void_p] * 3 + [c_int64, c_int64, c_int, c_void_p]),
    "cf_lowrank_q_reconstruct": (c_int, [c_void_p] * 3 + [c_int64, c_int64, c_int, c_void_p]),
    "cf_ipc_alloc": (c_int, [c_size_t, POINTER(c_void_p), c_void_p]),
    "cf_ipc_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "cf_ipc_close": (c_int, [c_void_p]),
    "cf_ipc_free": (c_int, [c_void_p]),
    "cf_p2p_put": (c_int, [c_void_p, c_size_t, c_int, _VPP, _VPP, c_void_p, c_void_p, c_void_p]),
    "cf_p2p_wait": (c_int, [c_int, _VPP, c_void_p, c_void_p, c_void_p]),
    "cf_sign_compress_put": (c_int, [c_int, c_int, c_int, _VPP, _VPP, c_int, c_int, _VPP, _VPP, c_void_p, c_void_p, c_int64, c_int64,
                                     c_void_p, c_size_t, c_void_p]),
    "cf_sign_decompress_batched_wait": (c_int, [c_int, c_int] + [_VPP] * 6 + [c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "cf_lse_merge": (c_int, [c_void_p] * 5 + [c_int64] * 4 + [c_void_p]),
    "cf_error_stats_workspace_bytes": (c_size_t, []),
    "cf_error_stats": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cf_host_scratch_bytes": (c_size_t, [c_int, c_int64, c_int64]),
    "cf_host_compress": (c_int, [c_int] + [c_void_p] * 4 + [c_int64, c_int64, c_void_p, c_size_t, c_void_p]),
    "cf_host_decompress": (c_int, [c_int] + [c_void_p] * 3 + [c_int64, c_int64, c_void_p, c_size_t, c_void_p]),
}

_lib = None


class NativeError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle.  Raises if the library is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} not found: build it with `python -m compactfusion_b200.build` "
                "(or __graft_entry__.build()).  compactfusion_b200 has no CPU fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if a declared symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().cf_last_error().decode("utf-8", "replace")
        raise NativeError(f"{what} failed (status {rc}): {msg}")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr() -> int:
    """cudaStream_t of torch's current stream on the current device (the raw getter skips building a Stream
    object: this runs once per hook call)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def require_cuda_half(t: torch.Tensor, name: str):
    assert t.dtype == torch.half, f"{name} must be FP16"
    if not t.is_cuda:
        raise NativeError(f"{name} must be a CUDA tensor: compactfusion_b200 has no CPU path")


_workspaces: dict = {}


def workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    """Per (device, stream) scratch buffer, grown geometrically, reused across calls."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream_ptr())
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes * 1.25), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def workspace_bytes(codec: int, n: int, c: int, rank: int = 0, batch: int = 1) -> int:
    return int(lib().cf_workspace_bytes(codec, n, c, rank, batch))


def ptr_array(tensors) -> ctypes.Array:
    arr = (c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr
