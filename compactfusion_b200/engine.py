"""Persistent-buffer, batched, CUDA-graph-capturable runtime for the compressed K/V exchange.

`compact_all_gather` / `_compact_ring_fwd` (main.py, ring.py) are the drop-in API: they
allocate per call and launch per tensor, like the reference.  This module is the B200-first
way to drive the same kernels for a whole denoising step (SURVEY.md section 7 hard part 4):

  * one compress launch pair for K *and* V of a layer (batched C-ABI entry points),
  * K and V payloads travel as ONE message per collective,
  * one decompress launch reconstructs K and V of all W origins, writing straight into the
    layer's global-sequence K/V buffers -- which double as the error-feedback cache and as
    the attention input (no torch.cat, no per-call allocation),
  * every buffer is allocated once, so a whole step (all layers) can be captured in a CUDA
    graph and replayed with a single launch.

  * transport: NCCL all-gather of the payloads (`transport="nccl"`, what the reference does,
    main.py:409), or one-sided NVLink stores into peer-mapped receive slots with device-side
    flags (`transport="p2p"`): no collective on the path, so the whole step, exchange
    included, is ONE CUDA graph.  The exchange is fused into the codec kernels
    (`cf_sign_compress_put`: they store codes, scales and finally the flags straight into
    every rank's slot); `CF_FUSED_PUT=0` compresses into a send buffer and pushes it with the
    separate `cf_p2p_put` kernel (csrc/cf_p2p.cu).  `transport="auto"` tries p2p and falls back.

Results are the ones `compact_all_gather` produces (same kernels, same wire format); only
the 1-ulp freedom of the mean-scale reductions applies (batched launches split rows over a
different number of CTAs).
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.distributed as dist

from . import _native as nv
from .utils import COMPACT_COMPRESS_TYPE as T

_CODEC = {T.BINARY: nv.CODEC_BINARY, T.INT2: nv.CODEC_INT2}
_LOWRANK = (T.LOW_RANK, T.LOW_RANK_Q)   # slowpath payloads the engines also carry (slowpath.py:54-75, :152-164)


class _DevicePtr:
    """`nbytes` of device memory at a raw address (a CUDA IPC region is not a torch allocation), exposed
    through __cuda_array_interface__ so that torch can wrap it without a copy."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _device_bytes(ptr: int, nbytes: int, device) -> torch.Tensor:
    return torch.as_tensor(_DevicePtr(ptr, nbytes), device=device)


class LocalWorld:
    """W virtual ranks of ONE process on one GPU: W engines whose receive regions are plain local
    allocations "mapped" into each other by address -- exactly what a peer mapping looks like to the
    kernels (fan-out table, slot offsets, flag / count arithmetic, flag-waiting reconstruct).  A
    decompress launch spins on flags, so the driver must enqueue every rank's put of a layer before any
    rank's decompress of that layer (`exchange_all`, `ring_all`); real concurrency and NVLink are what
    the multi-process runs add.  Used by the parity tests (a W = 4 / 8 exchange against the oracle on a
    1-GPU box) and by `bench.py --virtual-ranks`."""

    def __init__(self, world: int, layers: int, n_local: int, c: int, device=None, engine_cls=None, **kw):
        self.world = world
        self.engines = []
        self._regions = {}
        cls = engine_cls or PatchGatherEngine
        for r in range(world):
            self.engines.append(cls(layers, n_local, c, device=device, transport="p2p", local_world=self, rank=r, **kw))

    def map_regions(self, ctype, total, flags_bytes, slot_bytes):
        sts = self._regions.get(ctype)
        if sts is None:
            dev = self.engines[0].device
            mem = [torch.zeros(total, dtype=torch.uint8, device=dev) for _ in range(self.world)]
            ptrs = [m.data_ptr() for m in mem]
            sts = []
            for r, e in enumerate(self.engines):
                st = {"base": ptrs[r], "peers": list(ptrs), "flags_bytes": flags_bytes, "slot_bytes": slot_bytes,
                      "total_bytes": total, "ipc": False, "mem": mem,
                      "count": torch.zeros(e.layers, dtype=torch.int32, device=dev),
                      "ticket": torch.zeros(1, dtype=torch.int32, device=dev),
                      "error": torch.zeros(1, dtype=torch.int32, device=dev)}
                e._p2p[ctype] = st
                sts.append(st)
            self._regions[ctype] = sts
        return sts

    def exchange_all(self, layer: int, ks, vs, ctype):
        """One layer of one step for all W virtual ranks (ks[r], vs[r]: rank r's shards); returns the engines'
        (global_k, global_v) pairs."""
        if ctype == T.WARMUP:
            return [e.warmup(layer, ks[r], vs[r]) for r, e in enumerate(self.engines)]
        for r, e in enumerate(self.engines):
            e.send(layer, ks[r], vs[r], ctype)
        for e in self.engines:
            e.decompress(layer, ctype)
        return [(e.global_k[layer], e.global_v[layer]) for e in self.engines]

    def ring_all(self, layer: int, ks, vs, ctype):
        """The ring engines' consumption order: own shard first (hop 0), then origin (rank - s) mod W per hop."""
        if ctype == T.WARMUP:
            return [e.warmup(layer, ks[r], vs[r]) for r, e in enumerate(self.engines)]
        for r, e in enumerate(self.engines):
            e.send(layer, ks[r], vs[r], ctype)
        for e in self.engines:
            e.decompress(layer, ctype, origins=(e.rank,))
        for s in range(1, self.world):
            for e in self.engines:
                e.hop(layer, s, ctype)
        return [(e.global_k[layer], e.global_v[layer]) for e in self.engines]


class PatchGatherEngine:
    """Compressed patch-parallel K/V all-gather for `layers` attention layers (bs = 1).

    global_k[l], global_v[l]: (W * n_local, C) fp16 -- rows [r*n_local, (r+1)*n_local) hold
    origin r's shard.  After `exchange(l, k, v, ...)` they contain what every rank feeds to
    attention (identical on all ranks) and are the bases of the next step.
    """

    def __init__(self, layers: int, n_local: int, c: int, group=None, device=None, transport: str = "nccl",
                 inputs_stable: bool = False, local_world: "LocalWorld | None" = None, rank: int | None = None,
                 comp_rank: int | None = None):
        self.layers, self.n, self.c = layers, n_local, c
        self.comp_rank = comp_rank  # rank r of the LOW_RANK / LOW_RANK_Q payloads (CompactConfig.comp_rank)
        self.group = group
        self._local = local_world
        if local_world is not None:  # virtual ranks of one process (LocalWorld): no process group
            self.world, self.rank = local_world.world, int(rank)
        else:
            self.world = dist.get_world_size(group) if dist.is_initialized() else 1
            self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        W, n = self.world, n_local
        self.global_k = [torch.zeros((W * n, c), dtype=torch.half, device=self.device) for _ in range(layers)]
        self.global_v = [torch.zeros((W * n, c), dtype=torch.half, device=self.device) for _ in range(layers)]
        self._payload_numel = {}
        self._send = {}
        self._recv = {}
        self._ptr_cache = {}
        self._plans = {}   # ctype -> [plan per layer] (bound call sequences of `exchange`, see _build_plan)
        self.kernel_launches = 0  # launches of our kernels since the last reset
        nv.lib()  # fail loudly now if the extension is missing
        assert transport in ("nccl", "p2p", "auto")
        self.transport = "nccl"
        self._p2p = {}
        self._p2p_requested = transport
        # layers == 0: the layer count is discovered while the warm-up step(s) walk the model (`ensure_layer`, what
        # hybrid/attn_layer.py:176-179 does for mod_idx) and frozen by the first compressed exchange (`freeze`)
        self._frozen = layers > 0
        if self._frozen:
            self._pick_transport()
        # CF_FLAG_INPUTS_STABLE (include/compactb200.h) is a CALLER promise: the kernel launched right before a
        # compress / decompress on this stream did not write the K/V inputs or the cached bases that call reads,
        # so their first tiles may be fetched while that kernel drains.  True for a runtime that walks distinct
        # per-layer buffers with static inputs (bench.py passes inputs_stable=True); NOT true in a model, where
        # the projection / RoPE kernel that produces K and V is launched right before -- hence off by default.
        self._inputs_stable = bool(inputs_stable)
        self._flags = nv.FLAG_INPUTS_STABLE if (inputs_stable and layers >= 2) else 0
        # fused compress + put (cf_sign_compress_put): the codec kernels store the payload straight into every
        # rank's receive slot; CF_FUSED_PUT=0 keeps the separate put kernel (A/B)
        self.fused_put = os.environ.get("CF_FUSED_PUT", "1") != "0"
        self._publish_kernel = os.environ.get("CF_PUBLISH_MODE", "2") == "2"  # flags published by k_publish_flags
        # per-layer pointer-keyed graphs for `exchange` (the hooks' path); the whole-step graph of `capture_step`
        # does not need them
        # (opt-in, CF_LAYER_GRAPHS=1: measured SLOWER than eager launches on a B200 -- 3.46 vs 2.50 ms per FLUX step
        #  through the hooks at N = 1, profiles/r2_bench_n1_dropin_layer_graphs.json: a graph per layer gives up the
        #  programmatic-dependent-launch overlap between consecutive layers' kernels, which is worth more than the
        #  three launches it saves)
        self._layer_graphs = os.environ.get("CF_LAYER_GRAPHS", "0") == "1"
        self._per_layer_send = False
        self._side = None  # second stream of the overlapped step

    def _pick_transport(self):
        # (with a single layer a fast rank could overwrite a slot its peer is still reading: the
        #  per-layer slots are only hazard-free when another layer's exchange separates two uses)
        if self._p2p_requested in ("p2p", "auto") and self.world > 1 and self.layers >= 2:
            self.transport = "p2p"  # regions are mapped lazily per codec (payload size differs)

    def ensure_layer(self, layer: int):
        """Grow the per-layer buffers up to `layer` (lazy engines only; frozen once compression starts, because
        the receive regions of the one-sided transport are laid out per layer and mapped by every peer)."""
        if layer < len(self.global_k):
            return
        if self._frozen and self._p2p:
            raise nv.NativeError(f"layer {layer} first seen after the one-sided transport was mapped for "
                                 f"{self.layers} layers: every attention layer must run in the warm-up step")
        W = self.world
        while len(self.global_k) <= layer:
            self.global_k.append(torch.zeros((W * self.n, self.c), dtype=torch.half, device=self.device))
            self.global_v.append(torch.zeros((W * self.n, self.c), dtype=torch.half, device=self.device))
        self.layers = len(self.global_k)

    def freeze(self):
        """First compressed exchange: the layer count is final, choose the transport."""
        if not self._frozen:
            self._frozen = True
            self._flags = nv.FLAG_INPUTS_STABLE if (self._inputs_stable and self.layers >= 2) else 0
            self._pick_transport()

    def prepare(self, ctype) -> str:
        """Set up the transport for `ctype` now (collective call); returns the transport in use."""
        if self.transport == "p2p" and (ctype in _CODEC or ctype in _LOWRANK):
            self._p2p_region(ctype)
        return self.transport

    # -- one-sided transport setup -----------------------------------------------------------
    def _p2p_region(self, ctype):
        """Allocate this rank's receive region for `ctype`, exchange CUDA IPC handles and map every
        peer's region.  Layout: [flags: layers x W u32 | slots: layers x W x (K payload | V payload)]."""
        st = self._p2p.get(ctype)
        if st is not None:
            return st
        W, L = self.world, self.layers
        slot_bytes = 2 * self._pad_bytes(ctype)   # [K payload | V payload], each padded to a 16-byte multiple
        ok = True
        flags_bytes = (L * W * 4 + 255) // 256 * 256
        total = flags_bytes + L * W * slot_bytes
        if self._local is not None:
            assert ok, "payload sizes must be multiples of 16 bytes for the one-sided transport"
            return self._local.map_regions(ctype, total, flags_bytes, slot_bytes)[self.rank]
        lib = nv.lib()
        base_ptr, handle, err = ctypes.c_void_p(), ctypes.create_string_buffer(64), None
        if ok:
            rc = lib.cf_ipc_alloc(total, ctypes.byref(base_ptr), handle)
            if rc != 0:
                ok, err = False, lib.cf_last_error().decode()
        # every rank must take the same decision
        gathered = [None] * W
        dist.all_gather_object(gathered, (ok, handle.raw if ok else b"", os.getpid()), group=self.group)
        all_ok = all(g[0] for g in gathered)
        peers = [None] * W
        if all_ok:
            for r, (_, h, _pid) in enumerate(gathered):
                if r == self.rank:
                    peers[r] = base_ptr.value
                    continue
                pp = ctypes.c_void_p()
                rc = lib.cf_ipc_open(h, ctypes.byref(pp))
                if rc != 0:
                    all_ok, err = False, lib.cf_last_error().decode()
                    break
                peers[r] = pp.value
        flag = torch.tensor([1 if all_ok else 0], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            # nothing of a half-built transport may stay mapped / allocated
            for r, pp in enumerate(peers):
                if pp is not None and r != self.rank:
                    lib.cf_ipc_close(pp)
            if base_ptr.value:
                lib.cf_ipc_free(base_ptr)
            if self._p2p_requested == "p2p":
                raise nv.NativeError(f"p2p transport unavailable on rank {self.rank}: {err or 'a peer failed'}")
            self.transport = "nccl"
            self._p2p[ctype] = False
            return False
        st = {
            "base": base_ptr.value, "peers": peers, "flags_bytes": flags_bytes, "slot_bytes": slot_bytes,
            "total_bytes": total, "ipc": True,
            "count": torch.zeros(L, dtype=torch.int32, device=self.device),     # puts issued per layer slot
            "ticket": torch.zeros(1, dtype=torch.int32, device=self.device),
            "error": torch.zeros(1, dtype=torch.int32, device=self.device),
        }
        self._p2p[ctype] = st
        dist.barrier(group=self.group)
        return st

    def close(self):
        """Release the one-sided transport: unmap every peer region (cf_ipc_close) and free the own one
        (cf_ipc_free).  Collective in spirit -- call it on all ranks once no rank touches the slots any more
        (after a barrier).  Idempotent; the engine falls back to lazily re-creating regions if used again."""
        torch.cuda.synchronize(self.device)
        lib = nv.lib()
        for st in self._p2p.values():
            if not st or not st.get("ipc"):
                continue
            for r, pp in enumerate(st["peers"]):
                if pp is not None and r != self.rank:
                    lib.cf_ipc_close(pp)
            lib.cf_ipc_free(st["base"])
        self._p2p.clear()
        self._drop_call_caches()

    def slot_bytes(self, layer: int, origin: int, ctype, kv: int = 0) -> torch.Tensor:
        """The wire payload [codes | U | V] of tensor `kv` (0: K, 1: V) that `origin` delivered for `layer`, as it
        sits in THIS rank's receive memory (one-sided transport: the slot peers write over NVLink; otherwise
        the gather / send buffer), as a uint8 tensor view -- what parity checks decode with the oracle."""
        pn_bytes = self._numel(ctype) * 2
        st = self._p2p.get(ctype) if (self.transport == "p2p" and self.world > 1) else None
        if st:
            return _device_bytes(self._slot(st, st["base"], layer, origin) + kv * self._pad_bytes(ctype), pn_bytes,
                                 self.device)
        _, recv = self._buffers(ctype, layer)
        return recv[origin, kv].view(torch.uint8)[:pn_bytes]

    def _slot(self, st, region_base, layer, origin):
        return region_base + st["flags_bytes"] + (layer * self.world + origin) * st["slot_bytes"]

    def _flag(self, st, region_base, layer, origin):
        return region_base + (layer * self.world + origin) * 4

    def p2p_error(self) -> bool:
        """True if a device-side flag wait timed out (a peer never delivered its payload)."""
        return any(bool(st["error"].item()) for st in self._p2p.values() if st)

    # -- buffers ---------------------------------------------------------------------------
    def _pad_bytes(self, ctype):
        """Bytes one tensor's payload occupies in a send buffer / receive slot: the wire size rounded up to 16, so
        that K's and V's payloads both start 16-byte aligned (bulk copies) whatever the shard length."""
        return (self._numel(ctype) * 2 + 15) // 16 * 16

    def _numel(self, ctype):
        """fp16 elements of one tensor's wire payload."""
        if ctype in _LOWRANK:
            r = self.comp_rank
            assert r is not None and 1 <= r <= 64, "LOW_RANK payloads need comp_rank in [1, 64]"
            if ctype == T.LOW_RANK:
                return r * (self.n + self.c)                      # [U (N,r) | V (r,C)]
            assert self.n % 2 == 0 and self.c % 2 == 0, "LOW_RANK_Q packs row pairs of U and of V^T"
            return r * (self.n + self.c) // 4 + 4 * r             # [qU, sU, mU, qV^T, sV, mV]
        per_byte = 8 if ctype == T.BINARY else 4
        return self.n * (self.c // per_byte) // 2 + self.n + self.c

    def _buffers(self, ctype, layer: int = 0):
        """(send, recv) payload buffers.  One pair serves all layers, except on a single GPU in overlap mode
        (`_per_layer_send`): there recv aliases send, and the reconstruct of layer l runs while layer l+1 is
        being compressed, so every layer gets its own buffer."""
        key = (ctype, layer) if self._per_layer_send else ctype
        if key not in self._send:
            self._payload_numel[ctype] = self._numel(ctype)
            pn = self._pad_bytes(ctype) // 2   # padded: the buffers mirror the receive slots' layout
            self._send[key] = torch.empty((2, pn), dtype=torch.half, device=self.device)
            self._recv[key] = (self._send[key].view(1, 2, pn) if self.world == 1 else
                               torch.empty((self.world, 2, pn), dtype=torch.half, device=self.device))
        return self._send[key], self._recv[key]

    def _views(self, flat, ctype):
        per_byte = 8 if ctype == T.BINARY else 4
        qh = self.n * (self.c // per_byte) // 2
        return flat[:qh], flat[qh:qh + self.n], flat[qh + self.n:qh + self.n + self.c]

    def _shard(self, buf, r):
        return buf[r * self.n:(r + 1) * self.n]

    # -- one layer -------------------------------------------------------------------------
    def warmup(self, layer: int, k: torch.Tensor, v: torch.Tensor):
        """Uncompressed step (COMPACT_COMPRESS_TYPE.WARMUP): raw shards are gathered straight
        into the global buffers, which become the first bases (main.py:195-209, :351-366)."""
        k2, v2 = k.reshape(self.n, self.c), v.reshape(self.n, self.c)
        if not self._frozen:
            self.ensure_layer(layer)
        if self.world == 1:
            self.global_k[layer].copy_(k2)
            self.global_v[layer].copy_(v2)
        elif self._local is not None:  # virtual ranks: the "all-gather" is W local copies
            for e in self._local.engines:
                e._shard(e.global_k[layer], self.rank).copy_(k2)
                e._shard(e.global_v[layer], self.rank).copy_(v2)
        else:
            dist.all_gather_into_tensor(self.global_k[layer], k2, group=self.group)
            dist.all_gather_into_tensor(self.global_v[layer], v2, group=self.group)
        return self.global_k[layer], self.global_v[layer]

    def _compress_args(self, layer, k2, v2, ctype):
        # cached per (layer, codec); the input pointers are patched in place every call (a model hands over
        # freshly allocated K / V tensors each step: keying the cache on their addresses would grow it forever)
        key = ("c", layer, ctype)
        args = self._ptr_cache.get(key)
        if args is None:
            send, _ = self._buffers(ctype, layer)
            pk, uk, vk = self._views(send[0], ctype)
            pv, uv, vv = self._views(send[1], ctype)
            bases = [self._shard(self.global_k[layer], self.rank), self._shard(self.global_v[layer], self.rank)]
            ws_bytes = nv.workspace_bytes(_CODEC[ctype], self.n, self.c, 0, 2)
            ws = nv.workspace(ws_bytes, self.device)
            args = (nv.ptr_array([k2, v2]), nv.ptr_array(bases), nv.ptr_array([None, None]), nv.ptr_array([pk, pv]),
                    nv.ptr_array([uk, uv]), nv.ptr_array([vk, vv]), ws)
            self._ptr_cache[key] = args
        args[0][0], args[0][1] = k2.data_ptr(), v2.data_ptr()
        return args

    def _decompress_args_p2p(self, layer, ctype, st, origins=None):
        origins = tuple(range(self.world)) if origins is None else tuple(origins)
        key = ("dp", layer, ctype, origins)
        args = self._ptr_cache.get(key)
        if args is None:
            per_byte = 8 if ctype == T.BINARY else 4
            code_bytes = self.n * (self.c // per_byte)
            pn_bytes = self._pad_bytes(ctype)
            packed, us, vs, bases, flags = [], [], [], [], []
            for r in origins:
                slot = self._slot(st, st["base"], layer, r)
                for j, glob in enumerate((self.global_k[layer], self.global_v[layer])):
                    p0 = slot + j * pn_bytes
                    packed.append(p0)
                    us.append(p0 + code_bytes)
                    vs.append(p0 + code_bytes + 2 * self.n)
                    bases.append(self._shard(glob, r).data_ptr())
                    flags.append(self._flag(st, st["base"], layer, r))
            chunks = []
            for s0 in range(0, len(packed), nv.CF_MAX_BATCH):
                e0 = min(len(packed), s0 + nv.CF_MAX_BATCH)
                arr = lambda xs: (ctypes.c_void_p * (e0 - s0))(*xs[s0:e0])  # noqa: E731
                b = arr(bases)
                chunks.append((e0 - s0, arr(packed), arr(us), arr(vs), b, b, arr(flags)))
            args = chunks
            self._ptr_cache[key] = args
        return args

    def _decompress_args(self, layer, ctype, origins=None):
        origins = tuple(range(self.world)) if origins is None else tuple(origins)
        key = ("d", layer, ctype, origins)
        args = self._ptr_cache.get(key)
        if args is None:
            _, recv = self._buffers(ctype, layer)
            packed, us, vs, bases = [], [], [], []
            for r in origins:
                for j, glob in enumerate((self.global_k[layer], self.global_v[layer])):
                    p, u, v = self._views(recv[r, j], ctype)
                    packed.append(p)
                    us.append(u)
                    vs.append(v)
                    bases.append(self._shard(glob, r))
            chunks = []
            for s in range(0, len(packed), nv.CF_MAX_BATCH):
                e = min(len(packed), s + nv.CF_MAX_BATCH)
                b = nv.ptr_array(bases[s:e])
                chunks.append((e - s, nv.ptr_array(packed[s:e]), nv.ptr_array(us[s:e]), nv.ptr_array(vs[s:e]), b, b))
            args = chunks
            self._ptr_cache[key] = args
        return args

    def compress(self, layer, k, v, ctype, passes: int = nv.PASS_ALL):
        """K and V of this rank -> the send buffer (no cache update: the sender's own shard is
        updated by the decompress below, like every other origin; main.py:400-405).
        `passes` selects individual kernels of the call (bench.py times them one by one)."""
        k2, v2 = k.reshape(self.n, self.c), v.reshape(self.n, self.c)
        xs, bases, nones, pk, us, vs, ws = self._compress_args(layer, k2, v2, ctype)
        if passes == 0:  # argument arrays only (exchange's bound plan)
            return
        rc = nv.lib().cf_sign_compress_passes(_CODEC[ctype] | self._flags, passes, 2, xs, bases, nones, pk, us, vs, self.n, self.c,
                                              ws.data_ptr(), ws.numel(), nv.stream_ptr())
        nv.check(rc, "cf_sign_compress_passes")
        if ctype == T.BINARY:
            passes &= ~nv.PASS_ENCODE  # no cache update on the sender: BINARY has no third kernel
        self.kernel_launches += bin(passes).count("1")

    def fused(self, ctype) -> bool:
        """True if compress and exchange of `ctype` run as one fused call (no send buffer, no put kernel)."""
        return self.fused_put and self.transport == "p2p" and self.world > 1 and ctype in _CODEC

    def compress_put(self, layer, k, v, ctype, passes: int = nv.PASS_ALL):
        """Fused compress + one-sided exchange: the kernels write K's and V's payload into slot
        (layer, this rank) of every rank's receive region (own included) and publish the flags."""
        st = self._p2p_region(ctype)
        assert st, "compress_put needs the p2p transport"
        k2, v2 = k.reshape(self.n, self.c), v.reshape(self.n, self.c)
        key = ("cp", layer, ctype)
        args = self._ptr_cache.get(key)
        if args is None:
            W, pn_bytes = self.world, self._pad_bytes(ctype)
            bases = [self._shard(self.global_k[layer], self.rank), self._shard(self.global_v[layer], self.rank)]
            dst = (ctypes.c_void_p * (2 * W))(*[self._slot(st, st["peers"][q], layer, self.rank) + j * pn_bytes
                                               for j in range(2) for q in range(W)])
            flg = (ctypes.c_void_p * W)(*[self._flag(st, st["peers"][q], layer, self.rank) for q in range(W)])
            ws = nv.workspace(nv.workspace_bytes(_CODEC[ctype], self.n, self.c, 0, 2), self.device)
            args = (nv.ptr_array([k2, v2]), nv.ptr_array(bases), dst, flg, st["count"][layer:layer + 1].data_ptr(), ws)
            self._ptr_cache[key] = args
        xs, bases, dst, flg, count_ptr, ws = args
        xs[0], xs[1] = k2.data_ptr(), v2.data_ptr()
        if passes == 0:  # argument arrays only (exchange's bound plan)
            return
        rc = nv.lib().cf_sign_compress_put(_CODEC[ctype] | self._flags, passes, 2, xs, bases, self.world, self.rank, dst, flg, count_ptr,
                                           st["ticket"].data_ptr(), self.n, self.c, ws.data_ptr(), ws.numel(),
                                           nv.stream_ptr())
        nv.check(rc, "cf_sign_compress_put")
        last = nv.PASS_FINALIZE if ctype == T.BINARY else nv.PASS_ENCODE
        publish = 1 if (passes & last and self._publish_kernel) else 0  # k_publish_flags
        if ctype == T.BINARY:
            passes &= ~nv.PASS_ENCODE
        self.kernel_launches += bin(passes).count("1") + publish

    def gather(self, ctype, layer: int = 0):
        """Move this rank's [K payload | V payload] to every rank: NCCL all-gather, or one put
        kernel writing all W receive slots (own included) over NVLink + flag publication."""
        send, recv = self._buffers(ctype, layer)
        if self.world == 1:
            return
        st = self._p2p_region(ctype) if self.transport == "p2p" else None
        if not st:
            dist.all_gather_into_tensor(recv.view(self.world, -1), send.view(-1), group=self.group)
            return
        key = ("put", layer, ctype)
        args = self._ptr_cache.get(key)
        if args is None:
            W = self.world
            dst = (ctypes.c_void_p * W)(*[self._slot(st, st["peers"][q], layer, self.rank) for q in range(W)])
            flg = (ctypes.c_void_p * W)(*[self._flag(st, st["peers"][q], layer, self.rank) for q in range(W)])
            args = (dst, flg, st["count"][layer:layer + 1].data_ptr())
            self._ptr_cache[key] = args
        dst, flg, count_ptr = args
        rc = nv.lib().cf_p2p_put(send.data_ptr(), st["slot_bytes"], self.world, dst, flg, count_ptr,
                                 st["ticket"].data_ptr(), nv.stream_ptr())
        nv.check(rc, "cf_p2p_put")
        self.kernel_launches += 1

    def decompress(self, layer, ctype, origins=None):
        """Origins x {K, V} (default: all W): recon = base + dequant, in place in the global buffers.
        With the one-sided transport the kernel first waits for the flags of exactly these origins."""
        if ctype in _LOWRANK:
            return self._lowrank_decompress(layer, ctype, origins)
        st = self._p2p.get(ctype) if (self.transport == "p2p" and self.world > 1) else None
        if st:
            expected = st["count"][layer:layer + 1].data_ptr()
            for cnt, pk, us, vs, bases, recon, flags in self._decompress_args_p2p(layer, ctype, st, origins):
                rc = nv.lib().cf_sign_decompress_batched_wait(_CODEC[ctype] | self._flags, cnt, pk, us, vs, bases, recon, flags,
                                                              expected, st["error"].data_ptr(), self.n, self.c,
                                                              nv.stream_ptr())
                nv.check(rc, "cf_sign_decompress_batched_wait")
                self.kernel_launches += 1
            return
        for cnt, pk, us, vs, bases, recon in self._decompress_args(layer, ctype, origins):
            rc = nv.lib().cf_sign_decompress_batched_wait(_CODEC[ctype] | self._flags, cnt, pk, us, vs, bases, recon,
                                                          None, None, None, self.n, self.c, nv.stream_ptr())
            nv.check(rc, "cf_sign_decompress_batched_wait")
            self.kernel_launches += 1

    # -- LOW_RANK / LOW_RANK_Q payloads (the CogVideoX preset, examples/configs.py:87-97) ---------------------
    def _lowrank_compress(self, layer, k, v, ctype):
        """K and V -> [U | V] (LOW_RANK) or [qU, sU, mU, qV^T, sV, mV] (LOW_RANK_Q: int4 per column of U and of
        V^T, slowpath.py:62-75) in the send buffer; the projector subtracts the cached base on the fly."""
        from .compress_lowrank import lowrank_project, lowrank_q_pack
        n, c, r = self.n, self.c, self.comp_rank
        send, _ = self._buffers(ctype, layer)
        for j, (x, glob) in enumerate(((k, self.global_k[layer]), (v, self.global_v[layer]))):
            x2, base, payload = x.reshape(n, c), self._shard(glob, self.rank), send[j]
            if ctype == T.LOW_RANK:
                lowrank_project(x2, base, r, 2, u_out=payload[:n * r].view(n, r),
                                v_out=payload[n * r:n * r + r * c].view(r, c))
            else:
                u, vv, _ = lowrank_project(x2, base, r, 2)
                lowrank_q_pack(u, vv, out=payload)
            self.kernel_launches += 1

    def _lowrank_payload(self, layer, origin, ctype, kv):
        """fp16 view of the payload tensor `kv` that `origin` delivered for `layer`."""
        return self.slot_bytes(layer, origin, ctype, kv).view(torch.half)

    def _lowrank_decompress(self, layer, ctype, origins=None):
        """recon = base + U V for the given origins, in place in the global buffers; with the one-sided transport a
        one-warp kernel (cf_p2p_wait) first holds the stream until those origins' flags are up."""
        from .compress_lowrank import lowrank_q_reconstruct, lowrank_reconstruct
        origins = tuple(range(self.world)) if origins is None else tuple(origins)
        n, c, r = self.n, self.c, self.comp_rank
        st = self._p2p.get(ctype) if (self.transport == "p2p" and self.world > 1) else None
        if st:
            flags = (ctypes.c_void_p * len(origins))(*[self._flag(st, st["base"], layer, o) for o in origins])
            rc = nv.lib().cf_p2p_wait(len(origins), flags, st["count"][layer:layer + 1].data_ptr(),
                                      st["error"].data_ptr(), nv.stream_ptr())
            nv.check(rc, "cf_p2p_wait")
            self.kernel_launches += 1
        for o in origins:
            for j, glob in enumerate((self.global_k[layer], self.global_v[layer])):
                payload, shard = self._lowrank_payload(layer, o, ctype, j), self._shard(glob, o)
                if ctype == T.LOW_RANK:
                    lowrank_reconstruct(payload[:n * r].view(n, r), payload[n * r:].view(r, c), base=shard, out=shard)
                else:   # int4 decode of both factors inside the reconstruct kernel
                    lowrank_q_reconstruct(payload, n, c, r, base=shard, out=shard)
                self.kernel_launches += 1

    def send(self, layer: int, k: torch.Tensor, v: torch.Tensor, ctype):
        """Sender side of one layer: compress this rank's K and V and deliver the payloads to every rank."""
        assert ctype in _CODEC or ctype in _LOWRANK, f"engine supports BINARY / INT2 / LOW_RANK / LOW_RANK_Q, got {ctype}"
        self.freeze()
        if ctype in _LOWRANK:
            self._lowrank_compress(layer, k, v, ctype)
            self.gather(ctype, layer)
            return
        if self.fused(ctype):
            self.compress_put(layer, k, v, ctype)
        else:
            self.compress(layer, k, v, ctype)
            self.gather(ctype, layer)

    def _drop_call_caches(self):
        self._ptr_cache.clear()
        self._plans.clear()

    def exchange(self, layer: int, k: torch.Tensor, v: torch.Tensor, ctype):
        """One layer of one step; returns (global_k, global_v) ready for attention."""
        if ctype is T.WARMUP:
            return self.warmup(layer, k, v)
        plans = self._plans.get(ctype)
        plan = plans[layer] if (plans is not None and layer < len(plans)) else None
        if plan is None:
            plan = self._build_plan(layer, k, v, ctype)
        plan(k, v)
        return self.global_k[layer], self.global_v[layer]

    def _build_plan(self, layer, k, v, ctype):
        """The per-(layer, codec) call sequence of `exchange` with every argument bound once: the hooks run it once
        per attention layer per step with eager launches, so the host side of a layer is two C calls and a handful
        of Python operations.  (Rebuilt whenever `_ptr_cache` is cleared: new stream, graph capture.)"""
        self.freeze()
        generic = ctype in _LOWRANK or type(self).decompress is not PatchGatherEngine.decompress
        if not generic and self.fused(ctype):
            self.compress_put(layer, k, v, ctype, passes=0)          # builds and caches the argument arrays
            xs, bases, dst, flg, count_ptr, ws = self._ptr_cache[("cp", layer, ctype)]
            st = self._p2p[ctype]
            chunks = self._decompress_args_p2p(layer, ctype, st)
            lib, codec = nv.lib(), _CODEC[ctype] | self._flags
            put, dec = lib.cf_sign_compress_put, lib.cf_sign_decompress_batched_wait
            W, rank, n, c = self.world, self.rank, self.n, self.c
            ticket, err, ws_ptr, ws_n = st["ticket"].data_ptr(), st["error"].data_ptr(), ws.data_ptr(), ws.numel()
            n_launch = (3 if ctype == T.BINARY else 4) + len(chunks) - (0 if self._publish_kernel else 1)
            stream_ptr = nv.stream_ptr

            def plan(k_, v_):
                xs[0], xs[1] = k_.data_ptr(), v_.data_ptr()
                sp = stream_ptr()
                rc = put(codec, nv.PASS_ALL, 2, xs, bases, W, rank, dst, flg, count_ptr, ticket, n, c, ws_ptr, ws_n, sp)
                if rc:
                    nv.check(rc, "cf_sign_compress_put")
                for cnt, pk, us, vs, bs, recon, flags in chunks:
                    rc = dec(codec, cnt, pk, us, vs, bs, recon, flags, count_ptr, err, n, c, sp)
                    if rc:
                        nv.check(rc, "cf_sign_decompress_batched_wait")
                self.kernel_launches += n_launch
        elif not generic and self.world == 1:
            self.compress(layer, k, v, ctype, passes=0)
            xs, bases, nones, pk, us, vs, ws = self._ptr_cache[("c", layer, ctype)]
            chunks = self._decompress_args(layer, ctype)
            lib, codec = nv.lib(), _CODEC[ctype] | self._flags
            comp, dec = lib.cf_sign_compress_passes, lib.cf_sign_decompress_batched_wait
            n, c, ws_ptr, ws_n = self.n, self.c, ws.data_ptr(), ws.numel()
            n_launch = (2 if ctype == T.BINARY else 3) + len(chunks)
            stream_ptr = nv.stream_ptr

            def plan(k_, v_):
                xs[0], xs[1] = k_.data_ptr(), v_.data_ptr()
                sp = stream_ptr()
                rc = comp(codec, nv.PASS_ALL, 2, xs, bases, nones, pk, us, vs, n, c, ws_ptr, ws_n, sp)
                if rc:
                    nv.check(rc, "cf_sign_compress_passes")
                for cnt, dpk, dus, dvs, bs, recon in chunks:
                    rc = dec(codec, cnt, dpk, dus, dvs, bs, recon, None, None, None, n, c, sp)
                    if rc:
                        nv.check(rc, "cf_sign_decompress_batched_wait")
                self.kernel_launches += n_launch
        else:
            def plan(k_, v_):
                self.send(layer, k_, v_, ctype)
                self.decompress(layer, ctype)
        if not generic and self._layer_graphs:
            plan = self._graphed(plan)
        if os.environ.get("CF_PLAN_TIMING") == "1":   # host time spent inside the plan (the C calls), for tools
            import time
            inner, acc = plan, self.__dict__.setdefault("plan_host_s", [0.0, 0])

            def plan(k_, v_):  # noqa: F811
                t0 = time.perf_counter()
                inner(k_, v_)
                acc[0] += time.perf_counter() - t0
                acc[1] += 1
        plans = self._plans.setdefault(ctype, [])
        plans.extend([None] * (layer + 1 - len(plans)))
        plans[layer] = plan
        return plan

    def _graphed(self, eager):
        """Pointer-keyed CUDA graphs of one layer's launches.  A model hands over K / V tensors from the caching
        allocator, whose addresses repeat from step to step (or rotate among a few): when a layer sees the same pair
        of addresses a second time, its eager launch sequence is captured once for that pair and replayed from then
        on (one launch instead of four or five); an unknown pair runs the eager path.  At most 8 graphs per layer.
        Opt-in (`CF_LAYER_GRAPHS=1`): see the measurement note in __init__."""
        seen, graphs, launches = {}, {}, {}
        state = {"ok": True}

        def plan(k_, v_):
            ptrs = (k_.data_ptr(), v_.data_ptr())
            g = graphs.get(ptrs)
            if g is not None:
                g.replay()
                self.kernel_launches += launches[ptrs]
                return
            if state["ok"]:
                n_seen = seen.get(ptrs, 0) + 1
                if len(seen) > 64:
                    seen.clear()
                seen[ptrs] = n_seen
                if n_seen >= 2 and len(graphs) < 8:
                    try:
                        if torch.cuda.is_current_stream_capturing():
                            raise RuntimeError("already inside a capture (the whole-step graph)")
                        before = self.kernel_launches
                        g = torch.cuda.CUDAGraph()
                        cur = torch.cuda.current_stream()
                        side = torch.cuda.Stream(device=self.device)
                        side.wait_stream(cur)
                        with torch.cuda.stream(side):
                            g.capture_begin(capture_error_mode="thread_local")
                            try:
                                eager(k_, v_)
                            finally:
                                g.capture_end()
                        cur.wait_stream(side)
                        launches[ptrs] = self.kernel_launches - before   # kernels per replay (this call: below)
                        graphs[ptrs] = g
                        g.replay()
                        return
                    except Exception:  # noqa: BLE001 -- a capture that cannot be taken leaves the eager path in place
                        state["ok"] = False
            eager(k_, v_)
        return plan

    def check_errors(self):
        """Raise if a device-side flag wait timed out (a peer never delivered; the affected reconstructions were
        skipped, so that origin's cache is one step stale on this rank).  Synchronises: call it at step end."""
        if self.p2p_error():
            raise nv.NativeError(f"rank {self.rank}: a device-side wait for a peer's payload timed out (~2 s); "
                                 "the caches of the ranks have diverged -- reset the plugin state")

    # -- whole step ------------------------------------------------------------------------
    OVERLAP_LAG = 2  # layers the compress chain may run ahead of the reconstruct chain

    def can_overlap(self, ctype) -> bool:
        """The two-chain step needs hazard-free buffers between the chains: per-layer receive slots (one-sided
        transport) or per-layer send buffers (single GPU); the NCCL transport gathers every layer into one
        buffer.  The slot-reuse argument below needs layers >= 2 * OVERLAP_LAG + 1."""
        if ctype not in _CODEC or self.layers < 2 * self.OVERLAP_LAG + 1 or type(self).exchange is not PatchGatherEngine.exchange:
            return False
        return self.world == 1 or self.transport == "p2p"

    def step(self, ks, vs, ctype, overlap: bool = False):
        if overlap and self.can_overlap(ctype):
            return self._step_overlapped(ks, vs, ctype)
        for layer in range(self.layers):
            self.exchange(layer, ks[layer], vs[layer], ctype)

    def _step_overlapped(self, ks, vs, ctype):
        """One step as TWO chains: compress (+ put) of every layer on the current stream, reconstruct of every
        layer on a side stream, joined per layer by an event.  The compress side of layer l+1 (a short
        streaming pass plus a finalize kernel that mostly waits: partial sums, and at W > 1 a system-scope
        fence and remote flag updates) then fills the bubbles of the bandwidth-bound reconstruct of layer l
        instead of sitting in front of it.  Same kernels, same operands, bit-identical results.

        Slot reuse stays safe at W > 1 because the compress chain is held to OVERLAP_LAG = d layers ahead
        (put(l) waits for this rank's reconstruct(l - d)): a peer A overwrites slot (l, A) here in step t+1
        only after (l >= d) its own reconstruct_{t+1}(l - d), which needed OUR put_{t+1}(l - d), issued after
        our whole step t; or (l < d) after its step t, whose reconstruct_t(L-1) needed our put_t(L-1), issued
        after our reconstruct_t(L-1-d) and hence -- the reconstruct chain runs in layer order -- after our
        reconstruct_t(l), as l < d <= L-1-d."""
        if self.world == 1 and not self._per_layer_send:
            self._per_layer_send = True
            self._drop_call_caches()
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        side, lag = self._side, self.OVERLAP_LAG
        side.wait_stream(main)  # fork (inside a capture this pulls the side stream into the graph)
        done = []
        for layer in range(self.layers):
            if layer >= lag:
                main.wait_event(done[layer - lag])
            self.send(layer, ks[layer], vs[layer], ctype)
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ready)
                self.decompress(layer, ctype)
                ev = torch.cuda.Event()
                ev.record(side)
            done.append(ev)
        main.wait_stream(side)  # join

    def capture_step(self, ks, vs, ctype, warmup_iters: int = 1, overlap: bool = False):
        """Capture `step` (all layers, fixed input buffers) into a CUDA graph; returns the graph.
        The caller replays it after refreshing ks / vs contents in place.  `overlap`: the two-chain
        step (`_step_overlapped`) where `can_overlap` allows it."""
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup_iters):
                self.step(ks, vs, ctype, overlap)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._drop_call_caches()  # workspaces are keyed by stream: rebuild inside the capture
        g = torch.cuda.CUDAGraph()
        before = self.kernel_launches
        with torch.cuda.graph(g):
            self.step(ks, vs, ctype, overlap)
        self.launches_per_graph = self.kernel_launches - before
        self._drop_call_caches()
        return g


class RingExchangeEngine(PatchGatherEngine):
    """Compressed ring attention (`_compact_ring_fwd`, ring.py:120-275) on the engine's persistent buffers.

    The reference relays each origin's compressed payload W-1 hops around a NCCL P2P ring
    (ring.py:268-269) because a relay is what a ring of point-to-point links offers.  Behind an
    NVSwitch every peer is one hop away, so here the sender's codec kernels store its payload into
    ALL W receive slots at once (`cf_sign_compress_put`, or one all-gather on the NCCL transport)
    and the "ring" is only the ORDER in which a rank consumes the origins: hop s reconstructs
    origin (rank - s) mod W -- K and V in one launch that waits, on the device, for exactly that
    origin's flag -- and hands the block to attention while later payloads are still in flight.
    Hop 0 attends to the RAW local K/V (ring.py:197-208) while the own-shard cache receives the
    error-feedback reconstruction, computed by the same kernel every receiver runs on the same
    payload, so all W copies of an origin's base stay bit-identical.
    Non-causal attention only (every hop is consumed, like the DiT callers of ring.py).
    """

    def hop_origin(self, hop: int) -> int:
        return (self.rank - hop) % self.world

    def begin(self, layer: int, k: torch.Tensor, v: torch.Tensor, ctype):
        """Hop 0: compress this rank's K and V against the cached base, publish the payload to every
        rank and update the own-shard cache (compact_compress(..., update_cache=True), ring.py:184-185)."""
        if ctype == T.WARMUP:
            self.warmup(layer, k, v)
            return
        self.send(layer, k, v, ctype)
        self.decompress(layer, ctype, origins=(self.rank,))

    def hop(self, layer: int, hop: int, ctype):
        """Hop s >= 1: reconstruct origin (rank - s) mod W (cache update, ring.py:199-200); returns
        that origin's (n_local, C) K and V blocks (views into the global buffers)."""
        r = self.hop_origin(hop)
        if ctype != T.WARMUP and hop != 0:
            self.decompress(layer, ctype, origins=(r,))
        return self._shard(self.global_k[layer], r), self._shard(self.global_v[layer], r)

    def exchange(self, layer: int, k: torch.Tensor, v: torch.Tensor, ctype):
        """The hot path of one layer without the attention blocks (what bench.py times)."""
        self.begin(layer, k, v, ctype)
        for s in range(1, self.world):
            self.hop(layer, s, ctype)
        return self.global_k[layer], self.global_v[layer]

    def ring_forward(self, layer: int, q, k, v, ctype, softmax_scale=None, joint_tensor_key=None,
                     joint_tensor_value=None, joint_strategy: str = "none"):
        """q, k, v: (bs, s_local, h, d) fp16.  Returns (out (bs, s_local, h, d) fp16, lse (bs, h, s_local) fp32),
        the values `_compact_ring_fwd(..., causal=False)` returns."""
        from .attention import attn_forward, merge_out_and_lse
        from .ring import _joint_flags
        shape = k.shape
        assert v.shape == shape and shape[0] * shape[1] == self.n and shape[2] * shape[3] == self.c
        is_joint = _joint_flags(joint_tensor_key, joint_tensor_value, joint_strategy, ["front", "rear"])
        if softmax_scale is None:
            softmax_scale = q.shape[-1] ** (-0.5)
        k, v = k.contiguous(), v.contiguous()
        self.begin(layer, k, v, ctype)
        out = lse = None
        W = self.world
        for s in range(W):
            if s == 0:
                kk, vv = k, v
            else:
                ks_, vs_ = self.hop(layer, s, ctype)
                kk, vv = ks_.view(shape), vs_.view(shape)
            if is_joint and joint_strategy == "rear" and s + 1 == W:
                kk, vv = torch.cat([kk, joint_tensor_key], dim=1), torch.cat([vv, joint_tensor_value], dim=1)
            elif is_joint and joint_strategy == "front" and s == 0:
                kk, vv = torch.cat([joint_tensor_key, kk], dim=1), torch.cat([joint_tensor_value, vv], dim=1)
            block_out, block_lse = attn_forward(q, kk, vv, 0.0, softmax_scale, causal=False)
            out, lse = merge_out_and_lse(out, lse, block_out, block_lse)
        return out.to(q.dtype), lse
