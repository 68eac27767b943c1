"""Compressed ring attention (mirror of xfuser/compact/ring.py).

Each rank compresses its own K and V shard once (with error-feedback cache update), the
*compressed* payloads travel W-1 hops around the ring (send to rank+1, receive from rank-1,
relayed unchanged), every hop is decompressed against the per-origin cache, and the blocks
are merged with a log-sum-exp update.  Hop 0 uses the raw local K/V (ring.py:197-208).

B200-first changes: K and V payloads of a hop travel as ONE message (one isend/irecv pair
instead of two) and K+V of a hop are reconstructed in ONE batched launch.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import main as _m
from .attention import attn_forward, update_out_and_lse
from .main import compact_cache, compact_compress, compact_config, compact_decompress
from .prof import Profiler
from .utils import COMPACT_COMPRESS_TYPE

T = COMPACT_COMPRESS_TYPE


class RingComm:
    """P2P ring over a process group (restates yunchang.ring.utils.RingComm: batch_isend_irecv
    to rank+1 / from rank-1, `commit` launches, `wait` completes)."""

    def __init__(self, process_group):
        self._pg = process_group
        self._ops = []
        self._reqs = None
        self.rank = dist.get_rank(process_group)
        self.world_size = dist.get_world_size(process_group)
        self.send_rank = (self.rank + 1) % self.world_size
        self.recv_rank = (self.rank - 1) % self.world_size
        if process_group is not None:
            self.send_rank = dist.get_global_rank(process_group, self.send_rank)
            self.recv_rank = dist.get_global_rank(process_group, self.recv_rank)

    def send_recv(self, to_send: torch.Tensor, recv_tensor: torch.Tensor | None = None) -> torch.Tensor:
        res = torch.empty_like(to_send) if recv_tensor is None else recv_tensor
        self._ops.append(dist.P2POp(dist.isend, to_send, self.send_rank, group=self._pg))
        self._ops.append(dist.P2POp(dist.irecv, res, self.recv_rank, group=self._pg))
        return res

    def commit(self):
        if self._reqs is not None:
            raise RuntimeError("commit called twice")
        self._reqs = dist.batch_isend_irecv(self._ops)

    def wait(self):
        if self._reqs is None:
            raise RuntimeError("wait called before commit")
        for r in self._reqs:
            r.wait()
        self._reqs = None
        self._ops = []


_patch_fwd = None


def compact_fwd(q, k, v, dropout_p=0, softmax_scale=None, causal=True, window_size=(-1, -1), alibi_slopes=None,
                return_attn_probs=None, deterministic=False, attn_layer=None, group=None, joint_tensor_key=None,
                joint_tensor_value=None, joint_strategy="none", mod_idx=None, current_iter=None):
    """Entry point the xDiT long-context attention layer installs as `ring_attn_fn`
    (hybrid/attn_layer.py:59-64).  ring.py:36-70."""
    args = (q, k, v, dropout_p, softmax_scale, causal, window_size, alibi_slopes, return_attn_probs, deterministic,
            attn_layer, group, joint_tensor_key, joint_tensor_value, joint_strategy, mod_idx, current_iter)
    if compact_config().override_with_patch_gather_fwd:
        global _patch_fwd
        if _patch_fwd is None:
            from .patchpara.fwd import patch_gather_fwd   # (that module imports this one's package: first call)
            _patch_fwd = patch_gather_fwd
        return _patch_fwd(*args)
    return _compact_ring_fwd(*args)


def _joint_flags(joint_tensor_key, joint_tensor_value, joint_strategy, allowed):
    if joint_tensor_key is not None and joint_tensor_value is not None:
        if joint_strategy not in allowed:
            raise ValueError(f"joint_strategy: {joint_strategy} not supprted. supported joint strategy: {allowed}")
        return joint_strategy != "none"
    if joint_tensor_key is None and joint_tensor_value is None:
        return False
    raise ValueError("joint_tensor_key and joint_tensor_value should be None or not None simultaneously.")


def _decompress_kv(keys, payloads, ctype, shape):
    """K and V of one hop: one batched launch on the fastpath, else two plain calls."""
    cfg = compact_config()
    if cfg.fastpath and ctype in (T.BINARY, T.INT2) and cfg.comp_rank == -1:
        outs = _m._decompress_peers_batched(keys, payloads, ctype, _m._to_2d_shape(shape))
        return [o.view(shape) for o in outs]
    return [compact_decompress(key, p, ctype, shape, update_cache=True) for key, p in zip(keys, payloads)]


@Profiler.prof_func("compact._compact_ring_fwd")
def _compact_ring_fwd(q, k, v, dropout_p=0, softmax_scale=None, causal=True, window_size=(-1, -1),
                      alibi_slopes=None, return_attn_probs=None, deterministic=False, attn_layer=None, group=None,
                      joint_tensor_key=None, joint_tensor_value=None, joint_strategy="none", mod_idx=None,
                      current_iter=None):
    """ring.py:120-275."""
    assert alibi_slopes is None
    if softmax_scale is None:
        softmax_scale = q.shape[-1] ** (-0.5)
    is_joint = _joint_flags(joint_tensor_key, joint_tensor_value, joint_strategy, ["front", "rear"])
    comm = RingComm(group)
    W, me = comm.world_size, comm.rank
    q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    ctype = compact_config().compress_func(mod_idx, current_iter)  # same call for K and V (ring.py:180-181)
    shape = k.shape
    assert v.shape == shape

    from . import dropin
    ent = None
    if not causal and dropout_p == 0 and tuple(window_size) == (-1, -1):
        ent = dropin.hot(compact_config(), "ring", group, k, mod_idx, ctype)
    if ent is not None:
        # the ring on the persistent-buffer engine: the payload goes to every rank's slot at once, hop s consumes
        # origin (rank - s) mod W with one flag-waiting launch for K and V, the LSE merge is one fused kernel
        eng, layer = ent[0], ent[1]
        out, lse = eng.ring_forward(layer, q, k, v, ctype, softmax_scale, joint_tensor_key, joint_tensor_value,
                                    joint_strategy)
        return out, lse, None

    k_send = compact_compress(f"{mod_idx}-{me % W}-k", k, ctype, update_cache=True)
    v_send = compact_compress(f"{mod_idx}-{me % W}-v", v, ctype, update_cache=True)
    # one message per hop: [K payload | V payload]
    split = k_send.numel()
    msg = torch.cat([k_send.reshape(-1), v_send.reshape(-1)])

    out = lse = None
    for step in range(W):
        if step + 1 != W:
            nxt = comm.send_recv(msg)
            comm.commit()
        if step != 0:
            src = (me - step) % W
            k, v = _decompress_kv([f"{mod_idx}-{src}-k", f"{mod_idx}-{src}-v"],
                                  [msg[:split].view(k_send.shape), msg[split:].view(v_send.shape)], ctype, shape)
        key_to_use, value_to_use = k, v
        if is_joint and joint_strategy == "rear" and step + 1 == W:
            key_to_use = torch.cat([k, joint_tensor_key], dim=1)
            value_to_use = torch.cat([v, joint_tensor_value], dim=1)
        elif is_joint and joint_strategy == "front" and step == 0:
            key_to_use = torch.cat([joint_tensor_key, k], dim=1)
            value_to_use = torch.cat([joint_tensor_value, v], dim=1)
        if not causal or step <= me:
            block_out, block_lse = attn_forward(q, key_to_use, value_to_use, dropout_p, softmax_scale,
                                                causal=causal and step == 0, window_size=window_size)
            out, lse = update_out_and_lse(out, lse, block_out, block_lse)
        if step + 1 != W:
            with Profiler.scope("compact.ring.wait"):
                comm.wait()
            msg = nxt  # relay the compressed bytes unchanged (ring.py:268-269)

    out = out.to(q.dtype)
    lse = lse.squeeze(dim=-1).transpose(1, 2)
    if compact_config().check_cache_consistency:
        compact_cache().check_consistency(group=group)
    return out, lse, None
