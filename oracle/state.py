"""CPU restatement of the residual / error-feedback state machine and of the two
exchange schedules (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Follows xfuser/compact/main.py:169-270 (compact_compress), :322-388
(compact_decompress), :390-420 (compact_all_gather) and
xfuser/compact/ring.py:184-269 (the compress / relay / decompress order of the
compressed ring; attention itself is outside the oracle).
"""
from __future__ import annotations

import torch

from . import codecs

FASTPATH_TYPES = ("binary", "int2")


class OracleCompact:
    """One rank's compact state: config + cache (utils.py:123-160)."""

    def __init__(self, residual=1, ef=True, simulate=False, fastpath=False, comp_rank=-1,
                 sparse_ratio=None, delta_decay_factor=None):
        if residual == 0:
            assert not ef
        if residual == 2:
            assert ef
        if fastpath:
            assert ef and not simulate and residual == 1  # utils.py:87-90
        self.residual, self.ef, self.simulate, self.fastpath = residual, ef, simulate, fastpath
        self.comp_rank, self.sparse_ratio, self.decay = comp_rank, sparse_ratio, delta_decay_factor
        self.base, self.delta_base = {}, {}

    # cache ------------------------------------------------------------------
    def put(self, key, base, delta_base):
        self.base[key] = base
        self.delta_base[key] = delta_base

    # helpers ----------------------------------------------------------------
    @staticmethod
    def _to2d(shape):
        """main.py:180-185 / :333-342."""
        if len(shape) >= 4:
            d0 = 1
            for s in shape[:-2]:
                d0 *= s
            return (d0, shape[-2] * shape[-1])
        if len(shape) == 3:
            return (shape[0] * shape[1], shape[2])
        assert len(shape) == 2
        return tuple(shape)

    def _compress_fn(self, x, ctype):
        if self.simulate:
            return codecs.sim_compress(x, ctype, self.sparse_ratio, self.comp_rank)  # main.py:117-119
        return codecs.slowpath_compress(x, ctype, rank=self.comp_rank, sparse_ratio=self.sparse_ratio)

    def _decompress_fn(self, p, ctype, shape):
        if self.simulate:
            return p.view(shape)  # main.py:126-127
        return codecs.slowpath_decompress(p, shape, ctype, rank=self.comp_rank, sparse_ratio=self.sparse_ratio)

    # compress ---------------------------------------------------------------
    def compress(self, key, x, ctype, update_cache=False):
        """main.py:169-270.  `ctype` is the enum *value* string ("warmup", "binary", ...)."""
        orig = x.shape
        x = x.contiguous().view(self._to2d(orig))

        def cput(val, delta):
            if update_cache:
                self.put(key, val, delta)

        if ctype == "warmup":
            if self.fastpath or self.residual == 1:
                cput(x, None)
            elif self.residual == 2:
                b = self.base.get(key)
                cput(x, None if b is None else x - b)
            return x.view(orig)
        if self.fastpath:
            assert ctype in FASTPATH_TYPES
            base = self.base[key]
            fn = codecs.binary_quant if ctype == "binary" else codecs.int2_quant
            packed, u, v, nb = fn(x, base, update_cache)
            if update_cache:
                self.put(key, nb, None)
            return codecs.fastpath_payload(packed, u, v)
        if self.residual == 0:
            return self._compress_fn(x, ctype)
        if self.residual == 1:
            base = self.base[key]
            delta = x - base
            comp = self._compress_fn(delta, ctype)
            recon = base + self._decompress_fn(comp, ctype, x.shape)
            cput(recon if self.ef else x, None)  # main.py:233
            return comp
        base, db = self.base[key], self.delta_base[key]
        dd = x - base - db
        comp = self._compress_fn(dd, ctype)
        rdd = self._decompress_fn(comp, ctype, x.shape)
        cput(base + db + rdd, (db + rdd) * self.decay)  # main.py:250-256, :272-273
        return comp

    # decompress -------------------------------------------------------------
    def decompress(self, key, comp, ctype, shape, update_cache=False):
        """main.py:322-388."""
        orig = tuple(shape)
        s2 = self._to2d(orig)

        def cput(val, delta):
            if update_cache:
                self.put(key, val, delta)

        if ctype == "warmup":
            val = comp.view(s2)
            if self.fastpath or self.residual == 1:
                cput(val, None)
            elif self.residual == 2:
                b = self.base.get(key)
                cput(val, None if b is None else val - b)
            return val.view(orig)
        if self.fastpath:
            n, c = s2
            ipb = 8 if ctype == "binary" else 4
            packed, u, v = codecs.fastpath_split(comp, n, c, 1, ipb)
            fn = codecs.binary_dequant if ctype == "binary" else codecs.int2_dequant
            recon = fn(packed, u, v, self.base[key])
            cput(recon, None)
            return recon.view(orig)
        if self.residual == 0:
            return self._decompress_fn(comp, ctype, s2).view(orig)
        if self.residual == 1:
            recon = self.base[key] + self._decompress_fn(comp, ctype, s2)
            cput(recon, None)
            return recon.view(orig)
        base, db = self.base[key], self.delta_base[key]
        rdd = self._decompress_fn(comp, ctype, s2)
        recon = base + db + rdd
        cput(recon, (db + rdd) * self.decay)
        return recon.view(orig)


def all_gather_step(ranks: list[OracleCompact], tag: str, xs: list[torch.Tensor], ctype: str):
    """One compact_all_gather (main.py:390-420) emulated for all W ranks in one
    process.  Returns per-rank lists of W reconstructed tensors and the payloads."""
    w = len(ranks)
    payloads = [ranks[r].compress(f"{tag}-{r}", xs[r], ctype, update_cache=False) for r in range(w)]
    outs = []
    for r in range(w):
        outs.append([
            ranks[r].decompress(f"{tag}-{i}", payloads[i].clone(), ctype, xs[r].shape, update_cache=True)
            for i in range(w)
        ])
    return outs, payloads


def ring_step(ranks: list[OracleCompact], mod_idx: int, ks: list[torch.Tensor], ctype: str, suffix="k"):
    """The K (or V) side of one _compact_ring_fwd call (ring.py:184-269) for all W
    ranks: compress own shard with cache update, relay the *compressed* payload
    W-1 hops (send to rank+1, receive from rank-1), decompress each hop against
    the per-origin cache; hop 0 uses the raw local tensor (ring.py:197-208).
    Returns blocks[r][step] = the tensor rank r feeds attention at ring step `step`."""
    w = len(ranks)
    to_send = [ranks[r].compress(f"{mod_idx}-{r}-{suffix}", ks[r], ctype, update_cache=True) for r in range(w)]
    blocks = [[None] * w for _ in range(w)]
    for step in range(w):
        nxt = None
        if step + 1 != w:
            nxt = [to_send[(r - 1) % w] for r in range(w)]  # recv from rank-1
        for r in range(w):
            if step == 0:
                blocks[r][0] = ks[r]
            else:
                src = (r - step) % w
                blocks[r][step] = ranks[r].decompress(
                    f"{mod_idx}-{src}-{suffix}", to_send[r].clone(), ctype, ks[r].shape, update_cache=True)
        if nxt is not None:
            to_send = nxt
    return blocks
