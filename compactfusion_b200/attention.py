"""Attention block + LSE merge used by the ring / patch-parallel callers.

Attention is the *consumer* of the hot path, not part of it (SURVEY.md section 8a rows
a19/a20): it is a library call here -- flash-attn 2 when its extension imports and runs on
this GPU, else torch SDPA-style math with an explicit log-sum-exp.  `update_out_and_lse`
restates yunchang.ring.utils (third party, absent; `yunchang>=0.6.0` in the reference's
setup.py:35): out <- out - sigmoid(lse_b - lse) * (out - out_b); lse <- lse - logsigmoid(lse - lse_b).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

_flash_fwd = None
_flash_checked = False


def _try_flash():
    global _flash_fwd, _flash_checked
    if not _flash_checked:
        _flash_checked = True
        try:
            from flash_attn.flash_attn_interface import _flash_attn_forward
            _flash_fwd = _flash_attn_forward
        except Exception:
            _flash_fwd = None
    return _flash_fwd


def torch_attn_forward(q, k, v, softmax_scale, causal=False):
    """(b, s, h, d) layout.  Returns out (b, s_q, h, d) in q.dtype and lse (b, h, s_q) fp32."""
    qf, kf, vf = (t.transpose(1, 2).float() for t in (q, k, v))
    scores = torch.matmul(qf, kf.transpose(-1, -2)) * softmax_scale
    if causal:
        sq, sk = scores.shape[-2:]
        mask = torch.ones(sq, sk, dtype=torch.bool, device=q.device).tril(diagonal=sk - sq)
        scores = scores.masked_fill(~mask, float("-inf"))
    lse = torch.logsumexp(scores, dim=-1)
    out = torch.matmul(torch.exp(scores - lse.unsqueeze(-1)), vf)
    return out.transpose(1, 2).to(q.dtype), lse


_override = None
_fallback_warned = False


def set_attention_override(fn):
    """Replace the attention block by `fn(q, k, v, dropout_p, softmax_scale, causal, window_size) -> (out, lse)`
    (None restores the library call).  bench.py uses it to time the exchange hooks WITHOUT the attention that
    follows them; tests use it to observe what the hooks hand to attention."""
    global _override
    _override = fn


def attn_forward(q, k, v, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1)):
    """One attention block; returns (out (b,s,h,d), lse (b,h,s) fp32)."""
    global _flash_fwd, _fallback_warned
    if softmax_scale is None:
        softmax_scale = q.shape[-1] ** (-0.5)
    if _override is not None:
        return _override(q, k, v, dropout_p, softmax_scale, causal, window_size)
    fwd = _try_flash() if (q.is_cuda and q.dtype in (torch.half, torch.bfloat16)) else None
    if fwd is not None:
        try:
            res = fwd(q, k, v, dropout_p, softmax_scale, causal=causal, window_size_left=window_size[0],
                      window_size_right=window_size[1], softcap=0.0, alibi_slopes=None, return_softmax=False)
        except (RuntimeError, TypeError) as e:
            # only the two known "this wheel cannot serve this GPU / this signature" failures switch to the torch
            # path (once, loudly); anything else -- out of memory included -- is the caller's to see
            msg = str(e)
            if not (isinstance(e, TypeError) or "no kernel image" in msg or "is not supported" in msg
                    or "only supports" in msg):
                raise
            _flash_fwd = None
            if not _fallback_warned:
                _fallback_warned = True
                import warnings
                warnings.warn(f"flash-attn forward unusable here ({type(e).__name__}: {msg[:120]}); attention blocks "
                              "use the torch math path from now on (slow; dropout / windows unsupported)")
        else:
            # flash-attn >= 2.7: (out, softmax_lse, S_dmask, rng_state); <= 2.6.3: (out, q, k, v, out_padded,
            # softmax_lse, S_dmask, rng_state) (the reference branches on the version string, ring.py:236-262)
            if len(res) == 4:
                return res[0], res[1]
            if len(res) == 8:
                return res[0], res[5]
            raise RuntimeError(f"unexpected flash-attn forward return arity {len(res)}")
    if dropout_p != 0.0 or tuple(window_size) != (-1, -1):
        raise NotImplementedError("the torch attention path supports neither dropout nor sliding windows")
    return torch_attn_forward(q, k, v, softmax_scale, causal)


def update_out_and_lse(out, lse, block_out, block_lse):
    """Online-softmax merge of two attention blocks in fp32 (restated from yunchang.ring.utils).
    `block_lse` is (b, h, s); state `lse` is kept as (b, s, h, 1)."""
    block_out = block_out.to(torch.float32)
    block_lse = block_lse.transpose(-2, -1).unsqueeze(dim=-1)
    if out is None:
        return block_out, block_lse
    out = out - torch.sigmoid(block_lse - lse) * (out - block_out)
    lse = lse - F.logsigmoid(lse - block_lse)
    return out, lse


def merge_out_and_lse(out, lse, block_out, block_lse):
    """`update_out_and_lse` as ONE fused pass on the GPU (`cf_lse_merge`, csrc/cf_consumer.cu), in
    flash-attn's own layouts: state `out` (b, s, h, d) fp32 is updated in place, `lse` stays
    (b, h, s) fp32 (no transposes, no (b, s, h, 1) temporaries).  Returns the new (out, lse).
    Used by engine.RingExchangeEngine; CUDA tensors only (the library has no CPU path)."""
    from . import _native as nv
    if out is None:
        return block_out.to(torch.float32), block_lse.contiguous().to(torch.float32)
    if not out.is_cuda:
        raise nv.NativeError("merge_out_and_lse needs CUDA tensors: compactfusion_b200 has no CPU path")
    b, s, h, d = out.shape
    assert out.dtype == torch.float32 and out.is_contiguous()
    assert block_out.shape == out.shape and block_out.dtype == torch.half
    block_out = block_out.contiguous()
    block_lse = block_lse.contiguous()
    assert lse.shape == (b, h, s) and block_lse.shape == (b, h, s)
    assert lse.dtype == torch.float32 and block_lse.dtype == torch.float32 and lse.is_contiguous()
    new_lse = torch.empty_like(lse)
    rc = nv.lib().cf_lse_merge(out.data_ptr(), block_out.data_ptr(), lse.data_ptr(), block_lse.data_ptr(),
                               new_lse.data_ptr(), b, s, h, d, nv.stream_ptr())
    nv.check(rc, "cf_lse_merge")
    return out, new_lse
