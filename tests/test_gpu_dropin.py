"""The reference's hooks (`compact_fwd` -> `patch_gather_fwd` / `_compact_ring_fwd`, hybrid/attn_layer.py:59-64) on the
persistent-buffer engines (compactfusion_b200/dropin.py) against the per-call path and the oracle; all three modes
of `patch_gather_fwd` (patchpara/fwd.py:20-237: compact, synchronous, stale-async)."""
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist

from conftest import free_port, rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def one_rank_group():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    created = False
    if not dist.is_initialized():
        # NCCL, not gloo: the per-call and the uncompressed paths all-gather CUDA tensors
        torch.cuda.set_device(0)
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{free_port()}", world_size=1, rank=0,
                                device_id=torch.device("cuda:0"))
        created = True
    yield torch.device("cuda:0")
    from compactfusion_b200 import attention, dropin
    attention.set_attention_override(None)
    dropin.shutdown()
    if created:
        dist.destroy_process_group()


def _series(bs, s, h, d, steps, layers, dev, seed):
    g = torch.Generator().manual_seed(seed)
    x0 = [[torch.randn(bs, s, h, d, generator=g) for _ in range(3)] for _ in range(layers)]
    return [[[(0.97 ** t * x0[l][j] + 0.2 * torch.randn(bs, s, h, d, generator=g)).half().to(dev) for j in range(3)]
             for l in range(layers)] for t in range(steps)]  # [t][layer][q|k|v]


def _run_hooks(cfg_kw, data, engine: bool, monkeypatch, patch: bool):
    """Drive compact_fwd over steps x layers; returns per call (key, value) attention saw and (out, lse)."""
    import compactfusion_b200 as cf
    from compactfusion_b200 import attention, dropin
    from compactfusion_b200.attention import attn_forward
    T = cf.COMPACT_COMPRESS_TYPE
    monkeypatch.setenv("CF_DROPIN_ENGINE", "1" if engine else "0")
    codec = cfg_kw.pop("codec")
    kw = dict(enabled=True, compress_func=lambda l, s: codec if s >= 1 else T.WARMUP, comp_rank=-1, residual=1, ef=True,
              fastpath=True, **cfg_kw)
    cf.compact_init(cf.CompactConfig(**kw))
    seen, outs = [], []

    def spy(q, k, v, dropout_p, scale, causal, window):
        seen.append((k.clone(), v.clone()))
        attention.set_attention_override(None)
        try:
            return attn_forward(q, k, v, dropout_p, scale, causal, window)
        finally:
            attention.set_attention_override(spy)

    attention.set_attention_override(spy)
    try:
        for t, step in enumerate(data):
            cf.compact_set_step(t)
            for l, (q, k, v) in enumerate(step):
                out, lse, _ = cf.compact_fwd(q, k, v, causal=False, mod_idx=l, current_iter=t)
                outs.append((out.clone(), lse.clone()))
    finally:
        attention.set_attention_override(None)
    n_engines = len(dropin.engines())
    torch.cuda.synchronize()
    return seen, outs, n_engines


@pytest.mark.parametrize("codec", ["binary", "int2"])
@pytest.mark.parametrize("bs,s,h,d", [(1, 576, 24, 128), (2, 136, 16, 72)])
def test_patch_gather_hook_compact_mode_engine_vs_per_call_vs_oracle(one_rank_group, monkeypatch, codec, bs, s, h, d):
    """patch_gather_fwd, compact mode: the hook on the engine hands attention the same K / V (up to the 1-ulp scale
    freedom of the batched reductions) as the per-call compact_all_gather path, both track the oracle's
    `all_gather_step`, and the attention output follows."""
    dev = one_rank_group
    import compactfusion_b200 as cf
    from oracle.state import OracleCompact, all_gather_step
    T = cf.COMPACT_COMPRESS_TYPE
    steps, layers = 4, 3
    data = _series(bs, s, h, d, steps, layers, dev, seed=s + h)
    kw = lambda: dict(codec=T(codec), override_with_patch_gather_fwd=True,  # noqa: E731
                      patch_gather_fwd_config=cf.PatchConfig(True, False, 1))
    seen_e, outs_e, n_e = _run_hooks(kw(), data, True, monkeypatch, True)
    seen_p, outs_p, n_p = _run_hooks(kw(), data, False, monkeypatch, True)
    assert n_e == 1 and n_p == 0, "the engine must serve the fast configuration, and only when enabled"
    oracle = [OracleCompact(residual=1, ef=True, fastpath=True)]
    i = 0
    for t in range(steps):
        ct = codec if t >= 1 else "warmup"
        for l in range(layers):
            (ke, ve), (kp, vp) = seen_e[i], seen_p[i]
            assert ke.shape == (bs, s, h, d)
            ok, _ = all_gather_step(oracle, f"{l}-k", [data[t][l][1].cpu()], ct)
            ov, _ = all_gather_step(oracle, f"{l}-v", [data[t][l][2].cpu()], ct)
            if t == 0:
                assert torch.equal(ke, data[0][l][1]) and torch.equal(kp, data[0][l][1])
            assert rel_l2(ke, kp) < 2e-3 and rel_l2(ve, vp) < 2e-3, (t, l, rel_l2(ke, kp))
            assert rel_l2(ke, ok[0][0]) < 2e-3 and rel_l2(ve, ov[0][0]) < 2e-3, (t, l)
            assert torch.allclose(outs_e[i][0].float(), outs_p[i][0].float(), atol=5e-3)
            i += 1


@pytest.mark.parametrize("codec", ["binary", "int2"])
def test_ring_hook_engine_vs_per_call(one_rank_group, monkeypatch, codec):
    """_compact_ring_fwd on the ring engine == the per-call ring (W = 1: hop 0 attends to the RAW K / V, ring.py:197-208)."""
    dev = one_rank_group
    import compactfusion_b200 as cf
    from compactfusion_b200.attention import attn_forward
    T = cf.COMPACT_COMPRESS_TYPE
    bs, s, h, d, steps, layers = 1, 544, 24, 128, 3, 2
    data = _series(bs, s, h, d, steps, layers, dev, seed=9)
    seen_e, outs_e, n_e = _run_hooks(dict(codec=T(codec)), data, True, monkeypatch, False)
    seen_p, outs_p, n_p = _run_hooks(dict(codec=T(codec)), data, False, monkeypatch, False)
    assert n_e == 1 and n_p == 0
    i = 0
    for t in range(steps):
        for l in range(layers):
            q, k, v = data[t][l]
            assert torch.equal(seen_e[i][0], k) and torch.equal(seen_p[i][0], k)
            ref, ref_lse = attn_forward(q, k, v, 0.0, None, causal=False)
            for out, lse in (outs_e[i], outs_p[i]):
                assert out.shape == q.shape and lse.shape == (bs, h, s)
                assert torch.allclose(out.float(), ref.float(), atol=2e-3) and torch.allclose(lse, ref_lse, atol=1e-4)
            i += 1


def test_hooks_keep_the_per_call_path_for_other_configurations(one_rank_group, monkeypatch):
    """Stats logging, the consistency check or a non-fastpath codec are served by the per-call path (no engine)."""
    dev = one_rank_group
    import compactfusion_b200 as cf
    T = cf.COMPACT_COMPRESS_TYPE
    data = _series(1, 64, 8, 128, 2, 1, dev, seed=1)
    _, _, n = _run_hooks(dict(codec=T.BINARY, check_consist=True), data, True, monkeypatch, False)
    assert n == 0


@pytest.mark.parametrize("mode", ["sync", "async"])
def test_patch_gather_hook_uncompressed_modes(one_rank_group, monkeypatch, mode):
    """patch_gather_fwd without compression: synchronous all-gather and DistriFusion's stale-async mode
    (patchpara/fwd.py:103-173).  W = 1: attention must see exactly the fresh local K / V in both."""
    dev = one_rank_group
    import compactfusion_b200 as cf
    from compactfusion_b200.attention import attn_forward
    from compactfusion_b200.patchpara.fwd import patch_gather_fwd
    T = cf.COMPACT_COMPRESS_TYPE
    cfg = cf.CompactConfig(enabled=True, override_with_patch_gather_fwd=True,
                           patch_gather_fwd_config=cf.PatchConfig(False, mode == "async", 1),
                           compress_func=lambda l, s: T.WARMUP, comp_rank=-1, residual=1, ef=True, fastpath=True)
    cf.compact_init(cfg)
    data = _series(1, 64, 8, 128, 3, 2, dev, seed=2)
    for t, step in enumerate(data):
        for l, (q, k, v) in enumerate(step):
            out, lse, _ = patch_gather_fwd(q, k, v, causal=False, mod_idx=l, current_iter=t)
            ref, _ = attn_forward(q, k, v, 0.0, None, causal=False)
            assert torch.equal(out, ref)


WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["CF_ROOT"])
import torch, torch.distributed as dist
import compactfusion_b200 as cf
from compactfusion_b200 import attention, dropin
from compactfusion_b200.attention import attn_forward
T = cf.COMPACT_COMPRESS_TYPE
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
bs, s, h, d, layers, steps = 1, 576, 24, 128, 3, 4
def tensor(t, l, j, r):
    g = torch.Generator().manual_seed(1000 * r + 10 * l + j)
    x0 = torch.randn(bs, s, h, d, generator=g)
    g2 = torch.Generator().manual_seed(77 + 1000 * r + 10 * l + j + 100000 * t)
    return (0.97 ** t * x0 + 0.2 * torch.randn(bs, s, h, d, generator=g2)).half().to(dev)
def run(codec, patch, engine, transport):
    os.environ["CF_DROPIN_ENGINE"] = "1" if engine else "0"
    os.environ["CF_DROPIN_TRANSPORT"] = transport
    kw = dict(override_with_patch_gather_fwd=True, patch_gather_fwd_config=cf.PatchConfig(True, False, 1)) if patch else {}
    cf.compact_init(cf.CompactConfig(enabled=True, compress_func=lambda l, st: codec if st >= 1 else T.WARMUP, comp_rank=-1,
                                     residual=1, ef=True, fastpath=True, **kw))
    seen, outs = [], []
    def spy(q, k, v, p, scale, causal, window):
        seen.append((k.clone(), v.clone()))
        attention.set_attention_override(None)
        try:
            return attn_forward(q, k, v, p, scale, causal, window)
        finally:
            attention.set_attention_override(spy)
    attention.set_attention_override(spy)
    for t in range(steps):
        cf.compact_set_step(t)
        for l in range(layers):
            out, lse, _ = cf.compact_fwd(tensor(t, l, 0, rank), tensor(t, l, 1, rank), tensor(t, l, 2, rank), causal=False,
                                         mod_idx=l, current_iter=t)
            outs.append(out.clone())
    attention.set_attention_override(None)
    torch.cuda.synchronize()
    engs = dropin.engines()
    if engine:
        assert len(engs) == 1 and engs[0].transport == transport, (len(engs), engs[0].transport if engs else None)
        assert not engs[0].p2p_error()
        # every rank holds bit-identical caches of every origin (the error-feedback invariant)
        for g in engs[0].global_k + engs[0].global_v:
            both = [torch.empty_like(g) for _ in range(world)]
            dist.all_gather(both, g)
            assert all(torch.equal(b, both[0]) for b in both)
    return seen, outs
def rel(a, b):
    return float(torch.norm(a.float() - b.float()) / torch.norm(b.float()))
for codec in (T.BINARY, T.INT2):
    for patch in (True, False):
        ref_seen, ref_outs = run(codec, patch, False, "nccl")
        for transport in ("p2p", "nccl"):
            seen, outs = run(codec, patch, True, transport)
            assert len(seen) == len(ref_seen)
            for (k, v), (rk, rv) in zip(seen, ref_seen):
                assert k.shape == rk.shape and rel(k, rk) < 2e-3 and rel(v, rv) < 2e-3, (codec, patch, transport, rel(k, rk))
            for o, ro in zip(outs, ref_outs):
                assert torch.allclose(o.float(), ro.float(), atol=5e-3), (codec, patch, transport)
# the uncompressed modes of patch_gather_fwd: sync == attention over the gathered raw K/V; stale-async uses the
# PREVIOUS step's K/V of the peers and the fresh local shard (patchpara/fwd.py:113-173)
from compactfusion_b200.patchpara.fwd import patch_gather_fwd
for async_comm in (False, True):
    cf.compact_init(cf.CompactConfig(enabled=True, override_with_patch_gather_fwd=True,
                                     patch_gather_fwd_config=cf.PatchConfig(False, async_comm, 1),
                                     compress_func=lambda l, st: T.WARMUP, comp_rank=-1, residual=1, ef=True, fastpath=True))
    for t in range(3):
        for l in range(2):
            q, k, v = tensor(t, l, 0, rank), tensor(t, l, 1, rank), tensor(t, l, 2, rank)
            out, _, _ = patch_gather_fwd(q, k, v, causal=False, mod_idx=l, current_iter=t)
            stale = async_comm and t >= 1
            ks = [tensor(t - 1 if (stale and r != rank) else t, l, 1, r) for r in range(world)]
            vs = [tensor(t - 1 if (stale and r != rank) else t, l, 2, r) for r in range(world)]
            ref, _ = attn_forward(q, torch.cat(ks, dim=1), torch.cat(vs, dim=1), 0.0, None, causal=False)
            assert torch.equal(out, ref), (async_comm, t, l)
dropin.shutdown()
dist.barrier()
dist.destroy_process_group()
print("WORKER_OK", rank)
'''


@pytest.mark.multigpu(2)
def test_two_gpu_hooks_on_engines(tmp_path):
    """2 GPUs: both hooks x both codecs, engine (one-sided transport and NCCL) against the per-call path; the
    uncompressed sync / stale-async modes of patch_gather_fwd against plain attention."""
    script = tmp_path / "hooks2.py"
    script.write_text(WORKER)
    env = dict(os.environ, CF_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT=free_port(), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=900)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"WORKER_OK {r}" in o, o[-4000:]
