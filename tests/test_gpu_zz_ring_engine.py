"""RingExchangeEngine (compressed ring attention on persistent buffers, one-sided exchange), the fused
LSE merge and the one-pass error statistics, against the drop-in API / plain torch."""
import math
import os
import subprocess
import sys

import pytest
import torch

from conftest import free_port

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("shape", [(1, 130, 3, 64), (2, 257, 16, 72), (1, 544, 24, 128)])
def test_lse_merge_matches_torch(shape):
    dev = _cuda()
    from compactfusion_b200.attention import merge_out_and_lse, update_out_and_lse
    b, s, h, d = shape
    g = torch.Generator().manual_seed(s)
    blocks = [(torch.randn(b, s, h, d, generator=g).half().to(dev), (3 * torch.randn(b, h, s, generator=g)).to(dev))
              for _ in range(4)]
    out = lse = None
    ref_out = ref_lse = None
    for bo, bl in blocks:
        out, lse = merge_out_and_lse(out, lse, bo, bl)
        ref_out, ref_lse = update_out_and_lse(ref_out, ref_lse, bo, bl)
    torch.cuda.synchronize()
    # fp32 on both sides; expf / log1pf vs torch's sigmoid / logsigmoid differ by a few ulp per merge
    assert torch.allclose(out, ref_out, atol=2e-5, rtol=1e-5), float((out - ref_out).abs().max())
    assert torch.allclose(lse, ref_lse.squeeze(-1).transpose(1, 2), atol=5e-5, rtol=1e-6)
    # merging a block with a vanishing weight leaves the state alone; a dominant block replaces it
    tiny = torch.full((b, h, s), -80.0, device=dev)
    o2, l2 = merge_out_and_lse(out.clone(), lse, blocks[0][0], tiny)
    assert torch.allclose(o2, out, atol=2e-6) and torch.allclose(l2, lse, atol=2e-5)
    huge = torch.full((b, h, s), 80.0, device=dev)
    o3, l3 = merge_out_and_lse(out.clone(), lse, blocks[1][0], huge)
    assert torch.allclose(o3, blocks[1][0].float(), atol=2e-6) and torch.allclose(l3, huge, atol=1e-4)


@pytest.mark.parametrize("shape", [(8,), (130, 64), (1088, 3072), (4096, 3072)])
def test_error_stats_match_torch(shape):
    dev = _cuda()
    from compactfusion_b200.quality import QualityTrace, error_stats
    g = torch.Generator().manual_seed(len(shape) + shape[0])
    ref = torch.randn(*shape, generator=g).half().to(dev)
    test = (ref.float() + 0.05 * torch.randn(*shape, generator=g).to(dev)).half()
    got = error_stats(test, ref)
    d = test.double() - ref.double()
    sse, ssr = float((d * d).sum()), float((ref.double() ** 2).sum())
    assert abs(got["max_abs"] - float(d.abs().max())) <= 1e-6 * float(d.abs().max())
    assert abs(got["rel_l2"] - math.sqrt(sse / ssr)) <= 1e-5 * math.sqrt(sse / ssr)
    peak = float(ref.float().abs().max())
    assert abs(got["psnr_db"] - 10 * math.log10(peak * peak / (sse / ref.numel()))) < 1e-3
    # deterministic, reusable workspace, and exact zero on identical inputs
    assert error_stats(test, ref) == got
    same = error_stats(ref, ref)
    assert same["max_abs"] == 0.0 and same["rel_l2"] == 0.0 and same["psnr_db"] == math.inf
    trace = QualityTrace(3, dev)
    trace.record("a", test, ref)
    trace.record("b", ref, ref)
    rows = trace.rows()
    assert rows[0]["tag"] == "a" and rows[0]["max_abs"] == got["max_abs"] and rows[1]["max_abs"] == 0.0


def _kv(n, c, steps, layers, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    base = [[torch.randn(n, c, generator=g) for _ in range(2)] for _ in range(layers)]
    return [[[(0.97 ** t * base[l][j] + 0.2 * torch.randn(n, c, generator=g)).half().to(dev) for j in range(2)]
             for l in range(layers)] for t in range(steps)]


@pytest.mark.parametrize("codec", ["binary", "int2"])
def test_ring_engine_world1_equals_patch_engine_and_plain_attention(codec):
    """W = 1: the ring is hop 0 only -- the cache update must equal the patch engine's, and the
    attention output is plain attention over the RAW local K/V (ring.py:197-208)."""
    dev = _cuda()
    import compactfusion_b200 as cf
    from compactfusion_b200.attention import attn_forward
    from compactfusion_b200.engine import PatchGatherEngine, RingExchangeEngine
    T = cf.COMPACT_COMPRESS_TYPE
    ctype = T(codec)
    bs, s, h, d, layers, steps = 2, 136, 16, 72, 2, 4  # PixArt-like head geometry, C = 1152
    n, c = bs * s, h * d
    data = _kv(n, c, steps, layers, dev, seed=5)
    ring, patch = RingExchangeEngine(layers, n, c, device=dev), PatchGatherEngine(layers, n, c, device=dev)
    for t in range(steps):
        ct = ctype if t >= 1 else T.WARMUP
        for l in range(layers):
            k, v = data[t][l][0].view(bs, s, h, d), data[t][l][1].view(bs, s, h, d)
            q = data[t][l][0].flip(0).contiguous().view(bs, s, h, d)
            out, lse = ring.ring_forward(l, q, k, v, ct)
            gk, gv = patch.exchange(l, k, v, ct)
            assert torch.equal(ring.global_k[l], gk) and torch.equal(ring.global_v[l], gv), (t, l)
            ref, ref_lse = attn_forward(q, k, v, 0.0, None, causal=False)
            assert torch.allclose(out.float(), ref.float(), atol=2e-3)
            assert torch.allclose(lse, ref_lse, atol=1e-4)
    torch.cuda.synchronize()


WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["CF_ROOT"])
import torch, torch.distributed as dist
import compactfusion_b200 as cf
from compactfusion_b200.attention import attn_forward
from compactfusion_b200.engine import PatchGatherEngine, RingExchangeEngine
T = cf.COMPACT_COMPRESS_TYPE
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
bs, s, h, d, layers, steps = 2, 272, 24, 128, 3, 4
n, c = bs * s, h * d
def shard(t, l, j, r):
    g = torch.Generator().manual_seed(1000 * r + 10 * l + j)
    x0 = torch.randn(n, c, generator=g)
    g2 = torch.Generator().manual_seed(77 + 1000 * r + 10 * l + j + 100000 * t)
    return (0.97 ** t * x0 + 0.2 * torch.randn(n, c, generator=g2)).half()
for codec in (T.BINARY, T.INT2):
    for transport in ("p2p", "nccl"):
        ring = RingExchangeEngine(layers, n, c, device=dev, transport=transport)
        assert ring.prepare(codec) == transport
        patch = PatchGatherEngine(layers, n, c, device=dev, transport="nccl")
        for t in range(steps):
            ct = codec if t >= 1 else T.WARMUP
            for l in range(layers):
                k = shard(t, l, 0, rank).to(dev).view(bs, s, h, d)
                v = shard(t, l, 1, rank).to(dev).view(bs, s, h, d)
                q = shard(t, l, 0, (rank + 1) % world).to(dev).view(bs, s, h, d)
                before_k = [ring._shard(ring.global_k[l], r).clone() for r in range(world)]
                before_v = [ring._shard(ring.global_v[l], r).clone() for r in range(world)]
                out, lse = ring.ring_forward(l, q, k, v, ct)
                gk, gv = patch.exchange(l, k, v, ct)
                # every origin's cache is what the all-gather engine reconstructs (bit-identical on all ranks)
                assert torch.equal(ring.global_k[l], gk) and torch.equal(ring.global_v[l], gv), (codec, transport, t, l)
                # attention saw the raw local block at hop 0 and the reconstructions of the peers
                kk = [(k if r == rank else ring._shard(gk, r).view(bs, s, h, d)) for r in range(world)]
                vv = [(v if r == rank else ring._shard(gv, r).view(bs, s, h, d)) for r in range(world)]
                ref, ref_lse = attn_forward(q, torch.cat(kk, dim=1), torch.cat(vv, dim=1), 0.0, None, causal=False)
                assert torch.allclose(out.float(), ref.float(), atol=3e-3), float((out.float() - ref.float()).abs().max())
                assert torch.allclose(lse, ref_lse, atol=1e-3)
        assert not ring.p2p_error(), "a device-side flag wait timed out"
        if transport == "p2p":
            # the whole ring step (no attention) as ONE replayed CUDA graph == the eager all-gather engine
            ks = [shard(steps, l, 0, rank).to(dev) for l in range(layers)]
            vs = [shard(steps, l, 1, rank).to(dev) for l in range(layers)]
            g = ring.capture_step(ks, vs, codec, warmup_iters=0)
            assert ring.launches_per_graph == layers * ((2 if codec == T.BINARY else 3) + world)
            g.replay()
            torch.cuda.synchronize()
            for l in range(layers):
                gk, gv = patch.exchange(l, ks[l], vs[l], codec)
                assert torch.equal(ring.global_k[l], gk) and torch.equal(ring.global_v[l], gv), ("graph", codec, l)
            assert not ring.p2p_error()
        dist.barrier()
dist.destroy_process_group()
print("WORKER_OK", rank)
'''


def test_two_gpu_ring_engine(tmp_path):
    """2 GPUs (skipped on a 1-GPU box): ring consumption order over the one-sided transport and over NCCL."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "ring2.py"
    script.write_text(WORKER)
    env = dict(os.environ, CF_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT=free_port(), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"WORKER_OK {r}" in o, o[-4000:]


@pytest.mark.parametrize("mode", ["fastpath_binary", "sim_int4_r1", "sim_int2_r2"])
def test_log_stats_through_the_plugin_api(mode, tmp_path, capsys):
    """CompactConfig(log_stats=True): compact_compress records error / norm figures on the GPU without a
    synchronisation per call; the read-back equals torch reductions of the same tensors (stats.py:107-328)."""
    dev = _cuda()
    import compactfusion_b200 as cf
    from compactfusion_b200 import stats as st
    T = cf.COMPACT_COMPRESS_TYPE
    kw, ctype = {
        "fastpath_binary": (dict(residual=1, ef=True, fastpath=True, comp_rank=-1), T.BINARY),
        "sim_int4_r1": (dict(residual=1, ef=True, simulate=True, comp_rank=-1), T.INT4),
        "sim_int2_r2": (dict(residual=2, ef=True, simulate=True, comp_rank=-1, delta_decay_factor=0.5), T.INT2),
    }[mode]
    n, c, steps = 256, 512, 5
    shape = (1, n, 8, c // 8)
    g = torch.Generator().manual_seed(11)
    xs = [torch.randn(n, c, generator=g)]
    for _ in range(steps - 1):
        xs.append(0.95 * xs[-1] + 0.3 * torch.randn(n, c, generator=g))
    xs = [x.half().view(shape).to(dev) for x in xs]
    cf.compact_init(cf.CompactConfig(enabled=True, log_stats=True, compress_func=lambda l, s: ctype, **kw))
    cf.compact_set_inplace(True)  # must not alias the old base while it is being logged
    try:
        want = []
        warm = 2 if kw["residual"] == 2 else 1
        for t, x in enumerate(xs):
            cf.compact_set_step(t)
            ct = ctype if t >= warm else T.WARMUP
            base = cf.compact_cache().get_base("0-0-k")
            base = None if base is None else base.clone()
            comp = cf.compact_compress("0-0-k", x, ct, update_cache=True)
            if ct != T.WARMUP:
                new_base = cf.compact_cache().get_base("0-0-k")
                x2 = x.view(n, c).double()
                want.append(dict(error=float(torch.norm(x2 - new_base.double())), activation_norm=float(torch.norm(x2)),
                                 delta_norm=float(torch.norm(x2 - base.double())), comp_bytes=comp.numel() * 2,
                                 max_abs=float((x2 - new_base.double()).abs().max())))
        recs = st.stats_log().stats["0-0-k"]
        assert len(recs) == len(want) == steps - warm
        for r, w in zip(recs, want):
            for name in ("error", "activation_norm", "delta_norm"):
                assert abs(r[name] - w[name]) <= 2e-5 * w[name], (mode, name, r[name], w[name])
            assert abs(r["max_abs_error"] - w["max_abs"]) <= 1e-6 * max(w["max_abs"], 1e-3)
            assert r["compressed_size_bytes"] == w["comp_bytes"] and r["original_size_bytes"] == n * c * 2
            assert r["residual"] == kw["residual"] and (r["delta_delta_norm"] is not None) == (kw["residual"] == 2)
        assert recs[0]["activation_similarity"] is None and 0.5 < recs[1]["activation_similarity"] < 1.0
        cos = float(torch.nn.functional.cosine_similarity(xs[-1].double().flatten(), xs[-2].double().flatten(), dim=0))
        assert abs(recs[-1]["activation_similarity"] - cos) < 1e-4
        st.stats_verbose()
        d = st.dump_err_vs_steps(str(tmp_path))
        assert "avg comp error" in capsys.readouterr().out
        assert len(d["avg_comp_errors"]) == steps - warm and abs(d["avg_comp_errors"][0] - want[0]["error"]) <= 2e-5 * want[0]["error"]
    finally:
        cf.compact_set_inplace(False)
        st.stats_clear()


# The two-chain step is opt-in and has not run on a GPU yet: its tests are armed by CF_EXPERIMENTAL=1
# (tools/gpu_round.sh and tools/gpu_multi.sh set it) until the schedule has been measured and made default.
experimental = pytest.mark.skipif(os.environ.get("CF_EXPERIMENTAL", "0") != "1",
                                  reason="opt-in feature, not yet measured: set CF_EXPERIMENTAL=1")


@experimental
@pytest.mark.parametrize("codec", ["binary", "int2"])
def test_overlapped_step_equals_serial_world1(codec):
    """engine._step_overlapped (compress chain | reconstruct chain, per-layer events, lag 2): same kernels and
    operands as the serial step -> bit-identical caches, eagerly and as a replayed CUDA graph."""
    dev = _cuda()
    import compactfusion_b200 as cf
    from compactfusion_b200.engine import PatchGatherEngine
    T = cf.COMPACT_COMPRESS_TYPE
    ctype = T(codec)
    n, c, layers, steps = 576, 3072, 7, 4
    data = _kv(n, c, steps, layers, dev, seed=9)
    serial, over, graph_eng = (PatchGatherEngine(layers, n, c, device=dev) for _ in range(3))
    assert over.can_overlap(ctype) and not PatchGatherEngine(3, n, c, device=dev).can_overlap(ctype)
    ks = [data[0][l][0].clone() for l in range(layers)]
    vs = [data[0][l][1].clone() for l in range(layers)]
    for e in (serial, over, graph_eng):
        e.step(ks, vs, T.WARMUP, overlap=True)  # WARMUP ignores the flag
    graph = None
    for t in range(1, steps):
        for l in range(layers):
            ks[l].copy_(data[t][l][0])
            vs[l].copy_(data[t][l][1])
        serial.step(ks, vs, ctype)
        over.step(ks, vs, ctype, overlap=True)
        if graph is None:
            snap = [(a.clone(), b.clone()) for a, b in zip(graph_eng.global_k, graph_eng.global_v)]
            graph = graph_eng.capture_step(ks, vs, ctype, warmup_iters=1, overlap=True)
            for l, (a, b) in enumerate(snap):  # the capture's warm-up run advanced the cache: restore it
                graph_eng.global_k[l].copy_(a)
                graph_eng.global_v[l].copy_(b)
            assert graph_eng.launches_per_graph == layers * (3 if codec == "binary" else 4)
        graph.replay()
        torch.cuda.synchronize()
        for l in range(layers):
            assert torch.equal(over.global_k[l], serial.global_k[l]) and torch.equal(over.global_v[l], serial.global_v[l]), (t, l)
            assert torch.equal(graph_eng.global_k[l], serial.global_k[l]), ("graph", t, l)
            assert torch.equal(graph_eng.global_v[l], serial.global_v[l]), ("graph", t, l)


OVERLAP_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["CF_ROOT"])
import torch, torch.distributed as dist
import compactfusion_b200 as cf
from compactfusion_b200.engine import PatchGatherEngine
T = cf.COMPACT_COMPRESS_TYPE
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
n, c, layers, steps = 288, 3072, 8, 6
def shard(t, l, j):
    g = torch.Generator().manual_seed(1000 * rank + 10 * l + j)
    x0 = torch.randn(n, c, generator=g)
    g2 = torch.Generator().manual_seed(77 + 1000 * rank + 10 * l + j + 100000 * t)
    return (0.97 ** t * x0 + 0.2 * torch.randn(n, c, generator=g2)).half().to(dev)
for codec in (T.BINARY, T.INT2):
    serial = PatchGatherEngine(layers, n, c, device=dev, transport="nccl")
    over = PatchGatherEngine(layers, n, c, device=dev, transport="p2p")
    assert over.prepare(codec) == "p2p" and over.can_overlap(codec) and not serial.can_overlap(codec)
    ks = [shard(0, l, 0) for l in range(layers)]
    vs = [shard(0, l, 1) for l in range(layers)]
    serial.step(ks, vs, T.WARMUP)
    over.step(ks, vs, T.WARMUP)
    graph = None
    for t in range(1, steps):
        for l in range(layers):
            ks[l].copy_(shard(t, l, 0))
            vs[l].copy_(shard(t, l, 1))
        serial.step(ks, vs, codec)
        if t < 3:
            over.step(ks, vs, codec, overlap=True)       # eager two-chain steps
        else:
            if graph is None:
                graph = over.capture_step(ks, vs, codec, warmup_iters=0, overlap=True)
            for _ in range(1):
                graph.replay()                           # the same step as one graph with two branches
        torch.cuda.synchronize()
        for l in range(layers):
            assert torch.equal(over.global_k[l], serial.global_k[l]) and torch.equal(over.global_v[l], serial.global_v[l]), (codec, t, l)
    assert not over.p2p_error(), "a device-side flag wait timed out"
    dist.barrier()
dist.destroy_process_group()
print("WORKER_OK", rank)
'''


@experimental
def test_two_gpu_overlapped_step(tmp_path):
    """2 GPUs (skipped on a 1-GPU box): the two-chain step over the one-sided transport, eager and as a replayed
    graph, against the serial NCCL engine."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "overlap2.py"
    script.write_text(OVERLAP_WORKER)
    env = dict(os.environ, CF_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT=free_port(), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"WORKER_OK {r}" in o, o[-4000:]


def test_plain_c_client_round_trip(tmp_path):
    """examples/c_client.c: a C99 program (no torch) drives one BINARY residual round trip through the C ABI on
    host buffers and checks sign bits, reconstruction and the sender/receiver identity bit for bit with its own
    fp16 arithmetic (the checker is pinned against the oracle in tests/test_abi_and_host.py)."""
    _cuda()
    from compactfusion_b200 import _native as nv
    nv.lib()  # the library this process already uses (never rebuild a mapped .so)
    libdir = os.path.dirname(nv.LIB_PATH)
    exe = str(tmp_path / "c_client")
    cmd = ["gcc", "-std=c99", "-O2", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
           os.path.join(ROOT, "examples", "c_client.c"), "-o", exe, "-L", libdir, "-lcompactb200",
           f"-Wl,-rpath,{libdir}", "-L", "/usr/local/cuda/lib64", "-lcudart", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "C_CLIENT_OK 544x3072" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("name", ["rand_64x256", "rand_48x1152", "rand_130x64", "flux_k_96x512"])
def test_sim_int2_minmax_matches_reference_golden(name):
    """The 4-level min/max simulation codec (cf_int2mm_compress: the INT4 kernels at qmax = 3) against what
    the reference's sim_int2_minmax produced on the committed inputs: bit-exact (min/max are exact)."""
    dev = _cuda()
    import numpy as np
    from conftest import GOLDEN, assert_bits_equal, h16
    from compactfusion_b200.compress_quantize import sim_int2_minmax
    from compactfusion_b200.slowpath import sim_compress
    from compactfusion_b200.utils import COMPACT_COMPRESS_TYPE as T
    from oracle import codecs as oc
    g = np.load(os.path.join(GOLDEN, "codecs.npz"))
    d = h16(g[f"{name}/x"]) - h16(g[f"{name}/base"])
    got = sim_int2_minmax(d.to(dev))
    assert_bits_equal(got, h16(g[f"{name}/sim_int2_minmax"]), "sim_int2_minmax")
    assert torch.equal(sim_compress(d.to(dev), T.INT2_MINMAX), got)
    assert all(len(torch.unique(got[:, c])) <= 4 for c in range(0, got.shape[1], 17))  # 4 levels per channel
    odd = d[:-1].contiguous()  # odd N: padded with a copy of the last row inside the wrapper
    assert_bits_equal(sim_int2_minmax(odd.to(dev)), oc.sim_int2_minmax(odd), "sim_int2_minmax, odd N")


def test_quantized_cache_stores_int8_and_keeps_sender_and_receiver_identical(monkeypatch):
    """CompactConfig(quantized_cache=True) (deprecated in the reference, gated by COMPACT_ALLOW_DEPRECATED like
    there): bases are stored as per-channel int8 and dequantised on every read (utils.py:123-160)."""
    dev = _cuda()
    import compactfusion_b200 as cf
    from compactfusion_b200 import utils
    from compactfusion_b200.compress_quantize import dequantize_int8, quantize_int8
    from conftest import rel_l2
    monkeypatch.setattr(utils, "ALLOW_DEPRECATED", False)
    with pytest.raises(AssertionError):
        utils.CompactCache(quantize=True)
    with pytest.raises(AssertionError):
        cf.CompactConfig(enabled=True, residual=1, ef=True, quantized_cache=True)
    monkeypatch.setattr(utils, "ALLOW_DEPRECATED", True)
    g = torch.Generator().manual_seed(21)
    x = torch.randn(256, 512, generator=g).half().to(dev)
    cache = utils.CompactCache(quantize=True)
    cache.put("0-0-k", x, None)
    q, scale, zp, shape = cache.base["0-0-k"]
    assert q.dtype == torch.int8 and shape == x.shape
    got = cache.get_base("0-0-k")
    assert torch.equal(got, dequantize_int8(*quantize_int8(x))) and got.shape == x.shape
    assert cache.get_delta_base("0-0-k") is None and cache.get_base("missing") is None
    assert rel_l2(got, x) < 2e-2
    # through the plugin: residual 1 + EF on a quantised cache; both sides cache the same reconstruction
    T = cf.COMPACT_COMPRESS_TYPE
    cfg = cf.CompactConfig(enabled=True, residual=1, ef=True, simulate=True, quantized_cache=True, comp_rank=-1,
                           compress_func=lambda l, s: T.INT4 if s >= 1 else T.WARMUP)
    cf.compact_init(cfg)
    shape4 = (1, 256, 8, 64)
    xs = [x.view(shape4)]
    for t in range(1, 4):
        xs.append((0.97 * xs[-1].float() + 0.2 * torch.randn(shape4, generator=g).to(dev)).half())
    for t, xt in enumerate(xs):
        ct = cfg.compress_func(0, t)
        comp = cf.compact_compress("0-0-k", xt, ct, update_cache=True)
        rec = cf.compact_decompress("1-0-k", comp, ct, shape4, update_cache=True)
        assert isinstance(cf.compact_cache().base["0-0-k"], tuple)
        assert torch.equal(cf.compact_cache().get_base("0-0-k"), cf.compact_cache().get_base("1-0-k"))
        assert rel_l2(rec.reshape(-1), xt.reshape(-1)) < 0.3
