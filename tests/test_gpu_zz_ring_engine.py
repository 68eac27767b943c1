"""RingExchangeEngine (compressed ring attention on persistent buffers, one-sided exchange) against the
all-gather engine / plain attention, and the opt-in two-chain step against the serial one."""
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
                assert torch.allclose(out.float(), ref.float(), atol=8e-3), float((out.float() - ref.float()).abs().max())  # 2 fp16 ulps at |out| ~ 4
                assert torch.allclose(lse, ref_lse, atol=1e-3)
        assert not ring.p2p_error(), "a device-side flag wait timed out"
        if transport == "p2p":
            # the whole ring step (no attention) as ONE replayed CUDA graph == the eager all-gather engine
            ks = [shard(steps, l, 0, rank).to(dev) for l in range(layers)]
            vs = [shard(steps, l, 1, rank).to(dev) for l in range(layers)]
            g = ring.capture_step(ks, vs, codec, warmup_iters=0)
            assert ring.launches_per_graph == layers * ((3 if codec == T.BINARY else 4) + world)  # + k_publish_flags
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


@pytest.mark.multigpu(2)
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


# The two-chain step is opt-in and has not run on a GPU yet: its tests are armed by CF_EXPERIMENTAL=1
# (tools/gpu_round.sh and tools/gpu_multi.sh set it) until the schedule has been measured and made default.
experimental = pytest.mark.skipif(os.environ.get("CF_EXPERIMENTAL", "1") != "1",
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
@pytest.mark.multigpu(2)
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
