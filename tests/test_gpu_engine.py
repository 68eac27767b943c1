"""PatchGatherEngine (persistent buffers, batched launches, CUDA graph) vs the drop-in API."""
import os

import pytest
import torch

from conftest import free_port, rel_l2

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _data(n, c, steps, layers, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    base = [[torch.randn(n, c, generator=g) for _ in range(2)] for _ in range(layers)]
    for t in range(steps):
        out.append([[(0.97 ** t * base[l][j] + 0.2 * torch.randn(n, c, generator=g)).half().to(dev) for j in range(2)]
                    for l in range(layers)])
    return out  # out[t][layer][kv]


@pytest.mark.parametrize("codec", ["binary", "int2"])
def test_engine_matches_plain_api_world1(codec):
    dev = _cuda()
    import compactfusion_b200 as cf
    from compactfusion_b200.engine import PatchGatherEngine
    from compactfusion_b200.fastpath import binary_dequant_fastpath, int2_dequant_fastpath
    from compactfusion_b200.main import _payload_views
    T = cf.COMPACT_COMPRESS_TYPE
    ctype = T(codec)
    n, c, layers, steps = 1152, 3072, 3, 4
    data = _data(n, c, steps, layers, dev)
    eng = PatchGatherEngine(layers, n, c, device=dev)
    prev = None
    for t in range(steps):
        ct = ctype if t >= 1 else T.WARMUP
        for l in range(layers):
            before_k = eng.global_k[l].clone()
            gk, gv = eng.exchange(l, data[t][l][0], data[t][l][1], ct)
            if t == 0:
                assert torch.equal(gk, data[0][l][0]) and torch.equal(gv, data[0][l][1])
                continue
            # the payload the engine put on the wire decodes, through the plain API, to exactly
            # what the engine wrote into its global buffers
            send, recv = eng._buffers(ctype)
            packed, u, v = _payload_views(recv[0, 0][:eng._numel(ctype)].clone(), n, c, ctype)
            fn = binary_dequant_fastpath if codec == "binary" else int2_dequant_fastpath
            assert torch.equal(fn(packed, u, v, before_k), gk), (t, l)
            # sign bits of the wire codes are exactly (x - base >= 0)
            bits = (data[t][l][0] - before_k) >= 0
            if codec == "binary":
                got = ((packed.unsqueeze(-1) >> torch.arange(8, device=dev, dtype=torch.uint8)) & 1).view(n, c).bool()
            else:
                got = (((packed.unsqueeze(-1) >> (2 * torch.arange(4, device=dev, dtype=torch.uint8))) & 3) >> 1
                       ).view(n, c).bool()
            assert torch.equal(bits, got)
            # error feedback keeps the reconstruction close
            assert rel_l2(gk, data[t][l][0]) < 0.3
    torch.cuda.synchronize()


@pytest.mark.parametrize("codec", ["binary", "int2"])
def test_engine_fast_launch_options_are_bit_neutral(codec, monkeypatch):
    """L2 eviction hints, programmatic dependent launch and the early pipeline fill
    (CF_FLAG_INPUTS_STABLE) only change WHEN bytes move: a 4-step run of a 3-layer engine is
    bit-identical with all of them switched off."""
    dev = _cuda()
    import compactfusion_b200 as cf
    from compactfusion_b200.engine import PatchGatherEngine
    T = cf.COMPACT_COMPRESS_TYPE
    ctype = T(codec)
    n, c, layers, steps = 1150, 3072, 3, 4  # ragged: 1150 = 287 full tiles of 4 rows + 2
    data = _data(n, c, steps, layers, dev, seed=11)

    def run(fast):
        if not fast:
            monkeypatch.setenv("CF_L2_HINTS", "0")
            monkeypatch.setenv("CF_PDL", "0")
        eng = PatchGatherEngine(layers, n, c, device=dev)
        if not fast:
            eng._flags = 0
        for t in range(steps):
            ks = [data[t][l][0] for l in range(layers)]
            vs = [data[t][l][1] for l in range(layers)]
            eng.step(ks, vs, ctype if t >= 1 else T.WARMUP)
        torch.cuda.synchronize()
        monkeypatch.delenv("CF_L2_HINTS", raising=False)
        monkeypatch.delenv("CF_PDL", raising=False)
        return [g.clone() for g in eng.global_k + eng.global_v]

    fast, plain = run(True), run(False)
    for a, b in zip(fast, plain):
        assert torch.equal(a, b)


def test_engine_graph_replay_equals_eager():
    dev = _cuda()
    import compactfusion_b200 as cf
    from compactfusion_b200.engine import PatchGatherEngine
    T = cf.COMPACT_COMPRESS_TYPE
    n, c, layers = 576, 3072, 4
    data = _data(n, c, 3, layers, dev, seed=3)
    engines = [PatchGatherEngine(layers, n, c, device=dev) for _ in range(2)]
    ks = [data[1][l][0].clone() for l in range(layers)]
    vs = [data[1][l][1].clone() for l in range(layers)]
    for e in engines:
        e.step([data[0][l][0] for l in range(layers)], [data[0][l][1] for l in range(layers)], T.WARMUP)
    # eager: two compressed steps
    engines[0].step(ks, vs, T.BINARY)
    eager1 = [g.clone() for g in engines[0].global_k]
    ks2 = [data[2][l][0] for l in range(layers)]
    vs2 = [data[2][l][1] for l in range(layers)]
    engines[0].step(ks2, vs2, T.BINARY)
    # graph: capture on static input buffers, refresh their contents in place, replay
    snapshot_k = [g.clone() for g in engines[1].global_k]
    snapshot_v = [g.clone() for g in engines[1].global_v]
    graph = engines[1].capture_step(ks, vs, T.BINARY)
    for l in range(layers):  # capture (and its warm-up run) advanced the cache: restore it
        engines[1].global_k[l].copy_(snapshot_k[l])
        engines[1].global_v[l].copy_(snapshot_v[l])
    graph.replay()
    for l in range(layers):
        assert torch.equal(engines[1].global_k[l], eager1[l])
    for l in range(layers):
        ks[l].copy_(ks2[l])
        vs[l].copy_(vs2[l])
    graph.replay()
    torch.cuda.synchronize()
    for l in range(layers):
        assert torch.equal(engines[1].global_k[l], engines[0].global_k[l])
        assert torch.equal(engines[1].global_v[l], engines[0].global_v[l])
    assert engines[1].launches_per_graph == layers * 3


def test_host_buffer_c_abi_round_trip():
    """cf_host_compress / cf_host_decompress: pinned host buffers in, payload / reconstruction out."""
    dev = _cuda()
    from compactfusion_b200 import _native as nv
    from compactfusion_b200.fastpath import binary_dequant_fastpath, binary_quant_fastpath
    n, c = 544, 3072
    g = torch.Generator().manual_seed(4)
    x = torch.randn(n, c, generator=g).half().pin_memory()
    base = (x.float() + 0.2 * torch.randn(n, c, generator=g)).half().pin_memory()
    lib = nv.lib()
    for codec, per_byte in ((nv.CODEC_BINARY, 8), (nv.CODEC_INT2, 4)):
        nbytes = n * c // per_byte + 2 * n + 2 * c
        payload = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        new_base = torch.empty_like(x).pin_memory()
        recon = torch.empty_like(x).pin_memory()
        scratch = torch.empty(lib.cf_host_scratch_bytes(codec, n, c), dtype=torch.uint8, device=dev)
        rc = lib.cf_host_compress(codec, x.data_ptr(), base.data_ptr(), new_base.data_ptr(), payload.data_ptr(), n, c,
                                  scratch.data_ptr(), scratch.numel(), nv.stream_ptr())
        nv.check(rc, "cf_host_compress")
        rc = lib.cf_host_decompress(codec, payload.data_ptr(), base.data_ptr(), recon.data_ptr(), n, c,
                                    scratch.data_ptr(), scratch.numel(), nv.stream_ptr())
        nv.check(rc, "cf_host_decompress")
        assert torch.equal(recon, new_base), "host round trip: receiver != sender"
        if codec == nv.CODEC_BINARY:
            packed, u, v, nb = binary_quant_fastpath(x.to(dev), base.to(dev), -1, True)
            assert torch.equal(payload[:n * c // 8].view(n, c // 8), packed.cpu())
            assert torch.equal(new_base, nb.cpu())


@pytest.mark.parametrize("codec", ["binary", "int2"])
@pytest.mark.parametrize("n,c", [(576, 3072), (1150, 3072), (130, 1152), (64, 4096)])
def test_fused_compress_put_matches_compress_on_local_slots(n, c, codec):
    """cf_sign_compress_put with every destination in local memory (what a peer mapping looks like to the
    kernel): all n_dst slots receive exactly the payload cf_{binary,int2}_compress_batched writes, the flags
    carry the put count, and the flag-waiting decompress reconstructs from a slot."""
    import ctypes
    dev = _cuda()
    from compactfusion_b200 import _native as nv
    lib = nv.lib()
    g = torch.Generator().manual_seed(n + c)
    xs = [torch.randn(n, c, generator=g).half().to(dev) for _ in range(2)]
    bases = [(x.float().cpu() * 0.97 + 0.2 * torch.randn(n, c, generator=g)).half().to(dev) for x in xs]
    cid = nv.CODEC_BINARY if codec == "binary" else nv.CODEC_INT2
    code_b, n_dst = n * c // (8 if codec == "binary" else 4), 3
    pn = code_b + 2 * n + 2 * c
    pn_pad = (pn + 15) // 16 * 16
    # reference: the plain batched compress into a local payload
    ref = torch.zeros(2, pn_pad, dtype=torch.uint8, device=dev)
    ws = nv.workspace(nv.workspace_bytes(cid, n, c, 0, 2), dev)
    arr = lambda ptrs: (ctypes.c_void_p * len(ptrs))(*ptrs)  # noqa: E731
    plain_compress = lib.cf_binary_compress_batched if codec == "binary" else lib.cf_int2_compress_batched
    rc = plain_compress(2, nv.ptr_array(xs), nv.ptr_array(bases), arr([None, None]),
                                        arr([ref[t].data_ptr() for t in range(2)]),
                                        arr([ref[t].data_ptr() + code_b for t in range(2)]),
                                        arr([ref[t].data_ptr() + code_b + 2 * n for t in range(2)]), n, c,
                                        ws.data_ptr(), ws.numel(), nv.stream_ptr())
    nv.check(rc, "cf_*_compress_batched")
    slots = torch.zeros(n_dst, 2, pn_pad, dtype=torch.uint8, device=dev)
    flags = torch.zeros(n_dst, dtype=torch.int32, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    ticket = torch.zeros(1, dtype=torch.int32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    dst = arr([slots[q, t].data_ptr() for t in range(2) for q in range(n_dst)])
    flg = arr([flags[q:q + 1].data_ptr() for q in range(n_dst)])
    prev = 0
    for it in range(2):
        rc = lib.cf_sign_compress_put(cid, nv.PASS_ALL, 2, nv.ptr_array(xs), nv.ptr_array(bases), n_dst, 1, dst, flg,
                                      count.data_ptr(), ticket.data_ptr(), n, c, ws.data_ptr(), ws.numel(), nv.stream_ptr())
        nv.check(rc, "cf_sign_compress_put")
        torch.cuda.synchronize()
        # a flag counts the CTA arrivals of the publishing kernel; count is what "fully landed" looks like
        assert count.item() > prev and ticket.item() == 0
        assert flags.tolist() == [count.item()] * n_dst
        prev = count.item()
    for q in range(n_dst):
        assert torch.equal(slots[q, :, :pn], ref[:, :pn]), f"slot {q} differs from the plain compress payload"
    if (code_b // n) % 16 == 0:  # the flag-waiting decompress needs the pipelined kernel
        recon = [torch.empty_like(b) for b in bases]
        plain = [torch.empty_like(b) for b in bases]
        q = n_dst - 1
        pk = arr([slots[q, t].data_ptr() for t in range(2)])
        us = arr([slots[q, t].data_ptr() + code_b for t in range(2)])
        vs = arr([slots[q, t].data_ptr() + code_b + 2 * n for t in range(2)])
        wf = arr([flags[q:q + 1].data_ptr()] * 2)
        rc = lib.cf_sign_decompress_batched_wait(cid, 2, pk, us, vs, nv.ptr_array(bases), nv.ptr_array(recon), wf,
                                                 count.data_ptr(), err.data_ptr(), n, c, nv.stream_ptr())
        nv.check(rc, "cf_sign_decompress_batched_wait")
        plain_decompress = lib.cf_binary_decompress_batched if codec == "binary" else lib.cf_int2_decompress_batched
        rc = plain_decompress(2, pk, us, vs, nv.ptr_array(bases), nv.ptr_array(plain), n, c, nv.stream_ptr())
        nv.check(rc, "cf_*_decompress_batched")
        torch.cuda.synchronize()
        assert err.item() == 0
        for a, b in zip(recon, plain):
            assert torch.equal(a, b)
    # bad arguments are reported, not executed
    rc = lib.cf_sign_compress_put(nv.CODEC_INT4, nv.PASS_ALL, 2, nv.ptr_array(xs), nv.ptr_array(bases), n_dst, 1, dst, flg,
                                  count.data_ptr(), ticket.data_ptr(), n, c, ws.data_ptr(), ws.numel(), nv.stream_ptr())
    assert rc == -1
    rc = lib.cf_sign_compress_put(cid, nv.PASS_ALL, 2, nv.ptr_array(xs), nv.ptr_array(bases), 17, 1, dst, flg,
                                  count.data_ptr(), ticket.data_ptr(), n, c, ws.data_ptr(), ws.numel(), nv.stream_ptr())
    assert rc == -1


def _two_gpu_worker_source():
    return r'''
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
n, c, layers, steps = 576, 3072, 2, 4
def shard(t, l, j, r):
    g = torch.Generator().manual_seed(1000 * r + 10 * l + j)
    x0 = torch.randn(n, c, generator=g)
    g2 = torch.Generator().manual_seed(77 + 1000 * r + 10 * l + j + 100000 * t)
    return (0.97 ** t * x0 + 0.2 * torch.randn(n, c, generator=g2)).half()
for codec in (T.BINARY, T.INT2):
    # drop-in API: patch-parallel all-gather + ring, with the reference's consistency check
    cfg = cf.CompactConfig(enabled=True, override_with_patch_gather_fwd=True,
                           patch_gather_fwd_config=cf.PatchConfig(True, False, 1),
                           compress_func=lambda l, s: codec if s >= 1 else T.WARMUP, comp_rank=-1, residual=1,
                           ef=True, fastpath=True)
    cf.compact_init(cfg)
    eng = PatchGatherEngine(layers, n, c, device=dev)
    # one-sided NVLink transport (cf_p2p.cu): must reproduce the NCCL engine bit for bit
    eng_p2p = PatchGatherEngine(layers, n, c, device=dev, transport="p2p")
    assert eng_p2p.prepare(codec) == "p2p"
    for t in range(steps):
        ct = cfg.compress_func(0, t)
        for l in range(layers):
            k = shard(t, l, 0, rank).to(dev).view(1, n, 24, 128)
            v = shard(t, l, 1, rank).to(dev).view(1, n, 24, 128)
            k_list = cf.compact_all_gather(f"{l}-k", k, ct)
            v_list = cf.compact_all_gather(f"{l}-v", v, ct)
            gk, gv = eng.exchange(l, k, v, ct)
            pk, pv = eng_p2p.exchange(l, k, v, ct)
            assert torch.equal(pk, gk) and torch.equal(pv, gv), ("p2p != nccl engine", codec, t, l)
            for r in range(world):
                # every rank holds the same reconstruction of every origin (EF invariant) ...
                ref = k_list[r].reshape(n, c)
                both = [torch.empty_like(ref) for _ in range(world)]
                dist.all_gather(both, ref)
                assert all(torch.equal(b, both[0]) for b in both), ("plain", codec, t, l, r)
                # ... and it tracks the origin's true shard
                true = shard(t, l, 0, r).to(dev)
                err = float(torch.norm(ref.float() - true.float()) / torch.norm(true.float()))
                assert err < 0.3, err
                # engine == plain API up to the 1-ulp scale freedom of batched reductions
                e = gk[r * n:(r + 1) * n]
                assert float(torch.norm(e.float() - ref.float()) / torch.norm(ref.float())) < 2e-2
            eg = [torch.empty_like(gk) for _ in range(world)]
            dist.all_gather(eg, gk)
            assert all(torch.equal(b, eg[0]) for b in eg), ("engine", codec, t, l)
    assert not eng_p2p.p2p_error(), "a device-side flag wait timed out"
    # the p2p step is ONE CUDA graph (no collective inside): capture, replay, compare with eager
    ks = [shard(steps, l, 0, rank).to(dev) for l in range(layers)]
    vs = [shard(steps, l, 1, rank).to(dev) for l in range(layers)]
    snap = [eng_p2p.global_k[l].clone() for l in range(layers)]
    g = eng_p2p.capture_step(ks, vs, codec, warmup_iters=0)
    for l in range(layers):
        eng_p2p.global_k[l].copy_(snap[l])  # capture does not execute: caches are still the snapshot
    g.replay()
    torch.cuda.synchronize()
    graph_k = [eng_p2p.global_k[l].clone() for l in range(layers)]
    for l in range(layers):
        gk, _ = eng.exchange(l, ks[l], vs[l], codec)
        assert torch.equal(graph_k[l], gk), ("p2p graph replay != nccl eager", codec, l)
    assert not eng_p2p.p2p_error()
    cfg = cf.CompactConfig(enabled=True, compress_func=lambda l, s: codec if s >= 1 else T.WARMUP, comp_rank=-1,
                           residual=1, ef=True, fastpath=True, check_consist=True)
    cf.compact_init(cfg)
    for t in range(steps):
        k = shard(t, 0, 0, rank).to(dev).view(1, n, 24, 128)
        v = shard(t, 0, 1, rank).to(dev).view(1, n, 24, 128)
        q = shard(t, 1, 0, rank).to(dev).view(1, n, 24, 128)
        out, lse, _ = cf.compact_fwd(q, k, v, causal=False, mod_idx=0, current_iter=t)
        assert out.shape == q.shape and torch.isfinite(out.float()).all()
    assert cf.compact_cache().passed_count == steps
dist.destroy_process_group()
print("WORKER_OK", rank)
'''


@pytest.mark.multigpu(2)
def test_two_gpu_all_gather_and_ring(tmp_path):
    """NCCL path on 2 GPUs (skipped on a 1-GPU box): drop-in all-gather, engine, compressed ring."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "w2.py"
    script.write_text(_two_gpu_worker_source())
    env = dict(os.environ, CF_ROOT=root, MASTER_ADDR="127.0.0.1", MASTER_PORT=free_port(), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"WORKER_OK {r}" in o, o[-4000:]


# ---------------------------------------------------------------------------------------------------------
# W virtual ranks on ONE GPU (engine.LocalWorld) against the oracle: the multi-rank exchange -- fan-out table,
# per-(layer, origin) slots, flag / count arithmetic, flag-waiting batched reconstruct -- checked on a 1-GPU box
# ---------------------------------------------------------------------------------------------------------
def _world_data(world, n, c, steps, layers, seed):
    g = torch.Generator().manual_seed(seed)
    x0 = [[[torch.randn(n, c, generator=g) for _ in range(world)] for _ in range(2)] for _ in range(layers)]
    return [[[[(0.97 ** t * x0[l][j][r] + 0.2 * torch.randn(n, c, generator=g)).half() for r in range(world)]
              for j in range(2)] for l in range(layers)] for t in range(steps)]  # [t][layer][kv][rank]


@pytest.mark.parametrize("mode", ["patch", "ring"])
@pytest.mark.parametrize("codec", ["binary", "int2"])
@pytest.mark.parametrize("world,n,c", [(4, 144, 3072), (8, 72, 1152), (2, 288, 1536), (2, 290, 1536)])
def test_virtual_ranks_exchange_vs_oracle(world, n, c, codec, mode):
    """4 steps x 2 layers of the W-rank exchange through cf_sign_compress_put into every (virtual) rank's
    receive region.  Per step, on EVERY receiver: the payload that arrived from every origin carries the
    oracle's codes (bit-exact) and scales (<= 1 ulp), and the receiver's reconstruction is bit-identical to the
    oracle's dequant of that payload against the cached base (oracle/check.py; main.py:398-419,
    ring.py:184-200).  All receivers end bit-identical; an independent oracle run of `all_gather_step` /
    `ring_step` over the same 4 steps stays within the 1-ulp-scale drift."""
    dev = _cuda()
    import compactfusion_b200 as cf
    from compactfusion_b200.engine import LocalWorld, PatchGatherEngine, RingExchangeEngine
    from oracle import check as ocheck
    from oracle.state import OracleCompact, all_gather_step, ring_step
    T = cf.COMPACT_COMPRESS_TYPE
    ctype, layers, steps = T(codec), 2, 4
    data = _world_data(world, n, c, steps, layers, seed=world * 1000 + n)
    lw = LocalWorld(world, layers, n, c, device=dev, engine_cls=RingExchangeEngine if mode == "ring" else PatchGatherEngine)
    run = lw.ring_all if mode == "ring" else lw.exchange_all
    oracle = [OracleCompact(residual=1, ef=True, fastpath=True) for _ in range(world)]
    for t in range(steps):
        ct = ctype if t >= 1 else T.WARMUP
        for l in range(layers):
            ks = [data[t][l][0][r].to(dev) for r in range(world)]
            vs = [data[t][l][1][r].to(dev) for r in range(world)]
            before = [(e.global_k[l].cpu(), e.global_v[l].cpu()) for e in lw.engines]
            run(l, ks, vs, ct)
            torch.cuda.synchronize()
            after = [(e.global_k[l].cpu(), e.global_v[l].cpu()) for e in lw.engines]
            for q in range(1, world):  # the error-feedback invariant: every rank holds the same caches
                assert torch.equal(after[q][0], after[0][0]) and torch.equal(after[q][1], after[0][1]), (t, l, q)
            # the independent oracle run (its own scales, its own drift)
            for j, sfx in enumerate("kv"):
                xs = [data[t][l][j][r] for r in range(world)]
                if mode == "ring":
                    ring_step(oracle, l, xs, ct.value, suffix=sfx)
                else:
                    all_gather_step(oracle, f"{l}-{sfx}", xs, ct.value)
            if t == 0:
                for r in range(world):
                    assert torch.equal(after[0][0][r * n:(r + 1) * n], data[0][l][0][r])
                continue
            for q, e in enumerate(lw.engines):  # every receiver against the oracle
                for j in range(2):
                    sh = lambda g, r: g[r * n:(r + 1) * n]  # noqa: E731
                    verdict = ocheck.check_exchange(
                        codec, [data[t][l][j][r] for r in range(world)], [sh(before[q][j], r) for r in range(world)],
                        [e.slot_bytes(l, r, ctype, j).cpu().numpy() for r in range(world)],
                        [sh(after[q][j], r) for r in range(world)])
                    assert verdict["ok"], (t, l, q, j, verdict)
        assert not any(e.p2p_error() for e in lw.engines)
    # after 4 steps the independent oracle caches and the GPU caches agree up to the scale-ulp drift
    for l in range(layers):
        for j, sfx in enumerate("kv"):
            for r in range(world):
                key = f"{l}-{r}-{sfx}" if mode == "ring" else f"{l}-{sfx}-{r}"
                got = (lw.engines[0].global_k[l] if j == 0 else lw.engines[0].global_v[l])[r * n:(r + 1) * n]
                assert rel_l2(got, oracle[0].base[key]) < 2e-3, (l, sfx, r)


def test_virtual_ranks_graph_and_publish_modes(monkeypatch):
    """The W = 4 virtual-rank step as ONE replayed CUDA graph, and with the alternative flag publication
    (CF_PUBLISH_MODE=1: one CTA stores the new count instead of every CTA adding 1): bit-identical caches."""
    dev = _cuda()
    import compactfusion_b200 as cf
    from compactfusion_b200.engine import LocalWorld
    T = cf.COMPACT_COMPRESS_TYPE
    world, n, c, layers, steps = 4, 144, 3072, 3, 3
    data = _world_data(world, n, c, steps, layers, seed=77)

    def run(publish_mode, graph):
        monkeypatch.setenv("CF_PUBLISH_MODE", str(publish_mode))
        lw = LocalWorld(world, layers, n, c, device=dev)
        ks = [[data[0][l][0][r].to(dev) for r in range(world)] for l in range(layers)]
        vs = [[data[0][l][1][r].to(dev) for r in range(world)] for l in range(layers)]
        for l in range(layers):
            lw.exchange_all(l, ks[l], vs[l], T.WARMUP)
        g = None
        for t in range(1, steps):
            for l in range(layers):
                for r in range(world):
                    ks[l][r].copy_(data[t][l][0][r])
                    vs[l][r].copy_(data[t][l][1][r])
            if graph and t == 2:
                if g is None:
                    s = torch.cuda.Stream()
                    s.wait_stream(torch.cuda.current_stream())
                    g = torch.cuda.CUDAGraph()
                    for e in lw.engines:
                        e._ptr_cache.clear()
                    with torch.cuda.graph(g):
                        for l in range(layers):
                            lw.exchange_all(l, ks[l], vs[l], T.BINARY)
                g.replay()
            else:
                for l in range(layers):
                    lw.exchange_all(l, ks[l], vs[l], T.BINARY)
        torch.cuda.synchronize()
        assert not any(e.p2p_error() for e in lw.engines)
        return [x.clone() for x in lw.engines[world - 1].global_k + lw.engines[0].global_v]

    ref = run(0, False)
    for mode, graph in ((1, False), (0, True)):
        for a, b in zip(run(mode, graph), ref):
            assert torch.equal(a, b), (mode, graph)


@pytest.mark.parametrize("codec,rank", [("low-rank", 8), ("low-rank-int4", 32), ("low-rank", 32)])
@pytest.mark.parametrize("world", [1, 2])
def test_engine_carries_lowrank_payloads(codec, rank, world):
    """LOW_RANK / LOW_RANK_Q (CogVideoX's preset, examples/configs.py:87-97) through the engines: payload sizes
    are the reference's wire formats (slowpath.py:62-75), all (virtual) ranks hold bit-identical caches after every
    step, and the error-feedback reconstruction tracks the input as well as the per-call plugin path does
    (the random start of the projector differs per call, so the comparison is on quality, not bits)."""
    dev = _cuda()
    import compactfusion_b200 as cf
    from compactfusion_b200.engine import LocalWorld, PatchGatherEngine
    T = cf.COMPACT_COMPRESS_TYPE
    ctype, n, c, layers, steps = T(codec), 288, 1536, 2, 4
    g = torch.Generator().manual_seed(rank)
    # low-rank-plus-noise residuals, so that a rank-r code has something to find
    def series(seed):
        gg = torch.Generator().manual_seed(seed)
        x0 = torch.randn(n, c, generator=gg)
        out = [x0.half()]
        for _ in range(steps - 1):
            x0 = x0 + torch.randn(n, 6, generator=gg) @ torch.randn(6, c, generator=gg) * 0.05 + 0.01 * torch.randn(n, c, generator=gg)
            out.append(x0.half())
        return out
    data = [[[series(100 * l + 10 * j + r) for r in range(world)] for j in range(2)] for l in range(layers)]
    if world == 1:
        engines = [PatchGatherEngine(layers, n, c, device=dev, comp_rank=rank)]
        run = lambda l, ks, vs, ct: [engines[0].exchange(l, ks[0], vs[0], ct)]  # noqa: E731
    else:
        lw = LocalWorld(world, layers, n, c, device=dev, comp_rank=rank)
        engines, run = lw.engines, lw.exchange_all
    want_numel = rank * (n + c) if ctype == T.LOW_RANK else rank * (n + c) // 4 + 4 * rank
    assert engines[0]._numel(ctype) == want_numel
    # the per-call plugin path on the same inputs (rank 0's K of layer 0)
    cf.compact_init(cf.CompactConfig(enabled=True, compress_func=lambda l, s: ctype if s >= 1 else T.WARMUP,
                                     comp_rank=rank, residual=1, ef=True))
    for t in range(steps):
        ct = ctype if t >= 1 else T.WARMUP
        for l in range(layers):
            ks = [data[l][0][r][t].to(dev) for r in range(world)]
            vs = [data[l][1][r][t].to(dev) for r in range(world)]
            before = engines[0].global_k[l].clone()
            run(l, ks, vs, ct)
            torch.cuda.synchronize()
            for e in engines[1:]:
                assert torch.equal(e.global_k[l], engines[0].global_k[l]) and torch.equal(e.global_v[l], engines[0].global_v[l])
            if t >= 1:
                for r in range(world):
                    got, was = engines[0].global_k[l][r * n:(r + 1) * n], before[r * n:(r + 1) * n]
                    assert rel_l2(got, ks[r]) < 0.5 * rel_l2(was, ks[r]) + 1e-3, (t, l, r, rel_l2(got, ks[r]), rel_l2(was, ks[r]))
        x = data[0][0][0][t].to(dev).view(1, n, 12, c // 12)
        p = cf.compact_compress("0-0-k", x, ct, update_cache=True)
        ref = cf.compact_decompress("9-0-k", p, ct, x.shape, update_cache=True).view(n, c)
        if t >= 1:
            assert p.numel() == want_numel, "wire size differs from the plugin path's payload"
            mine = engines[0].global_k[0][:n]
            assert rel_l2(mine, x.view(n, c)) < 1.25 * rel_l2(ref, x.view(n, c)) + 2e-3
    assert not any(e.p2p_error() for e in engines)
