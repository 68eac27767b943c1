"""GPU tests of what landed after the round's GPU minutes were spent, other than the ring engine (that is
tests/test_gpu_zz_ring_engine.py): the fused LSE merge and one-pass error statistics kernels, statistics
logging through the plugin API, the plain-C client of the C ABI, the 4-level min/max simulation codec and the
int8-quantised cache.  Sorted late on purpose, so that `-x` reaches every earlier parity test first."""
import math
import os
import subprocess

import pytest
import torch

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
