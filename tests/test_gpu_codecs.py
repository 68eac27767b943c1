"""Parity of the sm_100a kernels (called through the C ABI via the Python mirror) against the
CPU oracle and the reference-generated golden vectors.  Needs a GPU: run with -m gpu.

Bars (SURVEY.md section 7 "hard parts" 2, section 8c):
  * sign bits, INT4 / INT8 / top-k codes and everything computed from exact statistics
    (min/max): bit-exact;
  * BINARY / INT2 scales are means whose fp32 summation order differs between any two
    implementations: at most 1 fp16 ulp apart, rel-L2 < 1e-3 (the reference's own bar,
    tests/compact/compress_fastpath_test.py:86-87);
  * everything downstream of the scales is bit-exact GIVEN IDENTICAL SCALE TENSORS (we feed
    the kernel's scales to the oracle); INT2 codes end to end >= 99.9 % byte-equal
    (compress_fastpath_test.py:134).
"""
import numpy as np
import pytest
import torch

from conftest import CODEC_CASES, assert_bits_equal, bits16, h16, rel_l2
from oracle import codecs as oc

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def ulp_diff(a: torch.Tensor, b: torch.Tensor) -> int:
    """max distance in fp16 ulps between two tensors of non-negative-ish scales"""
    x = bits16(a).astype(np.int32)
    y = bits16(b).astype(np.int32)
    return int(np.abs(x - y).max()) if x.size else 0


def make_xb(n, c, seed, scale_base=0.1):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, c, generator=g).half()
    base = (torch.randn(n, c, generator=g) * scale_base).half()
    return x, base


SHAPES = [(2048, 1024), (8192, 512), (1088, 3072), (130, 64), (77, 1152), (64, 8192), (16, 16384),
          (4608, 3072), (5001, 1152), (3000, 1536), (2050, 6144), (9, 3072)]


@pytest.fixture(params=["tma", "legacy"])
def kernel_path(request, monkeypatch):
    """Both implementations of the streaming kernels must meet the same parity bars: the
    bulk-async (TMA) pipelined ones (default) and the register-staged ones (fallback for
    unaligned / very wide shapes, forced here with CF_LEGACY_KERNELS=1)."""
    monkeypatch.setenv("CF_LEGACY_KERNELS", "1" if request.param == "legacy" else "0")
    return request.param


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("update_cache", [True, False])
def test_binary_fastpath_vs_oracle(shape, update_cache, kernel_path):
    dev = _cuda()
    from compactfusion_b200.fastpath import binary_dequant_fastpath, binary_quant_fastpath
    n, c = shape
    x, base = make_xb(n, c, 42 + n)
    packed, u, v, nb = binary_quant_fastpath(x.to(dev), base.to(dev), -1, update_cache)
    assert packed.shape == (n, c // 8) and u.shape == (n, 1) and v.shape == (c, 1)
    o_packed, o_u, o_v, _ = oc.binary_quant(x, base, False)
    # codes: bit-exact
    assert np.array_equal(packed.cpu().numpy(), o_packed), "packed sign bits differ"
    # scales: <= 1 ulp, rel-L2 1e-3
    assert ulp_diff(u.cpu(), o_u) <= 1 and ulp_diff(v.cpu(), o_v) <= 1
    assert rel_l2(u, o_u) < 1e-3 and rel_l2(v, o_v) < 1e-3
    # elementwise stage: bit-exact given identical scales
    recon = binary_dequant_fastpath(packed, u, v, base.to(dev))
    o_recon = oc.binary_dequant(o_packed, u.cpu(), v.cpu(), base)
    assert_bits_equal(recon, o_recon, "recon")
    if update_cache:
        assert_bits_equal(nb, o_recon, "new_base (sender == receiver)")
    else:
        assert nb is None


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("update_cache", [True, False])
def test_int2_fastpath_vs_oracle(shape, update_cache, kernel_path):
    dev = _cuda()
    from compactfusion_b200 import _native as nv
    from compactfusion_b200.fastpath import int2_dequant_fastpath, int2_quant_fastpath
    n, c = shape
    x, base = make_xb(n, c, 43 + n)
    xd, bd = x.to(dev), base.to(dev)
    packed, tok, chan, nb = int2_quant_fastpath(xd, bd, update_cache)
    assert packed.shape == (n, c // 4) and tok.shape == (n, 1) and chan.shape == (c, 1)
    o_packed, o_tok, o_chan, o_nb = oc.int2_quant(x, base, update_cache)
    assert ulp_diff(tok.cpu(), o_tok) <= 1 and ulp_diff(chan.cpu(), o_chan) <= 1
    mismatch = float((packed.cpu().numpy() != o_packed).mean())
    assert mismatch <= 1e-3, f"INT2 packed bytes mismatch ratio {mismatch}"
    # bit-exact given identical scale tensors: oracle run with the kernel's scales ...
    s_packed, _, _, s_nb = oc.int2_quant(x, base, update_cache, scales=(tok.cpu(), chan.cpu()))
    assert np.array_equal(packed.cpu().numpy(), s_packed), "codes differ given identical scales"
    if update_cache:
        assert_bits_equal(nb, s_nb, "new_base")
    # ... and kernel run with the oracle's scales
    p2 = torch.empty_like(packed)
    nb2 = torch.empty_like(xd)
    tok_d, chan_d = o_tok.to(dev), o_chan.to(dev)  # keep alive: the call only sees raw pointers
    rc = nv.lib().cf_int2_encode_with_scales(xd.data_ptr(), bd.data_ptr(), tok_d.data_ptr(), chan_d.data_ptr(),
                                             nb2.data_ptr(), p2.data_ptr(), n, c, nv.stream_ptr())
    nv.check(rc, "cf_int2_encode_with_scales")
    assert np.array_equal(p2.cpu().numpy(), o_packed)
    assert_bits_equal(nb2, oc.int2_quant(x, base, True)[3], "new_base with oracle scales")
    recon = int2_dequant_fastpath(packed, tok, chan, bd)
    assert_bits_equal(recon, oc.int2_dequant(packed.cpu().numpy(), tok.cpu(), chan.cpu(), base), "recon")


@pytest.mark.parametrize("shape", [(2048, 1024), (512, 4096), (1088, 3072), (130, 64), (64, 8192)])
def test_int4_int8_bit_exact(shape, kernel_path):
    dev = _cuda()
    from compactfusion_b200.compress_quantize import (dequantize_int4, dequantize_int8, quantize_int4, quantize_int8,
                                                      sim_int4)
    n, c = shape
    x, base = make_xb(n, c, 44 + n)
    d = x - base
    dd = d.to(dev)
    q, s, m = quantize_int4(dd)
    oq, os_, om = oc.int4_quantize(d)
    assert_bits_equal(s, os_, "int4 scale")
    assert_bits_equal(m, om, "int4 min")
    assert np.array_equal(q.cpu().numpy(), oq), "int4 codes"
    assert_bits_equal(dequantize_int4(q, s, m), oc.int4_dequantize(oq, os_, om), "int4 deq")
    assert_bits_equal(sim_int4(dd, 0), oc.sim_int4(d, 0), "sim_int4 dim0")
    if n <= 2048:
        assert_bits_equal(sim_int4(dd, 1), oc.sim_int4(d, 1), "sim_int4 dim1")
    q8, s8, z8 = quantize_int8(dd)
    oq8, os8, oz8 = oc.int8_quantize(d)
    assert_bits_equal(s8, os8, "int8 scale")
    assert np.array_equal(z8.cpu().numpy(), oz8.numpy()), "int8 zero point"
    assert np.array_equal(q8.cpu().numpy(), oq8), "int8 codes"
    assert_bits_equal(dequantize_int8(q8, s8, z8), oc.int8_dequantize(oq8, os8, oz8), "int8 deq")


def test_int4_int8_exact_quotient_edge_cases(kernel_path):
    """The encode kernels replace the per-element IEEE division by a reciprocal multiply plus an
    exactness check (csrc/cf_minmax_codecs.cu: quot_for_rn16).  Columns built to sit ON the rounding
    ties, with tiny / huge / zero scales and subnormal quotients must still give the oracle's codes."""
    dev = _cuda()
    from compactfusion_b200.compress_quantize import dequantize_int4, dequantize_int8, quantize_int4, quantize_int8
    n, c = 4096, 512
    g = torch.Generator().manual_seed(2024)
    d = torch.randn(n, c, generator=g) * torch.exp(3.0 * torch.randn(1, c, generator=g))
    k = torch.randint(0, 16, (n, 64), generator=g).float()
    lo = torch.randn(1, 64, generator=g)
    step = torch.rand(1, 64, generator=g) * 0.3 + 1e-3
    d[:, :64] = lo + (k + 0.5) * step                 # exact ties of the 16-level grid
    d[:, 64:96] = (k[:, :32] + 0.5) / 256.0 * 7.0     # ties of the 256-level grid
    d[:, 96:112] *= 1e-6                               # subnormal fp16 values
    d[:, 112:120] = (torch.randn(n, 8, generator=g) * 9000.0).clamp(-30000, 30000)  # large, finite max - min
    d[:, 120] = 0.0                                    # constant columns (zero scale)
    d[:, 121] = 1.5
    d[:, 122] = 2.25
    d[0, 123] = 30000.0
    d[1, 123] = -30000.0
    d = d.clamp(-30000, 30000).half()
    dd = d.to(dev)
    q, s, m = quantize_int4(dd)
    oq, os_, om = oc.int4_quantize(d)
    assert_bits_equal(s, os_, "int4 scale")
    assert_bits_equal(m, om, "int4 min")
    live = (os_.float().view(-1) != 0).numpy()         # zero-scale columns: the reference's cast is undefined
    assert np.array_equal(q.cpu().numpy()[:, live], oq[:, live]), "int4 codes"
    got, want = dequantize_int4(q, s, m).cpu(), oc.int4_dequantize(oq, os_, om)
    assert_bits_equal(got[:, live], want[:, live], "int4 deq")
    # (a constant column has a zero scale: we define code 0, so it reconstructs exactly)
    assert_bits_equal(got[:, 120:123], d[:, 120:123], "int4 constant columns reconstruct exactly")
    q8, s8, z8 = quantize_int8(dd)
    oq8, os8, oz8 = oc.int8_quantize(d)
    assert_bits_equal(s8, os8, "int8 scale")
    assert np.array_equal(z8.cpu().numpy(), oz8.numpy()), "int8 zero point"
    assert np.array_equal(q8.cpu().numpy(), oq8), "int8 codes"
    assert_bits_equal(dequantize_int8(q8, s8, z8), oc.int8_dequantize(oq8, os8, oz8), "int8 deq")


def test_int4_error_feedback_round_trip_baseline_config(kernel_path):
    """BASELINE config 0 at full size (4096 x 3072, INT4 residual + error feedback): the fused
    sender update equals what the receiver reconstructs from the wire bytes, codes match the
    oracle on a 256-column slice, and the EF residual stays bounded over the steps."""
    dev = _cuda()
    from compactfusion_b200 import _native as nv
    from compactfusion_b200.compress_quantize import _minmax_compress, dequantize_int4
    n, c = 4096, 3072
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(n, c, generator=g, device=dev)
    base = x.half()
    errs = []
    for t in range(1, 5):
        x = x + 0.05 * torch.randn(n, c, generator=g, device=dev)
        xh = x.half()
        codes, s, m, new_base = _minmax_compress(nv.CODEC_INT4, xh, base, want_recon=True)
        recv = base + dequantize_int4(codes, s, m)
        assert torch.equal(recv, new_base), "receiver != sender after step %d" % t
        d = (xh - base)[:, :256].cpu()
        oq, os_, om = oc.int4_quantize(d)
        assert np.array_equal(codes[:, :256].cpu().numpy(), oq)
        errs.append(rel_l2(new_base, xh))
        base = new_base
    assert max(errs) < 0.02 and errs[-1] < 1.5 * errs[0] + 1e-3, errs


def test_int4_fused_residual_matches_composition(kernel_path):
    """cf_int4_compress(x, base, new_base) == base + dequant(quant(x - base)) bit-exactly."""
    dev = _cuda()
    from compactfusion_b200 import _native as nv
    from compactfusion_b200.compress_quantize import _minmax_compress
    x, base = make_xb(512, 1536, 7)
    codes, s, m, recon = _minmax_compress(nv.CODEC_INT4, x.to(dev), base.to(dev), want_recon=True)
    d = x - base
    oq, os_, om = oc.int4_quantize(d)
    assert np.array_equal(codes.cpu().numpy(), oq)
    assert_bits_equal(recon, base + oc.int4_dequantize(oq, os_, om), "fused int4 EF")


@pytest.mark.parametrize("m", [2, 4, 8, 16])
@pytest.mark.parametrize("shape", [(1024, 2048), (256, 8192), (3, 1024)])
def test_topk_bit_exact(m, shape):
    dev = _cuda()
    from compactfusion_b200.compress_topk import sim_topk, topk_compress, topk_decompress, topk_sparsify
    a, w = shape
    g = torch.Generator().manual_seed(42 + m)
    x = torch.randn(a, w, generator=g).half()
    rows = x.view(-1, 1024)
    val, idx = topk_compress(rows.to(dev), m)
    oval, oidx = oc.topk_compress(rows, m)
    assert_bits_equal(val, oval, "topk values")
    assert np.array_equal(idx.cpu().numpy(), oidx), "topk indices"
    ref = oc.sim_topk(x, m)
    assert_bits_equal(topk_decompress(val, idx, m).view(a, w), ref, "decompress")
    assert_bits_equal(topk_sparsify(rows.to(dev), m).view(a, w), ref, "sparsify")
    assert_bits_equal(sim_topk(x.to(dev), m), ref, "sim_topk")


def test_topk_ties_pick_lowest_index():
    dev = _cuda()
    from compactfusion_b200.compress_topk import topk_compress
    x = torch.zeros(1, 1024, dtype=torch.half)
    x[0, 2], x[0, 3] = 1.0, -1.0
    x[0, 5], x[0, 6] = -2.0, 2.0
    val, idx = topk_compress(x.to(dev), 4)
    assert float(val[0, 0]) == 1.0 and float(val[0, 1]) == -2.0
    assert int(idx[0, 0]) == ((2 << 4) | 1)


@pytest.mark.parametrize("name", CODEC_CASES)
def test_against_reference_goldens(golden_codecs, name, kernel_path):
    """CUDA kernels on the committed inputs vs what the reference itself produced."""
    dev = _cuda()
    from compactfusion_b200.compress_quantize import (quantize_int2, quantize_int4, quantize_int8, sim_binary,
                                                      sim_int2, sim_int4, dequantize_int4, dequantize_int8)
    g = golden_codecs
    x, base = h16(g[f"{name}/x"]), h16(g[f"{name}/base"])
    d = (x - base)
    dd = d.to(dev)
    # exact codecs
    q, s, m = quantize_int4(dd)
    assert np.array_equal(q.cpu().numpy(), g[f"{name}/int4_packed"])
    assert_bits_equal(s, h16(g[f"{name}/int4_scale"]), "int4 scale")
    assert_bits_equal(m, h16(g[f"{name}/int4_min"]), "int4 min")
    assert_bits_equal(dequantize_int4(q, s, m), h16(g[f"{name}/int4_deq"]), "int4 deq")
    assert_bits_equal(sim_int4(dd, 0), h16(g[f"{name}/sim_int4_d0"]), "sim_int4")
    q8, s8, z8 = quantize_int8(dd)
    assert np.array_equal(q8.cpu().numpy(), g[f"{name}/int8_q"])
    assert np.array_equal(z8.cpu().numpy(), g[f"{name}/int8_zp"])
    assert_bits_equal(dequantize_int8(q8, s8, z8), h16(g[f"{name}/int8_deq"]), "int8 deq")
    # mean-scale codecs: signs exact, values within the reference's own tolerances
    sb = sim_binary(dd, rank=-1).cpu()
    ref_sb = h16(g[f"{name}/sim_binary"])
    assert torch.equal(torch.signbit(sb), torch.signbit(ref_sb)), "binary signs"
    assert rel_l2(sb, ref_sb) < 1e-3
    p2, chan, tok = quantize_int2(dd)
    assert float((p2.cpu().numpy() != g[f"{name}/int2_packed"]).mean()) <= 1e-3
    assert ulp_diff(chan.cpu(), h16(g[f"{name}/int2_chan"])) <= 1
    assert ulp_diff(tok.cpu(), h16(g[f"{name}/int2_tok"])) <= 1
    assert rel_l2(sim_int2(dd), h16(g[f"{name}/sim_int2"])) < 2e-2


def test_batched_equals_single():
    dev = _cuda()
    from compactfusion_b200 import _native as nv
    from compactfusion_b200.fastpath import binary_quant_fastpath, int2_quant_fastpath
    n, c, nb = 576, 3072, 5
    xs, bs = [], []
    for i in range(nb):
        x, b = make_xb(n, c, 100 + i)
        xs.append(x.to(dev))
        bs.append(b.to(dev))
    for codec, single, per_byte in ((nv.CODEC_BINARY, lambda x, b: binary_quant_fastpath(x, b, -1, True), 8),
                                    (nv.CODEC_INT2, lambda x, b: int2_quant_fastpath(x, b, True), 4)):
        ref = [single(x, b) for x, b in zip(xs, bs)]
        packed = [torch.empty((n, c // per_byte), dtype=torch.uint8, device=dev) for _ in range(nb)]
        u = [torch.empty((n, 1), dtype=torch.half, device=dev) for _ in range(nb)]
        v = [torch.empty((c, 1), dtype=torch.half, device=dev) for _ in range(nb)]
        newb = [torch.empty((n, c), dtype=torch.half, device=dev) for _ in range(nb)]
        ws = nv.workspace(nv.workspace_bytes(codec, n, c, 0, nb), dev)
        fn = nv.lib().cf_binary_compress_batched if codec == nv.CODEC_BINARY else nv.lib().cf_int2_compress_batched
        rc = fn(nb, nv.ptr_array(xs), nv.ptr_array(bs), nv.ptr_array(newb), nv.ptr_array(packed), nv.ptr_array(u),
                nv.ptr_array(v), n, c, ws.data_ptr(), ws.numel(), nv.stream_ptr())
        nv.check(rc, "compress_batched")
        for i in range(nb):
            assert torch.equal(packed[i], ref[i][0])
            # the batched launch splits rows over fewer CTAs: partial-sum order may differ by 1 ulp
            assert ulp_diff(u[i].cpu(), ref[i][1].cpu()) <= 1 and ulp_diff(v[i].cpu(), ref[i][2].cpu()) <= 1
        dfn = nv.lib().cf_binary_decompress_batched if codec == nv.CODEC_BINARY else nv.lib().cf_int2_decompress_batched
        rec = [torch.empty((n, c), dtype=torch.half, device=dev) for _ in range(nb)]
        rc = dfn(nb, nv.ptr_array(packed), nv.ptr_array(u), nv.ptr_array(v), nv.ptr_array(bs), nv.ptr_array(rec), n, c,
                 nv.stream_ptr())
        nv.check(rc, "decompress_batched")
        for i in range(nb):
            assert torch.equal(rec[i], newb[i]), "batched receiver != batched sender"


@pytest.mark.parametrize("codec", ["binary", "int2"])
def test_full_size_properties(codec):
    """BASELINE config-1 shape (4096 x 3072): size-independent properties instead of the oracle."""
    dev = _cuda()
    from compactfusion_b200.fastpath import (binary_dequant_fastpath, binary_quant_fastpath, int2_dequant_fastpath,
                                             int2_quant_fastpath)
    n, c = 4096, 3072
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(n, c, generator=g, device=dev).half()
    base = (x.float() + 0.3 * torch.randn(n, c, generator=g, device=dev)).half()
    if codec == "binary":
        packed, u, v, nb = binary_quant_fastpath(x, base, -1, True)
        recon = binary_dequant_fastpath(packed, u, v, base)
        bits = (x - base) >= 0
        got = ((packed.unsqueeze(-1) >> torch.arange(8, device=dev, dtype=torch.uint8)) & 1).view(n, c).bool()
        assert torch.equal(bits, got), "sign bits"
        step = (u.float() * v.float().t()).half()
        assert torch.equal(nb, base + torch.where(bits, step, -step)), "EF update"
    else:
        packed, u, v, nb = int2_quant_fastpath(x, base, True)
        recon = int2_dequant_fastpath(packed, u, v, base)
    assert torch.equal(recon, nb), "sender new_base != receiver recon"
    # error feedback contracts the residual: |x - new_base| < |x - base| in norm
    assert torch.norm(x.float() - nb.float()) < torch.norm(x.float() - base.float())
    # scales are positive and token scales average ~1
    assert float(v.min()) > 0 and abs(float(u.float().mean()) - 1.0) < 1e-2
    # determinism
    again = (binary_quant_fastpath(x, base, -1, True) if codec == "binary" else int2_quant_fastpath(x, base, True))
    assert torch.equal(again[0], packed) and torch.equal(again[1], u) and torch.equal(again[2], v)


def test_in_place_update_and_null_base():
    dev = _cuda()
    from compactfusion_b200 import _native as nv
    from compactfusion_b200.compress_quantize import _sign_compress, sim_binary
    x, base = make_xb(300, 1024, 5)
    xd, bd = x.to(dev), base.to(dev)
    ref = _sign_compress(nv.CODEC_BINARY, xd, bd, True)
    b2 = bd.clone()
    out = _sign_compress(nv.CODEC_BINARY, xd, b2, True, new_base=b2)  # new_base aliases base
    assert torch.equal(out[3], ref[3]) and out[3].data_ptr() == b2.data_ptr()
    # base == NULL: the bare dequantised tensor (sim_binary)
    sb = sim_binary(xd, rank=-1)
    p0, u0, v0, _ = oc.binary_quant(x, torch.zeros_like(x), False)
    assert torch.equal(torch.signbit(sb.cpu()), torch.signbit(oc.binary_dequant(p0, u0, v0, None)))


def test_cpu_tensor_fails_loudly():
    from compactfusion_b200 import _native as nv
    from compactfusion_b200.fastpath import binary_quant_fastpath
    x, base = make_xb(8, 64, 1)
    with pytest.raises(nv.NativeError):
        binary_quant_fastpath(x, base, -1, True)
