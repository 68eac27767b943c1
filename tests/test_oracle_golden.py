"""Pins the CPU oracle (oracle/codecs.py, oracle/state.py) against golden vectors
produced by the reference itself (oracle/make_goldens.py, run where
/root/reference exists).  Integer codes and fp16 values must match bit-exactly:
both sides are eager torch CPU arithmetic of the same expressions.
"""
import numpy as np
import pytest
import torch

from conftest import LRQ_WIRE_CASES, CODEC_CASES, assert_bits_equal, h16, rel_l2
from oracle import codecs
from oracle.state import OracleCompact


def _inputs(g, name):
    x, base = h16(g[f"{name}/x"]), h16(g[f"{name}/base"])
    return x, base, x - base


@pytest.mark.parametrize("name", CODEC_CASES)
def test_binary_matches_sim_binary(golden_codecs, name):
    x, base, d = _inputs(golden_codecs, name)
    ref = h16(golden_codecs[f"{name}/sim_binary"])
    assert_bits_equal(codecs.sim_binary(d), ref, "sim_binary")
    # fastpath restatement: base + deq == base + sim_binary(delta), and deq alone == sim_binary
    packed, u, v, nb = codecs.binary_quant(x, base, True)
    assert_bits_equal(codecs.binary_dequant(packed, u, v, None), ref, "binary deq")
    assert_bits_equal(nb, base + ref, "binary new_base")
    assert_bits_equal(codecs.binary_dequant(packed, u, v, base), nb, "sender/receiver identity")
    # packing convention: bit c%8 of byte c//8, 1 <=> delta >= 0
    n, c = d.shape
    bits = (d >= 0).numpy()
    for (r, col) in [(0, 0), (n - 1, c - 1), (n // 2, 13 % c), (3 % n, 8 % c)]:
        assert ((packed[r, col // 8] >> (col % 8)) & 1) == int(bits[r, col])


@pytest.mark.parametrize("name", CODEC_CASES)
def test_int2_matches_reference(golden_codecs, name):
    g = golden_codecs
    x, base, d = _inputs(g, name)
    packed, tok, chan, nb = codecs.int2_quant(x, base, True)
    assert np.array_equal(packed, g[f"{name}/int2_packed"])
    assert_bits_equal(tok, h16(g[f"{name}/int2_tok"]), "tok")
    assert_bits_equal(chan.t(), h16(g[f"{name}/int2_chan"]), "chan")
    deq = codecs.int2_dequant(packed, tok, chan, None)
    assert_bits_equal(deq, h16(g[f"{name}/int2_deq"]), "deq")
    assert_bits_equal(nb, base + deq, "new_base")
    assert_bits_equal(codecs.sim_int2(d), h16(g[f"{name}/sim_int2"]), "sim_int2")
    # the reference's own claim: sim == codec
    assert_bits_equal(deq, h16(g[f"{name}/sim_int2"]), "codec vs sim")
    assert_bits_equal(codecs.sim_int2_minmax(d), h16(g[f"{name}/sim_int2_minmax"]), "sim_int2_minmax")


@pytest.mark.parametrize("name", CODEC_CASES)
def test_int4_matches_reference(golden_codecs, name):
    g = golden_codecs
    _, _, d = _inputs(g, name)
    packed, scale, mn = codecs.int4_quantize(d)
    assert np.array_equal(packed, g[f"{name}/int4_packed"])
    assert_bits_equal(scale, h16(g[f"{name}/int4_scale"]), "scale")
    assert_bits_equal(mn, h16(g[f"{name}/int4_min"]), "min")
    assert_bits_equal(codecs.int4_dequantize(packed, scale, mn), h16(g[f"{name}/int4_deq"]), "deq")
    assert_bits_equal(codecs.sim_int4(d, 0), h16(g[f"{name}/sim_int4_d0"]), "sim d0")
    assert_bits_equal(codecs.sim_int4(d, 1), h16(g[f"{name}/sim_int4_d1"]), "sim d1")
    assert_bits_equal(codecs.int4_dequantize(packed, scale, mn), h16(g[f"{name}/sim_int4_d0"]), "codec vs sim")


@pytest.mark.parametrize("name", CODEC_CASES)
def test_int8_matches_reference(golden_codecs, name):
    g = golden_codecs
    _, _, d = _inputs(g, name)
    q, scale, zp = codecs.int8_quantize(d)
    assert np.array_equal(q, g[f"{name}/int8_q"])
    assert_bits_equal(scale, h16(g[f"{name}/int8_scale"]), "scale")
    assert np.array_equal(zp.numpy(), g[f"{name}/int8_zp"])
    assert_bits_equal(codecs.int8_dequantize(q, scale, zp), h16(g[f"{name}/int8_deq"]), "deq")


@pytest.mark.parametrize("m", [2, 4, 8, 16])
def test_topk_matches_sim_topk(golden_codecs, m):
    x = h16(golden_codecs["topk/x"])
    ref = h16(golden_codecs[f"topk/sim_m{m}"])
    assert_bits_equal(codecs.sim_topk(x, m), ref, "sim_topk")
    val, idx = codecs.topk_compress(x, m)
    assert val.shape == (4, 1024 // m) and idx.shape == (4, 512 // m)
    assert_bits_equal(codecs.topk_decompress(val, idx, m), ref, "codec round trip")


def test_topk_tie_break_lowest_index():
    x = torch.zeros(1, 1024, dtype=torch.half)
    x[0, 2] = 1.0
    x[0, 3] = -1.0  # same magnitude: index 2 wins
    val, idx = codecs.topk_compress(x, 4)
    assert float(val[0, 0]) == 1.0 and (idx[0, 0] >> 4) == 2


def test_subspace_iter_matches_reference(golden_codecs):
    g = golden_codecs
    a = h16(g["lowrank/a"])
    q0 = torch.from_numpy(g["lowrank/q0"])
    u, v, q = codecs.subspace_iter(a, 4, 2, init_q=q0)
    assert_bits_equal(u, h16(g["lowrank/u"]), "U")
    assert_bits_equal(v, h16(g["lowrank/v"]), "V")
    assert_bits_equal(q, h16(g["lowrank/q"]), "Q")
    torch.manual_seed(123)
    u2, v2, _ = codecs.subspace_iter(a, 8, 2)
    assert_bits_equal((u2.float() @ v2.float()).half(), h16(g["lowrank/seed123_r8_uv"]), "seeded U@V")
    # and it is a good approximation of a nearly rank-6 matrix
    assert rel_l2(u2.float() @ v2.float(), a) < 0.05


@pytest.mark.parametrize("name,ctype,kw", [
    ("low_rank_r8", "low-rank", dict(rank=8)),
    ("low_rank_q_r4", "low-rank-int4", dict(rank=4)),
])
def test_slowpath_payloads(golden_slowpath, name, ctype, kw):
    g = golden_slowpath
    x = h16(g["x"])
    torch.manual_seed(123)
    p = codecs.slowpath_compress(x, ctype, **kw)
    assert_bits_equal(p, h16(g[f"{name}/payload"]), "payload")
    assert_bits_equal(codecs.slowpath_decompress(p, x.shape, ctype, **kw), h16(g[f"{name}/recon"]), "recon")
    torch.manual_seed(123)
    assert_bits_equal(codecs.sim_compress(x, ctype, **kw), h16(g[f"{name}/sim"]), "sim")


@pytest.mark.parametrize("name,n,c,r", LRQ_WIRE_CASES)
def test_lowrank_q_wire_codec_on_fixed_factors(golden_lowrank_q_wire, name, n, c, r):
    """slowpath.py:69-75 / :156-164 on fixed U, V: the oracle's int4 assembles the reference's payload bit for bit
    and decodes it to the reference's reconstruction (fp16 matmul: fp32 accumulation, one rounding)."""
    g = golden_lowrank_q_wire
    u, v = h16(g[f"{name}/u"]), h16(g[f"{name}/v"])
    parts = []
    for t in (u, v.t().contiguous()):
        q, s, m = codecs.int4_quantize(t)
        parts += [torch.from_numpy(q.copy()).contiguous().view(-1).view(torch.half), s.reshape(-1), m.reshape(-1)]
    payload = torch.cat(parts)
    want = h16(g[f"{name}/payload"])
    assert_bits_equal(payload, want, "LOW_RANK_Q payload from fixed factors")
    rec = codecs.slowpath_decompress(want, (n, c), "low-rank-int4", rank=r)
    ref = h16(g[f"{name}/recon"])
    assert rel_l2(rec, ref) < 1e-3 and float((rec.float() - ref.float()).abs().max()) <= 2e-2


STATE_FLAVOURS = {
    "sim_int4_r1_ef": (dict(residual=1, ef=True, simulate=True), "int4"),
    "sim_binary_r1_ef": (dict(residual=1, ef=True, simulate=True), "binary"),
    "sim_int2_r2_ef": (dict(residual=2, ef=True, simulate=True, delta_decay_factor=0.5), "int2"),
    "sim_int4_r1_noef": (dict(residual=1, ef=False, simulate=True), "int4"),
    "sim_int4_r0": (dict(residual=0, ef=False, simulate=True), "int4"),
    "real_lowrank4_r1_ef": (dict(residual=1, ef=True, simulate=False, comp_rank=4), "low-rank"),
}


@pytest.mark.parametrize("name", list(STATE_FLAVOURS))
def test_state_machine_matches_reference(golden_state, name):
    """compact_compress / compact_decompress (main.py:169-388) over 6 steps."""
    g = golden_state
    kw, ctype = STATE_FLAVOURS[name]
    n, c, steps = 32, 128, 6
    shape4 = (1, n, 4, c // 4)
    warm = 2 if kw["residual"] == 2 else 1
    sender, receiver = OracleCompact(**kw), OracleCompact(**kw)
    for t in range(steps):
        x = h16(g["xs"][t]).view(shape4)
        ct = ctype if t >= warm else "warmup"
        torch.manual_seed(1000 + t)
        comp = sender.compress("0-0-k", x, ct, update_cache=True)
        rec = receiver.decompress("0-0-k", comp, ct, shape4, update_cache=True)
        assert_bits_equal(comp.reshape(-1), h16(g[f"{name}/comp{t}"]), f"comp{t}")
        assert_bits_equal(rec.reshape(-1), h16(g[f"{name}/recon{t}"]), f"recon{t}")
        if kw["residual"] != 0:
            assert_bits_equal(sender.base["0-0-k"].reshape(-1), h16(g[f"{name}/send_base{t}"]), f"send_base{t}")
            assert_bits_equal(receiver.base["0-0-k"].reshape(-1), h16(g[f"{name}/recv_base{t}"]), f"recv_base{t}")
        if f"{name}/send_dbase{t}" in g:
            assert_bits_equal(sender.delta_base["0-0-k"].reshape(-1), h16(g[f"{name}/send_dbase{t}"]), f"send_dbase{t}")
    if kw.get("ef"):
        # error-feedback invariant: sender and receiver caches are bit-identical
        assert_bits_equal(sender.base["0-0-k"], receiver.base["0-0-k"], "EF invariant")


def test_lowrank_product_does_not_depend_on_orthonormalising_the_random_start():
    """compactfusion_b200.compress_lowrank._init_q skips the reference's library QR of the random start
    (compress_lowrank.py:40-42): with the reference's own algorithm (the oracle restatement) U V from a raw
    Gaussian Q0 equals U V from qr(Q0) to fp16 output rounding, with and without low-rank structure."""
    from oracle import codecs
    g = torch.Generator().manual_seed(0)
    n, c, r = 272, 384, 8
    low = torch.randn(n, r, generator=g) @ torch.randn(r, c, generator=g)
    for a in ((low + 0.3 * torch.randn(n, c, generator=g)).half(), torch.randn(n, c, generator=g).half()):
        q0 = torch.randn(c, r, generator=g)
        qo, _ = torch.linalg.qr(q0)
        u1, v1, _ = codecs.subspace_iter(a, r, 2, init_q=q0)
        u2, v2, _ = codecs.subspace_iter(a, r, 2, init_q=qo)
        p1, p2 = u1.float() @ v1.float(), u2.float() @ v2.float()
        assert float((p1 - p2).norm() / p2.norm()) < 2e-4
