"""LOW_RANK / LOW_RANK_Q on the GPU.  Only U @ V is comparable between implementations (QR sign
and basis conventions are arbitrary); tolerances are the reference's
(tests/compact/compress_slowpath_test.py: 5e-2 for LOW_RANK_Q, 1e-1 vs exact SVD)."""
import pytest
import torch

from conftest import LRQ_WIRE_CASES, assert_bits_equal, h16, rel_l2
from oracle import codecs as oc

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def test_subspace_iter_matches_reference_golden(golden_codecs):
    dev = _cuda()
    from compactfusion_b200.compress_lowrank import subspace_iter
    g = golden_codecs
    a = h16(g["lowrank/a"])
    q0 = torch.from_numpy(g["lowrank/q0"])
    u, v, q = subspace_iter(a.to(dev), 4, 2, init_q=q0.to(dev))
    assert u.shape == (96, 4) and v.shape == (4, 256) and q.shape == (256, 4)
    ref = h16(g["lowrank/u"]).float() @ h16(g["lowrank/v"]).float()
    got = u.float().cpu() @ v.float().cpu()
    assert rel_l2(got, ref) < 2e-3
    # U orthonormal, Q orthonormal
    eye = torch.eye(4)
    assert torch.allclose(u.float().cpu().t() @ u.float().cpu(), eye, atol=5e-3)
    assert torch.allclose(q.float().cpu().t() @ q.float().cpu(), eye, atol=5e-3)


@pytest.mark.parametrize("shape", [(128, 128), (32, 256), (512, 32)])
@pytest.mark.parametrize("rank", [1, 2])
def test_subspace_iter_converges_to_svd(shape, rank):
    """The reference's test_subspace_iter (compress_slowpath_test.py:190-220): 100 iterations
    on a nearly rank-r matrix vs exact truncated SVD, tol 1e-1."""
    dev = _cuda()
    from compactfusion_b200.compress_lowrank import subspace_iter, svd
    n, c = shape
    torch.manual_seed(42)
    for _ in range(3):
        low = torch.randn(n, rank, device=dev) @ torch.randn(rank, c, device=dev)
        a = (low + 0.01 * torch.norm(low) * torch.randn(n, c, device=dev) / (n * c) ** 0.5).half()
        pu, pv, _ = subspace_iter(a, rank, num_iters=100)
        su, sv = svd(a, rank)
        assert rel_l2(pu.float() @ pv.float(), su.float() @ sv.float()) < 1e-1


@pytest.mark.parametrize("shape", [(1088, 3072), (1024, 2048), (300, 520)])
@pytest.mark.parametrize("rank", [8, 12, 32, 64])
def test_lowrank_project_vs_oracle(shape, rank):
    """Fused residual projector on activation-like data vs the oracle's subspace_iter started
    from the same Q0."""
    dev = _cuda()
    from compactfusion_b200.compress_lowrank import lowrank_project, lowrank_reconstruct
    n, c = shape
    g = torch.Generator().manual_seed(5 + rank)
    sig = torch.exp(0.5 * torch.randn(c, generator=g))
    x = (torch.randn(n, 24, generator=g) @ torch.randn(24, c, generator=g) * 0.3 + torch.randn(n, c, generator=g) * sig).half()
    base = (x.float() * 0.9 + 0.1 * torch.randn(n, c, generator=g)).half()
    q0, _ = torch.linalg.qr(torch.randn(c, rank, generator=g))
    u, v, _ = lowrank_project(x.to(dev), base.to(dev), rank, 2, init_q=q0.to(dev))
    d = x - base
    ou, ov, _ = oc.subspace_iter(d, rank, 2, init_q=q0)
    ref = ou.float() @ ov.float()
    got = u.float().cpu() @ v.float().cpu()
    # same subspace iteration from the same start: products agree to fp16-factor rounding
    assert rel_l2(got, ref) < 1e-2, rel_l2(got, ref)
    # and the approximation quality is the same
    e_got, e_ref = rel_l2(got, d), rel_l2(ref, d)
    assert abs(e_got - e_ref) < 2e-3, (e_got, e_ref)
    # fused reconstruct == base + fp16(U V)
    rec = lowrank_reconstruct(u, v, base.to(dev))
    bare = lowrank_reconstruct(u, v, None)
    assert_bits_equal(rec, base.to(dev) + bare, "fused base add")
    assert rel_l2(bare, (u.float() @ v.float())) < 1e-3


@pytest.mark.parametrize("ctype,tol", [("low-rank", 2e-3), ("low-rank-int4", 5e-2)])
@pytest.mark.parametrize("shape", [(1024, 2048), (256, 8192)])
def test_slowpath_compress_decompress_vs_sim(ctype, tol, shape):
    """compress_slowpath_test.py:140-188 (only LOW_RANK_Q is enabled there; LOW_RANK added)."""
    dev = _cuda()
    from compactfusion_b200.slowpath import sim_compress, slowpath_compress, slowpath_decompress
    from compactfusion_b200.utils import COMPACT_COMPRESS_TYPE as T
    t = T(ctype)
    torch.manual_seed(42)
    for i in range(2):
        x = torch.randn(shape, dtype=torch.half, device=dev)
        for rank in (2, 8):
            with torch.random.fork_rng(devices=[dev]):
                torch.manual_seed(42 + i)
                sim = sim_compress(x, t, rank=rank)
            with torch.random.fork_rng(devices=[dev]):
                torch.manual_seed(42 + i)
                payload = slowpath_compress(x, t, rank=rank)
            n, c = shape
            expect = rank * (n + c) if ctype == "low-rank" else rank * (n + c) // 4 + 4 * rank
            assert payload.numel() == expect and payload.dtype == torch.half
            rec = slowpath_decompress(payload, x.shape, t, rank=rank)
            assert rel_l2(rec, sim) < tol


@pytest.mark.parametrize("n,c,rank", [(1024, 2048, 8), (290, 1536, 4), (2304, 3072, 32), (64, 128, 1), (130, 264, 20), (512, 3072, 64),
                                      (256, 1024, 16), (190, 648, 48), (4388, 3072, 32)])
@pytest.mark.parametrize("with_base", [False, True])
def test_lowrank_q_fused_decode_matches_two_step(n, c, rank, with_base):
    """cf_lowrank_q_reconstruct (decode of both int4 factors inside the reconstruct) against the sequence of
    slowpath.py:156-164: dequantize_int4 of U and of V^T (checked against the oracle's dequant), transpose,
    product, residual add.  Same MMA order as the unfused kernel, so the reconstructions agree bit for bit."""
    dev = _cuda()
    from compactfusion_b200.compress_lowrank import lowrank_q_reconstruct, lowrank_reconstruct
    from compactfusion_b200.compress_quantize import dequantize_int4
    from oracle import codecs as oc
    g = torch.Generator().manual_seed(n + c + rank)
    qu = torch.randint(0, 256, (n // 2, rank), dtype=torch.uint8, generator=g)
    qv = torch.randint(0, 256, (c // 2, rank), dtype=torch.uint8, generator=g)
    su, sv = [(torch.rand(rank, generator=g) * 0.02 + 1e-3).half() for _ in range(2)]
    mu, mv = [(-torch.rand(rank, generator=g) * 0.1).half() for _ in range(2)]
    flat = lambda t: t.contiguous().view(-1).view(torch.half)  # noqa: E731
    payload = torch.cat([flat(qu), su, mu, flat(qv), sv, mv]).to(dev)
    base = torch.randn(n, c, generator=g).half().to(dev) if with_base else None
    u = dequantize_int4(qu.to(dev), su.view(1, rank).to(dev), mu.view(1, rank).to(dev))
    vt = dequantize_int4(qv.to(dev), sv.view(1, rank).to(dev), mv.view(1, rank).to(dev))
    assert_bits_equal(u.cpu(), oc.int4_dequantize(qu.numpy(), su.view(1, rank), mu.view(1, rank)), "U decode vs oracle")
    assert_bits_equal(vt.cpu(), oc.int4_dequantize(qv.numpy(), sv.view(1, rank), mv.view(1, rank)), "V^T decode vs oracle")
    want = lowrank_reconstruct(u, vt.t().contiguous(), base)
    got = lowrank_q_reconstruct(payload, n, c, rank, base=base)
    assert_bits_equal(got, want, "fused LOW_RANK_Q decode")
    if with_base:   # in place, as the engines call it
        buf = base.clone()
        lowrank_q_reconstruct(payload, n, c, rank, base=buf, out=buf)
        assert_bits_equal(buf, want, "in place")


@pytest.mark.parametrize("n,c,rank", [(1024, 2048, 8), (290, 1536, 4), (2304, 3072, 32), (64, 128, 2), (130, 264, 20), (512, 3072, 64)])
def test_lowrank_q_pack_matches_two_quantize_calls(n, c, rank):
    """cf_lowrank_q_pack against slowpath.py:62-75 spelled out: quantize_int4(U), quantize_int4(V^T) (each checked
    against the oracle), concatenated.  Bit-exact, including a constant column (zero scale)."""
    dev = _cuda()
    from compactfusion_b200.compress_lowrank import lowrank_q_pack
    from compactfusion_b200.compress_quantize import quantize_int4
    from oracle import codecs as oc
    g = torch.Generator().manual_seed(3 * n + c + rank)
    u = (torch.randn(n, rank, generator=g) * 0.05).half()
    v = (torch.randn(rank, c, generator=g) * 3).half()
    u[:, 0] = 0.25   # constant column: scale 0, codes NaN -> 0
    parts = []
    for t in (u, v.t().contiguous()):
        q, s, m = quantize_int4(t.to(dev))
        oq, os_, om = oc.int4_quantize(t)
        assert_bits_equal(s.cpu(), os_, "scale vs oracle")
        assert_bits_equal(m.cpu(), om, "min vs oracle")
        assert (q.cpu().numpy() == oq).all(), "codes vs oracle"
        parts += [q.contiguous().view(-1).view(torch.half), s.view(-1), m.view(-1)]
    want = torch.cat(parts)
    got = lowrank_q_pack(u.to(dev), v.to(dev))
    assert got.numel() == (n * rank + c * rank) // 4 + 4 * rank
    assert_bits_equal(got, want, "LOW_RANK_Q payload")


@pytest.mark.parametrize("name,n,c,r", LRQ_WIRE_CASES)
def test_lowrank_q_wire_codec_vs_reference_goldens(golden_lowrank_q_wire, name, n, c, r):
    """Fixed factors U, V through the reference's own quantize_int4 + concatenation (slowpath.py:69-75) and its
    slowpath_decompress (slowpath.py:156-164), recorded by oracle/make_goldens.py: cf_lowrank_q_pack reproduces the
    reference's payload bit for bit; cf_lowrank_q_reconstruct decodes the reference's payload to the reference's
    reconstruction up to the rounding of the fp16 product (summation order is the only freedom)."""
    dev = _cuda()
    from compactfusion_b200.compress_lowrank import lowrank_q_pack, lowrank_q_reconstruct
    g = golden_lowrank_q_wire
    u, v = h16(g[f"{name}/u"]), h16(g[f"{name}/v"])
    want, ref = h16(g[f"{name}/payload"]), h16(g[f"{name}/recon"])
    got = lowrank_q_pack(u.to(dev), v.to(dev)).cpu()
    assert_bits_equal(got, want, "LOW_RANK_Q payload vs the reference's")
    rec = lowrank_q_reconstruct(want.to(dev), n, c, r).cpu()
    assert rel_l2(rec, ref) < 1e-3, rel_l2(rec, ref)
    assert float((rec.float() - ref.float()).abs().max()) <= 2e-2


def test_lowrank_state_machine_ef_invariant():
    """residual 1 + EF with the real LOW_RANK wire codec (examples/configs.py:63-85)."""
    dev = _cuda()
    import compactfusion_b200 as cf
    T = cf.COMPACT_COMPRESS_TYPE
    n, c, steps = 544, 3072, 4
    shape = (1, n, 24, 128)
    g = torch.Generator().manual_seed(1)
    x0 = torch.randn(n, c, generator=g)
    xs = [(x0 + 0.1 * t * torch.randn(n, c, generator=g)).half().view(shape).to(dev) for t in range(steps)]
    cfg = cf.CompactConfig(enabled=True, compress_func=lambda l, s: T.LOW_RANK if s >= 1 else T.WARMUP, comp_rank=8,
                           residual=1, ef=True)
    cf.compact_init(cfg)
    for t, x in enumerate(xs):
        ct = cfg.compress_func(0, t)
        comp = cf.compact_compress("0-0-k", x, ct, update_cache=True)
        rec = cf.compact_decompress("1-0-k", comp, ct, shape, update_cache=True)
        if t >= 1:
            assert comp.numel() == 8 * (n + c)
        assert torch.equal(cf.compact_cache().get_base("0-0-k"), cf.compact_cache().get_base("1-0-k"))
        assert rel_l2(rec, x) < 0.5  # iid step noise is not low-rank: rank 8 only tracks the slow part
