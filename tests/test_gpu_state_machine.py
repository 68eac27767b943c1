"""compact_compress / compact_decompress on the GPU vs the reference-generated goldens and
the CPU oracle state machine (main.py:169-388)."""
import numpy as np
import pytest
import torch

from conftest import assert_bits_equal, h16, rel_l2
from oracle.state import OracleCompact

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _types():
    from compactfusion_b200.utils import COMPACT_COMPRESS_TYPE as T
    return T


def _run(cfg_kw, ctype_name, xs, shape, warm, seed_fn=None):
    """Sender ("0-0-k") and receiver ("1-0-k") in one process, like oracle/make_goldens.py."""
    import compactfusion_b200 as cf
    T = _types()
    ctype = T(ctype_name)
    cfg = cf.CompactConfig(enabled=True, compress_func=lambda l, s: ctype if s >= warm else T.WARMUP, **cfg_kw)
    cf.compact_init(cfg)
    outs = []
    for t, x in enumerate(xs):
        cf.compact_set_step(t)
        ct = cfg.compress_func(0, t)
        if seed_fn:
            seed_fn(t)
        comp = cf.compact_compress("0-0-k", x, ct, update_cache=True)
        rec = cf.compact_decompress("1-0-k", comp, ct, shape, update_cache=True)
        outs.append((comp.reshape(-1).clone(), rec.reshape(-1).clone(),
                     cf.compact_cache().get_base("0-0-k"), cf.compact_cache().get_base("1-0-k")))
    return outs


def test_config1_sim_int4_ef_matches_reference_golden(golden_state):
    """BASELINE configs[0] in miniature: INT4 simulate + residual 1 + EF.  min/max are exact,
    so every step must match the reference bit for bit."""
    dev = _cuda()
    g = golden_state
    n, c, steps = 32, 128, 6
    shape = (1, n, 4, c // 4)
    xs = [h16(g["xs"][t]).view(shape).to(dev) for t in range(steps)]
    for name, kw in (("sim_int4_r1_ef", dict(residual=1, ef=True, simulate=True, comp_rank=-1)),
                     ("sim_int4_r1_noef", dict(residual=1, ef=False, simulate=True, comp_rank=-1)),
                     ("sim_int4_r0", dict(residual=0, ef=False, simulate=True, comp_rank=-1))):
        outs = _run(kw, "int4", xs, shape, warm=1)
        for t, (comp, rec, sb, rb) in enumerate(outs):
            assert_bits_equal(comp, h16(g[f"{name}/comp{t}"]), f"{name} comp{t}")
            assert_bits_equal(rec, h16(g[f"{name}/recon{t}"]), f"{name} recon{t}")
            if kw["residual"]:
                assert_bits_equal(sb.reshape(-1), h16(g[f"{name}/send_base{t}"]), f"{name} send_base{t}")
                assert_bits_equal(rb.reshape(-1), h16(g[f"{name}/recv_base{t}"]), f"{name} recv_base{t}")


@pytest.mark.parametrize("name,ctype,kw,warm", [
    ("sim_binary_r1_ef", "binary", dict(residual=1, ef=True, simulate=True, comp_rank=-1), 1),
    ("sim_int2_r2_ef", "int2", dict(residual=2, ef=True, simulate=True, comp_rank=-1, delta_decay_factor=0.5), 2),
])
def test_sim_mean_scale_codecs_track_reference(golden_state, name, ctype, kw, warm):
    """Mean-scale codecs: 1-ulp scale differences feed back through EF, so compare with the
    reference's own tolerance (rel-L2 2e-2 on reconstructions) at every step."""
    dev = _cuda()
    g = golden_state
    n, c, steps = 32, 128, 6
    shape = (1, n, 4, c // 4)
    xs = [h16(g["xs"][t]).view(shape).to(dev) for t in range(steps)]
    outs = _run(kw, ctype, xs, shape, warm=warm)
    for t, (comp, rec, sb, rb) in enumerate(outs):
        assert rel_l2(rec, h16(g[f"{name}/recon{t}"])) < 2e-2, f"{name} recon{t}"
        assert torch.equal(sb, rb), f"{name}: caches differ at step {t}"
    assert torch.equal(outs[-1][2], outs[-1][3]), "EF invariant: sender and receiver caches identical"


@pytest.mark.parametrize("codec", ["binary", "int2"])
def test_fastpath_state_machine_vs_oracle(codec):
    """Fastpath (the production configs, examples/configs.py:39-61): wire payload layout,
    sender/receiver cache identity over 8 steps, FLUX W=8 shard shape."""
    dev = _cuda()
    import compactfusion_b200 as cf
    n, c, steps = 576, 3072, 8
    shape = (1, n, 24, 128)
    g = torch.Generator().manual_seed(3)
    x0 = torch.randn(n, c, generator=g)
    xs = [(0.97 ** t * x0 + 0.2 * torch.randn(n, c, generator=g)).half().view(shape) for t in range(steps)]
    kw = dict(residual=1, ef=True, simulate=False, fastpath=True, comp_rank=-1)
    outs = _run(kw, codec, [x.to(dev) for x in xs], shape, warm=1)
    snd, rcv = OracleCompact(**kw), OracleCompact(**kw)
    per_byte = 8 if codec == "binary" else 4
    for t, (comp, rec, sb, rb) in enumerate(outs):
        ct = codec if t >= 1 else "warmup"
        o_comp = snd.compress("k", xs[t], ct, update_cache=True)
        o_rec = rcv.decompress("k", o_comp, ct, shape, update_cache=True)
        assert comp.numel() == o_comp.numel()
        assert torch.equal(sb, rb), f"step {t}: sender and receiver caches differ"
        assert torch.equal(rec.view(n, c), rb), "decompress returns the cached tensor's value"
        if t == 1:  # first compressed step: bases are identical on both sides, compare closely
            nb = n * c // per_byte
            mism = float((comp.view(torch.uint8)[:nb].cpu() != o_comp.view(torch.uint8)[:nb]).float().mean())
            assert mism <= (0 if codec == "binary" else 1e-3)
            assert rel_l2(rec, o_rec.reshape(-1)) < (1e-3 if codec == "binary" else 2e-2)
        assert rel_l2(rec, o_rec.reshape(-1)) < 5e-2
        # compression error stays bounded under error feedback
        assert rel_l2(rec, xs[t].reshape(-1)) < 0.5


def test_inplace_cache_mode_is_equivalent():
    dev = _cuda()
    import compactfusion_b200 as cf
    n, c = 256, 1024
    shape = (1, n, 8, 128)
    g = torch.Generator().manual_seed(9)
    xs = [torch.randn(n, c, generator=g).half().view(shape).to(dev) for _ in range(5)]
    keep = [x.clone() for x in xs]
    kw = dict(residual=1, ef=True, simulate=False, fastpath=True, comp_rank=-1)
    ref = _run(kw, "binary", xs, shape, warm=1)
    cf.compact_set_inplace(True)
    try:
        got = _run(kw, "binary", xs, shape, warm=1)
    finally:
        cf.compact_set_inplace(False)
    for (c0, r0, _, _), (c1, r1, _, _) in zip(ref, got):
        # payloads hold code bytes viewed as fp16 (NaN patterns): compare bit patterns
        assert torch.equal(c0.view(torch.int16), c1.view(torch.int16)) and torch.equal(r0, r1)
    for x, k in zip(xs, keep):  # no caller tensor (incl. the warm-up one cached as base) was overwritten
        assert torch.equal(x, k)
