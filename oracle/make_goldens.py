"""Generate tests/golden/*.npz by running the REFERENCE itself (imported from
/root/reference, eager, CPU) on seeded inputs.  TEST INFRASTRUCTURE ONLY.

Run in the build container (the only place /root/reference exists):

    TORCHDYNAMO_DISABLE=1 python oracle/make_goldens.py

The fixtures are committed; tests/test_oracle_golden.py pins the oracle
restatement (oracle/codecs.py, oracle/state.py) against them on any machine,
and the -m gpu tests pin the CUDA kernels against the same files.
fp16 tensors are stored as their uint16 bit patterns (exact).
"""
import os
import sys

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle.ref_loader import load_reference, REFERENCE_ROOT  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def bits(t: torch.Tensor) -> np.ndarray:
    """fp16 tensor -> uint16 bit pattern array (exact storage)."""
    assert t.dtype == torch.half
    return t.contiguous().view(torch.int16).numpy().view(np.uint16).copy()


def make_inputs(seed, n, c, realistic=None):
    g = torch.Generator().manual_seed(seed)
    if realistic is not None:
        x = realistic[:n, :c].clone()
        base = (x.float() * 0.97 + 0.2 * torch.randn(n, c, generator=g)).half()
    else:
        x = torch.randn(n, c, generator=g).half()  # tests/compact/compress_fastpath_test.py:57-58
        base = (torch.randn(n, c, generator=g) * 0.1).half()
    return x.contiguous(), base.contiguous()


def codec_goldens(ref):
    from xfuser.compact import compress_quantize as cq
    from xfuser.compact import compress_topk as ct
    from xfuser.compact import compress_lowrank as cl

    act_path = os.path.join(REFERENCE_ROOT, "compact_plot", "activation_dump", "20-0-k_step10.pt")
    act = torch.load(act_path, map_location="cpu").half().reshape(-1, 3072)
    cases = [
        ("rand_64x256", 42, 64, 256, None),
        ("rand_48x1152", 43, 48, 1152, None),
        ("rand_130x64", 44, 130, 64, None),
        ("flux_k_96x512", 45, 96, 512, act),
    ]
    out = {}
    for name, seed, n, c, real in cases:
        x, base = make_inputs(seed, n, c, real)
        d = x - base
        out[f"{name}/x"] = bits(x)
        out[f"{name}/base"] = bits(base)
        # BINARY: sim_binary (compress_quantize.py:300) is the reference's ground truth for the fastpath
        out[f"{name}/sim_binary"] = bits(cq.sim_binary(d, rank=-1))
        # INT2: eager quantize/dequantize + sim
        p, chan, tok = cq.quantize_int2(d)
        out[f"{name}/int2_packed"] = p.numpy()
        out[f"{name}/int2_chan"] = bits(chan)
        out[f"{name}/int2_tok"] = bits(tok)
        out[f"{name}/int2_deq"] = bits(cq.dequantize_int2(p, chan, tok))
        out[f"{name}/sim_int2"] = bits(cq.sim_int2(d))
        out[f"{name}/sim_int2_minmax"] = bits(cq.sim_int2_minmax(d))
        # INT4
        p4, s4, m4 = cq.quantize_int4(d)
        out[f"{name}/int4_packed"] = p4.numpy()
        out[f"{name}/int4_scale"] = bits(s4)
        out[f"{name}/int4_min"] = bits(m4)
        out[f"{name}/int4_deq"] = bits(cq.dequantize_int4(p4, s4, m4))
        out[f"{name}/sim_int4_d0"] = bits(cq.sim_int4(d, dim=0))
        out[f"{name}/sim_int4_d1"] = bits(cq.sim_int4(d, dim=1))
        # INT8
        q8, s8, z8 = cq.quantize_int8(d)
        out[f"{name}/int8_q"] = q8.numpy()
        out[f"{name}/int8_scale"] = bits(s8)
        out[f"{name}/int8_zp"] = z8.numpy()
        out[f"{name}/int8_deq"] = bits(cq.dequantize_int8(q8, s8, z8))
    # top-k: sim_topk on tie-free data (torch.topk tie order is unspecified)
    g = torch.Generator().manual_seed(7)
    xt = torch.randn(4, 1024, generator=g).half()
    out["topk/x"] = bits(xt)
    for m in (2, 4, 8, 16):
        out[f"topk/sim_m{m}"] = bits(ct.sim_topk(xt.clone(), m))
    # subspace iteration with an explicit init_q and with the seeded global RNG
    g = torch.Generator().manual_seed(11)
    a = (torch.randn(96, 6, generator=g) @ torch.randn(6, 256, generator=g) + 0.05 * torch.randn(96, 256, generator=g)).half()
    q0 = torch.randn(256, 4, generator=g)
    u, v, q = cl.subspace_iter(a, 4, 2, init_q=q0)
    out["lowrank/a"] = bits(a)
    out["lowrank/q0"] = q0.numpy()
    out["lowrank/u"], out["lowrank/v"], out["lowrank/q"] = bits(u), bits(v), bits(q)
    torch.manual_seed(123)
    u2, v2, _ = cl.subspace_iter(a, 8, 2)
    out["lowrank/seed123_r8_uv"] = bits((u2.float() @ v2.float()).half())
    np.savez_compressed(os.path.join(OUT, "codecs.npz"), **out)
    print("codecs.npz:", len(out), "arrays")


def slowpath_goldens(ref):
    from xfuser.compact import slowpath as sp
    from xfuser.compact.utils import COMPACT_COMPRESS_TYPE as T

    out = {}
    g = torch.Generator().manual_seed(21)
    x = (torch.randn(64, 5, generator=g) @ torch.randn(5, 256, generator=g) + 0.1 * torch.randn(64, 256, generator=g)).half()
    out["x"] = bits(x)
    for name, t, kw in (
        ("low_rank_r8", T.LOW_RANK, dict(rank=8)),
        ("low_rank_q_r4", T.LOW_RANK_Q, dict(rank=4)),
    ):
        torch.manual_seed(123)
        p = sp.slowpath_compress(x, t, **kw)
        out[f"{name}/payload"] = bits(p)
        out[f"{name}/recon"] = bits(sp.slowpath_decompress(p, x.shape, t, **kw))
        torch.manual_seed(123)
        out[f"{name}/sim"] = bits(sp.sim_compress(x, t, **kw))
    np.savez_compressed(os.path.join(OUT, "slowpath.npz"), **out)
    print("slowpath.npz:", len(out), "arrays")


def state_machine_goldens(ref):
    """Drive the reference's own compact_compress / compact_decompress
    (xfuser/compact/main.py:169,322) for several CompactConfig flavours.
    Sender key "0-0-k", receiver key "1-0-k" (two caches in one process)."""
    from xfuser.compact import main as cm
    from xfuser.compact.utils import CompactConfig, COMPACT_COMPRESS_TYPE as T

    n, c, steps = 32, 128, 6
    shape4 = (1, n, 4, c // 4)
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(n, c, generator=g)
    xs = [(x0 + 0.05 * t * torch.randn(n, c, generator=g)).half().view(shape4).contiguous() for t in range(steps)]
    out = {"xs": np.stack([bits(x) for x in xs])}
    flavours = {
        "sim_int4_r1_ef": (dict(residual=1, ef=True, simulate=True, comp_rank=-1), T.INT4),
        "sim_binary_r1_ef": (dict(residual=1, ef=True, simulate=True, comp_rank=-1), T.BINARY),
        "sim_int2_r2_ef": (dict(residual=2, ef=True, simulate=True, comp_rank=-1, delta_decay_factor=0.5), T.INT2),
        "sim_int4_r1_noef": (dict(residual=1, ef=False, simulate=True, comp_rank=-1), T.INT4),
        "sim_int4_r0": (dict(residual=0, ef=False, simulate=True, comp_rank=-1), T.INT4),
        "real_lowrank4_r1_ef": (dict(residual=1, ef=True, simulate=False, comp_rank=4), T.LOW_RANK),
    }
    for name, (kw, ctype) in flavours.items():
        warm = 2 if kw["residual"] == 2 else 1
        cfg = CompactConfig(enabled=True, compress_func=lambda l, s, w=warm, ct=ctype: ct if s >= w else T.WARMUP, **kw)
        cm.compact_init(cfg)
        for t in range(steps):
            cm.compact_set_step(t)
            ct = cfg.compress_func(0, t)
            torch.manual_seed(1000 + t)  # low-rank draws from the global RNG
            comp = cm.compact_compress("0-0-k", xs[t], ct, update_cache=True)
            rec = cm.compact_decompress("1-0-k", comp, ct, shape4, update_cache=True)
            out[f"{name}/comp{t}"] = bits(comp.reshape(-1))
            out[f"{name}/recon{t}"] = bits(rec.reshape(-1))
            for who, key in (("send", "0-0-k"), ("recv", "1-0-k")):
                b = cm.compact_cache().get_base(key)
                if b is not None:  # residual 0 keeps no cache
                    out[f"{name}/{who}_base{t}"] = bits(b.reshape(-1))
                db = cm.compact_cache().get_delta_base(key)
                if db is not None:
                    out[f"{name}/{who}_dbase{t}"] = bits(db.reshape(-1))
    np.savez_compressed(os.path.join(OUT, "state_machine.npz"), **out)
    print("state_machine.npz:", len(out), "arrays")


def lowrank_q_wire_goldens(ref):
    """The LOW_RANK_Q wire codec on FIXED factors: the payload the reference assembles from U (N, r) and V (r, C)
    (its own quantize_int4 on U and on V^T, concatenated as in slowpath.py:69-75) and what its slowpath_decompress
    makes of that payload (slowpath.py:156-164).  Pins cf_lowrank_q_pack / cf_lowrank_q_reconstruct (and the oracle's
    int4) against the reference itself, without the random start of subspace_iter in the way."""
    from xfuser.compact import compress_quantize as cq
    from xfuser.compact import slowpath as sp
    from xfuser.compact.utils import COMPACT_COMPRESS_TYPE as T

    out = {}
    for name, seed, n, c, r in (("n64_c256_r4", 7, 64, 256, 4), ("n130_c264_r20", 8, 130, 264, 20), ("n96_c512_r32", 9, 96, 512, 32)):
        g = torch.Generator().manual_seed(seed)
        u = (torch.randn(n, r, generator=g) * 0.05).half()
        v = (torch.randn(r, c, generator=g) * 3).half()
        qu, su, mu = cq.quantize_int4(u)
        qv, sv, mv = cq.quantize_int4(v.t())   # (slowpath.py:70: the transposed view)
        parts = [qu.view(torch.half).reshape(-1), su.reshape(-1), mu.reshape(-1),
                 qv.view(torch.half).reshape(-1), sv.reshape(-1), mv.reshape(-1)]
        payload = torch.cat([p_.contiguous() for p_ in parts])
        out[f"{name}/u"], out[f"{name}/v"] = bits(u), bits(v)
        out[f"{name}/payload"] = bits(payload)
        out[f"{name}/recon"] = bits(sp.slowpath_decompress(payload, (n, c), T.LOW_RANK_Q, rank=r))
    np.savez_compressed(os.path.join(OUT, "lowrank_q_wire.npz"), **out)
    print("lowrank_q_wire.npz:", len(out), "arrays")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # reduction order independent of the host's core count
    ref = load_reference()
    only = sys.argv[1:]   # e.g. `python oracle/make_goldens.py lowrank_q_wire`: just that file
    for name, fn in (("codecs", codec_goldens), ("slowpath", slowpath_goldens), ("state_machine", state_machine_goldens),
                     ("lowrank_q_wire", lowrank_q_wire_goldens)):
        if not only or name in only:
            fn(ref)
