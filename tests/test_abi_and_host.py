"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the header
declares, shapes / payload sizes / config rules match the reference, the product fails loudly
without CUDA, and the host-side exchange logic (all-gather, ring) runs under gloo with
world_size 2 (codec calls replaced by the oracle IN THE TEST ONLY)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import free_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from compactfusion_b200 import build as cf_build
    return cf_build.build()


def test_library_exports_every_declared_symbol(built_lib):
    from compactfusion_b200 import _native as nv
    hdr = open(os.path.join(ROOT, "include", "compactb200.h")).read()
    declared = set(re.findall(r"CF_API\s+[\w\s\*]+?\b(cf_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no CF_API declarations parsed"
    assert declared == set(nv.SYMBOLS), (declared ^ set(nv.SYMBOLS))
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (cf_[a-z0-9_]+)", out))
    assert declared <= exported, declared - exported
    lib = nv.lib()  # binds every symbol with its prototype
    assert lib.cf_abi_version() == 1


def test_ctypes_prototypes_match_the_header():
    """Every prototype in include/compactb200.h against the ctypes binding: same arity, and per parameter the
    same class (int / int64_t / size_t / pointer / pointer-to-pointer / float pointer); same return type."""
    import ctypes
    from compactfusion_b200 import _native as nv
    hdr = open(os.path.join(ROOT, "include", "compactb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    protos = re.findall(r"CF_API\s+([\w\s\*]+?)\b(cf_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S)
    assert {name for _, name, _ in protos} == set(nv.SYMBOLS)

    def classify_c(decl):
        decl = " ".join(decl.replace("const", " ").split())
        if decl in ("void", ""):
            return None
        stars = decl.count("*")
        base = decl.replace("*", " ").split()
        base = " ".join(base[:-1]) if len(base) > 1 else base[0]  # drop the parameter name
        if stars >= 2:
            return "pp"
        if stars == 1:
            return "p"
        return {"int": "int", "int64_t": "i64", "size_t": "size", "cf_stream_t": "p", "unsigned": "int"}[base]

    def classify_py(t):
        if t is ctypes.c_int:
            return "int"
        if t is ctypes.c_int64:
            return "i64"
        if t is ctypes.c_size_t:
            return "size"
        if t in (ctypes.c_void_p, ctypes.c_char_p):
            return "p"
        if t is nv._VPP:
            return "pp"
        if isinstance(t, type) and issubclass(t, ctypes._Pointer):
            return "p" if t._type_ is not ctypes.c_void_p else "pp"
        raise AssertionError(f"unclassified ctypes type {t}")

    for ret, name, params in protos:
        res, args = nv.SYMBOLS[name]
        c_args = [classify_c(p) for p in params.split(",")]
        c_args = [a for a in c_args if a is not None]
        py_args = [classify_py(a) for a in args]
        # a void** OUT parameter (cf_ipc_alloc / cf_ipc_open) is bound as POINTER(c_void_p): both "pp"
        assert c_args == py_args, f"{name}: header {c_args} vs ctypes {py_args}"
        c_ret = classify_c(ret.strip() + " _") or "void"
        assert c_ret == classify_py(res), f"{name}: return {c_ret} vs {classify_py(res)}"


def test_workspace_sizes_without_gpu(built_lib):
    from compactfusion_b200 import _native as nv
    for codec in (nv.CODEC_BINARY, nv.CODEC_INT2, nv.CODEC_INT4, nv.CODEC_INT8):
        b = nv.workspace_bytes(codec, 4096, 3072)
        assert 0 < b < 64 << 20
    assert nv.workspace_bytes(nv.CODEC_BINARY, 4096, 3072, 0, 4) == 4 * nv.workspace_bytes(nv.CODEC_BINARY, 4096, 3072)
    assert nv.workspace_bytes(nv.CODEC_LOWRANK, 4096, 3072, 32) > 4096 * 32 * 4


def test_product_fails_loudly_on_cpu_tensors(built_lib):
    import compactfusion_b200 as cf
    from compactfusion_b200 import _native as nv
    from compactfusion_b200.compress_quantize import quantize_int4, sim_binary
    x = torch.randn(8, 64).half()
    for fn in (lambda: sim_binary(x, rank=-1), lambda: quantize_int4(x)):
        with pytest.raises(nv.NativeError):
            fn()
    T = cf.COMPACT_COMPRESS_TYPE
    cf.compact_init(cf.CompactConfig(enabled=True, compress_func=lambda l, s: T.BINARY, comp_rank=-1, residual=1,
                                     ef=True, fastpath=True))
    cf.compact_compress("0-0-k", x, T.WARMUP, update_cache=True)
    with pytest.raises(nv.NativeError):
        cf.compact_compress("0-0-k", x, T.BINARY, update_cache=True)


def test_missing_library_raises(monkeypatch, built_lib):
    from compactfusion_b200 import _native as nv
    monkeypatch.setattr(nv, "_lib", None)
    monkeypatch.setattr(nv, "LIB_PATH", "/nonexistent/libcompactb200.so")
    with pytest.raises(nv.NativeError):
        nv.lib()


def test_payload_sizes_match_reference_wire_formats():
    """SURVEY.md App-A / BASELINE.md section 2 byte counts for (4096, 3072)."""
    from compactfusion_b200.main import fastpath_payload_numel
    from compactfusion_b200.utils import COMPACT_COMPRESS_TYPE as T
    assert fastpath_payload_numel(4096, 3072, T.BINARY) * 2 == 1_587_200
    assert fastpath_payload_numel(4096, 3072, T.INT2) * 2 == 3_160_064
    from oracle import codecs as oc
    x = torch.randn(64, 256).half()
    p, u, v, _ = oc.binary_quant(x, torch.zeros_like(x), False)
    assert oc.fastpath_payload(p, u, v).numel() == fastpath_payload_numel(64, 256, T.BINARY)


def test_config_rules_match_reference():
    import compactfusion_b200 as cf
    T = cf.COMPACT_COMPRESS_TYPE
    assert T("low-rank-int4") is T.LOW_RANK_Q and T.BINARY.value == "binary"
    with pytest.raises(AssertionError):
        cf.CompactConfig(enabled=True, residual=0, ef=True)
    with pytest.raises(AssertionError):
        cf.CompactConfig(enabled=True, residual=2, ef=False)
    with pytest.raises(AssertionError):
        cf.CompactConfig(enabled=True, residual=1, ef=True, simulate=True, fastpath=True)
    with pytest.raises(AssertionError):
        cf.CompactConfig(enabled=True, residual=1, ef=True, patch_gather_fwd_config=cf.PatchConfig(True, False, 1))
    with pytest.raises(AssertionError):
        cf.PatchConfig(use_compact=True, async_comm=True, async_warmup=1)
    cfg = cf.CompactConfig(enabled=True, override_with_patch_gather_fwd=True,
                           patch_gather_fwd_config=cf.PatchConfig(True, False, 1),
                           compress_func=lambda l, s: T.INT2 if s >= 1 else T.WARMUP, comp_rank=-1, residual=1, ef=True,
                           fastpath=True)
    assert cfg.get_compress_type() == "INT2"
    assert cf.CompactConfig().get_compress_type() == "NO_COMPACT"


def test_warmup_and_shape_rules_on_cpu():
    """WARMUP needs no kernel: passthrough + cache of the caller's tensor (main.py:195-209)."""
    import compactfusion_b200 as cf
    from compactfusion_b200.main import _to_2d_shape
    T = cf.COMPACT_COMPRESS_TYPE
    assert _to_2d_shape((2, 5, 4, 8)) == (10, 32) and _to_2d_shape((2, 5, 32)) == (10, 32)
    cf.compact_init(cf.CompactConfig(enabled=True, compress_func=lambda l, s: T.WARMUP, comp_rank=-1, residual=1,
                                     ef=True, fastpath=True))
    cf.compact_set_step(0)
    assert cf.compact_get_step() == 0
    x = torch.randn(1, 6, 2, 8).half()
    out = cf.compact_compress("3-0-k", x, T.WARMUP, update_cache=True)
    assert out.data_ptr() == x.data_ptr() and out.shape == x.shape
    assert cf.compact_cache().get_base("3-0-k").shape == (6, 16)
    rec = cf.compact_decompress("3-1-k", out, T.WARMUP, x.shape, update_cache=True)
    assert torch.equal(rec, x)
    cf.compact_reset()
    assert cf.compact_cache().get_base("3-0-k") is None and cf.compact_get_step() is None


WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch
import torch.distributed as dist
import compactfusion_b200 as cf
from compactfusion_b200 import main as cm, ring as cr, compress_quantize as cq, fastpath as fp
from compactfusion_b200.utils import COMPACT_COMPRESS_TYPE as T
from oracle import codecs as oc
from oracle.state import OracleCompact, all_gather_step, ring_step

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")

# ---- TEST-ONLY stand-ins for the CUDA codec calls (the product has no CPU path) ----
def sign_compress(codec, x, base, update_cache, packed=None, u=None, v=None, new_base=None):
    fn = oc.binary_quant if codec == 1 else oc.int2_quant
    p, uu, vv, nb = fn(x, base if base is not None else torch.zeros_like(x), update_cache)
    packed.copy_(torch.from_numpy(p)); u.copy_(uu); v.copy_(vv)
    if update_cache:
        new_base.copy_(nb)
    return packed, u, v, (new_base if update_cache else None)
def peers(tags, payloads, ctype, shape2d):
    n, c = shape2d
    outs = []
    for t, p in zip(tags, payloads):
        pk, u, v = cm._payload_views(p, n, c, ctype)
        fn = oc.binary_dequant if ctype == T.BINARY else oc.int2_dequant
        o = fn(pk.numpy(), u, v, cm._cache.get_base(t))
        cm._put(t, o)
        outs.append(o)
    return outs
cm._sign_compress = sign_compress
cm._decompress_peers_batched = peers

n, c, steps = 16, 64, 4
shape = (1, n, 2, 32)
g = torch.Generator().manual_seed(7)
x0 = [torch.randn(n, c, generator=g) for _ in range(world)]
xs = [[(x0[r] + 0.1 * t * torch.randn(n, c, generator=g)).half().view(shape) for r in range(world)] for t in range(steps)]

for codec in (T.BINARY, T.INT2):
    # ---------------- compressed all-gather (patch parallel) ----------------
    cfg = cf.CompactConfig(enabled=True, override_with_patch_gather_fwd=True,
                           patch_gather_fwd_config=cf.PatchConfig(True, False, 1),
                           compress_func=lambda l, s: codec if s >= 1 else T.WARMUP, comp_rank=-1, residual=1,
                           ef=True, fastpath=True)
    cf.compact_init(cfg)
    oracle_ranks = [OracleCompact(residual=1, ef=True, fastpath=True) for _ in range(world)]
    for t in range(steps):
        ct = cfg.compress_func(0, t)
        got = cf.compact_all_gather("5-k", xs[t][rank], ct)
        want, _ = all_gather_step(oracle_ranks, "5-k", xs[t], ct.value)
        assert len(got) == world
        for i in range(world):
            assert torch.equal(got[i], want[rank][i]), (codec, t, i)
    # ---------------- compressed ring ----------------
    cfg = cf.CompactConfig(enabled=True, compress_func=lambda l, s: codec if s >= 1 else T.WARMUP, comp_rank=-1,
                           residual=1, ef=True, fastpath=True, check_consist=True)
    cf.compact_init(cfg)
    seen = []
    real_attn = cr.attn_forward
    def spy(q, k, v, *a, **kw):
        seen.append((k.clone(), v.clone()))
        return real_attn(q, k, v, *a, **kw)
    cr.attn_forward = spy
    ok_ranks = [OracleCompact(residual=1, ef=True, fastpath=True) for _ in range(world)]
    ov_ranks = [OracleCompact(residual=1, ef=True, fastpath=True) for _ in range(world)]
    for t in range(steps):
        seen.clear()
        q = xs[t][rank].float()
        out, lse, _ = cf.compact_fwd(q.half(), xs[t][rank], xs[t][rank].flip(1).contiguous(), causal=False,
                                     mod_idx=2, current_iter=t)
        ks = ring_step(ok_ranks, 2, xs[t], cfg.compress_func(2, t).value, "k")
        vs = ring_step(ov_ranks, 2, [x.flip(1).contiguous() for x in xs[t]], cfg.compress_func(2, t).value, "v")
        assert len(seen) == world
        for s in range(world):
            assert torch.equal(seen[s][0], ks[rank][s]) and torch.equal(seen[s][1], vs[rank][s]), (codec, t, s)
        # blockwise LSE merge == attention over the concatenated sequence
        kcat = torch.cat([b for b in ks[rank]], dim=1)
        vcat = torch.cat([b for b in vs[rank]], dim=1)
        ref, ref_lse = real_attn(q.half(), kcat, vcat, 0.0, None, False)
        assert torch.allclose(out.float(), ref.float(), atol=2e-3), float((out.float() - ref.float()).abs().max())
        assert torch.allclose(lse, ref_lse, atol=1e-3)
    cr.attn_forward = real_attn
    assert cf.compact_cache().passed_count == steps

# ---------------- patch_gather_fwd, all three modes (patchpara/fwd.py:20-237) ----------------
from compactfusion_b200.patchpara import fwd as pf
seen = []
real_attn = pf.attn_forward
def spy2(q, k, v, *a, **kw):
    seen.append((k.clone(), v.clone()))
    return real_attn(q, k, v, *a, **kw)
pf.attn_forward = spy2
vs_ = [[x.flip(1).contiguous() for x in step] for step in xs]
for mode in ("compact", "sync", "async"):
    cfg = cf.CompactConfig(enabled=True, override_with_patch_gather_fwd=True,
                           patch_gather_fwd_config=cf.PatchConfig(mode == "compact", mode == "async", 1),
                           compress_func=lambda l, s: T.BINARY if s >= 1 else T.WARMUP, comp_rank=-1, residual=1,
                           ef=True, fastpath=True)
    cf.compact_init(cfg)
    ok_ranks = [OracleCompact(residual=1, ef=True, fastpath=True) for _ in range(world)]
    ov_ranks = [OracleCompact(residual=1, ef=True, fastpath=True) for _ in range(world)]
    for t in range(steps):
        seen.clear()
        out, lse, _ = cf.compact_fwd(xs[t][rank], xs[t][rank], vs_[t][rank], causal=False, mod_idx=3, current_iter=t)
        assert len(seen) == 1 and out.shape == xs[t][rank].shape
        if mode == "compact":     # every origin's reconstruction, own shard included (main.py:410-419)
            ct = cfg.compress_func(3, t).value
            wk, _ = all_gather_step(ok_ranks, "3-k", xs[t], ct)
            wv, _ = all_gather_step(ov_ranks, "3-v", vs_[t], ct)
            want_k, want_v = torch.cat(wk[rank], dim=1), torch.cat(wv[rank], dim=1)
        else:                      # raw shards; stale-async: the peers' PREVIOUS step, the fresh local shard
            stale = mode == "async" and t >= 1
            want_k = torch.cat([xs[t - 1 if (stale and r != rank) else t][r] for r in range(world)], dim=1)
            want_v = torch.cat([vs_[t - 1 if (stale and r != rank) else t][r] for r in range(world)], dim=1)
        assert torch.equal(seen[0][0], want_k) and torch.equal(seen[0][1], want_v), (mode, t)
pf.attn_forward = real_attn
dist.destroy_process_group()
print("WORKER_OK", rank)
'''


def test_exchange_host_logic_gloo_world2(tmp_path, built_lib):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=free_port(), OMP_NUM_THREADS="2")
    procs = []
    for r in range(2):
        e = dict(env, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"WORKER_OK {r}" in o, o[-3000:]


def test_shim_installs_under_reference_module_path():
    """Every name xDiT and the reference's tests import from xfuser.compact.* 
    resolves to this package after shim.install()."""
    import importlib
    import compactfusion_b200.shim as shim
    shim.install()
    try:
        wanted = {
            "xfuser.compact.main": ["compact_config", "compact_get_step", "compact_set_step", "CompactConfig",
                                    "compact_init", "compact_reset", "compact_hello", "compact_cache",
                                    "compact_compress", "compact_decompress", "compact_all_gather", "allgather_cache"],
            "xfuser.compact.ring": ["compact_fwd"],
            "xfuser.compact.utils": ["CompactConfig", "CompactCache", "COMPACT_COMPRESS_TYPE", "ALLOW_DEPRECATED"],
            "xfuser.compact.patchpara.df_utils": ["PatchConfig"],
            "xfuser.compact.stats": ["stats_verbose", "stats_verbose_steps", "plot_eigenvalues", "save_eigenvalues",
                                     "dump_err_vs_steps"],
            "xfuser.compact.fastpath": ["binary_quant_fastpath", "binary_dequant_fastpath", "int2_quant_fastpath",
                                        "int2_dequant_fastpath"],
            "xfuser.compact.compress_quantize": ["quantize_1bit", "dequantize_1bit", "quantize_int2", "dequantize_int2",
                                                 "sim_int2", "sim_binary", "quantize_int4", "dequantize_int4",
                                                 "sim_int4", "quantize_int8", "dequantize_int8"],
            "xfuser.compact.compress_topk": ["topk_compress", "topk_decompress", "sim_topk"],
            "xfuser.compact.compress_lowrank": ["subspace_iter", "svd"],
            "xfuser.compact.slowpath": ["slowpath_compress", "slowpath_decompress", "sim_compress"],
        }
        for mod, names in wanted.items():
            m = importlib.import_module(mod)
            assert m.__name__.startswith("compactfusion_b200"), mod
            for n in names:
                assert hasattr(m, n), f"{mod}.{n}"
    finally:
        shim.uninstall()
    assert "xfuser.compact.main" not in sys.modules


def test_stats_logger_host_logic(monkeypatch, tmp_path, capsys):
    """StatsLogger bookkeeping (record fields, deferred read-back, similarity from norms, summaries, dump
    formats of plot.py:413-560) with the device reduction replaced by torch-CPU IN THE TEST ONLY."""
    from compactfusion_b200 import stats as st

    def cpu_pair(self, a, b):
        if not self._tables or self._used == st._CHUNK_ROWS:
            self._tables.append(torch.zeros((st._CHUNK_ROWS, 4), dtype=torch.float32))
            self._used = 0
        d = a.double() - b.double()
        self._tables[-1][self._used] = torch.tensor([float((d * d).sum()), float((b.double() ** 2).sum()),
                                                     float(d.abs().max()), float(b.double().abs().max())])
        self._used += 1
        return (len(self._tables) - 1, self._used - 1)

    monkeypatch.setattr(st.StatsLogger, "_pair", cpu_pair)
    monkeypatch.setattr(st, "_CHUNK_ROWS", 5)  # several tables
    st.stats_clear()
    g = torch.Generator().manual_seed(3)
    xs = {k: [torch.randn(16, 32, generator=g).half()] for k in ("0-0-k", "0-0-v")}
    for k in xs:
        for _ in range(3):
            xs[k].append((0.9 * xs[k][-1].float() + 0.3 * torch.randn(16, 32, generator=g)).half())
    want = {}
    for k, seq in xs.items():
        base = seq[0]
        for t in range(1, 4):
            x = seq[t]
            recv = (base.float() + 0.5 * (x.float() - base.float())).half()
            payload = torch.zeros(x.numel() // 8, dtype=torch.half)
            st.log(k, base, None, x, recv, payload, 1)
            want[(k, t - 1)] = dict(error=float(torch.norm(x.double() - recv.double())),
                                    activation_norm=float(torch.norm(x.double())),
                                    delta_norm=float(torch.norm(x.double() - base.double())),
                                    prev=seq[t - 1] if t > 1 else None, x=x)
            base = recv
    logger = st.stats_log()
    assert logger._dirty
    stats = logger.stats
    assert not logger._dirty and set(stats) == set(xs)
    for (k, i), w in want.items():
        r = stats[k][i]
        for name in ("error", "activation_norm", "delta_norm"):
            assert abs(r[name] - w[name]) <= 1e-5 * max(1.0, w[name]), (k, i, name)
        assert r["residual"] == 1 and r["original_size_bytes"] == 16 * 32 * 2 and r["compressed_size_bytes"] == 128
        assert r["total_error"] is None and r["delta_delta_norm"] is None
        assert abs(r["rel_l2"] - w["error"] / w["activation_norm"]) < 1e-6
        if w["prev"] is None:
            assert r["activation_similarity"] is None and r["delta_before_feedback_norm"] is None
        else:
            cos = float(torch.nn.functional.cosine_similarity(w["x"].double().flatten(), w["prev"].double().flatten(), dim=0))
            assert abs(r["activation_similarity"] - cos) < 1e-5
            assert abs(r["delta_before_feedback_norm"] - float(torch.norm(w["x"].double() - w["prev"].double()))) < 1e-4
    assert logger.total_original_volume == 6 * 1024 and logger.total_compressed_volume == 6 * 128
    st.stats_verbose()
    st.stats_verbose_steps(steps=[0, 7], keys=["0-0-k"])
    out = capsys.readouterr().out
    assert "Ratio 8.00x" in out and "avg comp error" in out and "Step 7 is out of range" in out and "=== Step 0 ===" in out
    d = st.dump_err_vs_steps(str(tmp_path))
    saved = torch.load(os.path.join(tmp_path, "average_error_vs_steps.pt"))
    assert saved == d and saved["steps"] == [0, 1, 2] and saved["avg_total_errors"] == [None] * 3
    assert abs(saved["avg_comp_errors"][1] - (want[("0-0-k", 1)]["error"] + want[("0-0-v", 1)]["error"]) / 2) < 1e-4
    n = st.dump_norms_sim_vs_steps(str(tmp_path))
    assert n["avg_act_similarities"][0] is None and n["avg_act_similarities"][1] is not None
    assert os.path.exists(os.path.join(tmp_path, "average_norms_and_similarity_vs_steps.pt"))
    # residual 2 adds the delta-delta norm; a None reconstruction leaves the error fields empty
    st.stats_clear()
    x, base, db = xs["0-0-k"][1], xs["0-0-k"][0], (0.1 * xs["0-0-v"][0].float()).half()
    st.log("1-0-k", base, db, x, None, None, 2)
    r = st.stats_log().stats["1-0-k"][0]
    assert r["error"] is None and r["compressed_size_bytes"] == 0
    assert abs(r["delta_delta_norm"] - float(torch.norm(x.double() - (base + db).double()))) < 1e-4
    assert abs(r["activation_norm"] - float(torch.norm(x.double()))) < 1e-4
    with pytest.raises(ValueError):
        st.log("1-0-k", base, None, x, x, None, 3)
    st.stats_clear()
    st.stats_verbose()
    assert "No statistics logged." in capsys.readouterr().out


def test_header_is_plain_c(tmp_path):
    """include/compactb200.h is the drop-in boundary: it must compile as C99 and as C++ with no warnings and
    without CUDA or torch headers on the include path."""
    src = tmp_path / "t.c"
    src.write_text('#include "compactb200.h"\nint main(void) { return cf_abi_version() != CF_ABI_VERSION; }\n')
    inc = os.path.join(ROOT, "include")
    for cmd in (["gcc", "-std=c99"], ["g++", "-x", "c++", "-std=c++17"]):
        r = subprocess.run(cmd + ["-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-c", str(src), "-o",
                                  str(tmp_path / "t.o")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def _build_c_client(tmp_path, built_lib):
    exe = str(tmp_path / "c_client")
    libdir = os.path.dirname(built_lib)
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           "-I", "/usr/local/cuda/include", os.path.join(ROOT, "examples", "c_client.c"), "-o", exe,
           "-L", libdir, "-lcompactb200", f"-Wl,-rpath,{libdir}", "-L", "/usr/local/cuda/lib64", "-lcudart", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_plain_c_client_links_and_its_checker_agrees_with_the_oracle(tmp_path, built_lib):
    """examples/c_client.c (C99, no torch): links against the shared library, the device-free calls work, and
    its bit-exact property checker accepts the ORACLE's BINARY round trip and rejects corrupted ones -- so the
    GPU run of the same program (tests/test_gpu_zy_consumers_stats_abi.py) is judged by a checked checker."""
    from oracle import codecs as oc
    exe = _build_c_client(tmp_path, built_lib)
    r = subprocess.run([exe, "--no-gpu"], capture_output=True, text=True)
    assert r.returncode == 0 and "C_NO_GPU_OK abi=1" in r.stdout, r.stdout + r.stderr
    n, c = 48, 256
    g = torch.Generator().manual_seed(2)
    x = (torch.randn(n, c, generator=g) * 3).half()
    base = (x.float() + 0.4 * torch.randn(n, c, generator=g)).half()
    base[0, :8] = x[0, :8]                      # zero deltas: sign bit 1
    x[1, 0], base[1, 0] = 6.0e-8, 1.2e-7        # subnormal operands
    packed, u, v, _ = oc.binary_quant(x, base, False)
    recon = oc.binary_dequant(packed, u, v, base)
    payload = packed.tobytes() + u.numpy().tobytes() + v.numpy().tobytes()

    def dump(path, recon_t, payload_b):
        with open(path, "wb") as f:
            f.write(np.array([n, c], dtype=np.int64).tobytes())
            f.write(x.numpy().tobytes() + base.numpy().tobytes() + payload_b + recon_t.numpy().tobytes())

    good = str(tmp_path / "good.bin")
    dump(good, recon, payload)
    r = subprocess.run([exe, "--check", good], capture_output=True, text=True)
    assert r.returncode == 0 and "C_CHECK_OK" in r.stdout, r.stdout + r.stderr
    bad_recon = recon.clone()
    bad_recon.view(torch.int16)[5, 7] ^= 1      # one ulp off
    dump(str(tmp_path / "bad1.bin"), bad_recon, payload)
    r = subprocess.run([exe, "--check", str(tmp_path / "bad1.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "recon (5,7)" in r.stderr
    flipped = bytearray(payload)
    flipped[3 * (c // 8) + 2] ^= 0x10           # sign bit of element (3, 20)
    dump(str(tmp_path / "bad2.bin"), recon, bytes(flipped))
    r = subprocess.run([exe, "--check", str(tmp_path / "bad2.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "sign bit (3,20)" in r.stderr


@pytest.mark.skipif(not os.path.isdir("/root/reference/xfuser"), reason="reference tree not mounted (GPU box)")
def test_every_import_of_the_plugin_in_the_reference_tree_resolves():
    """Data-driven version of the test above: walk the reference checkout (xDiT hooks, examples, benchmark
    scripts and the reference's own tests), collect every `from xfuser.compact... import a, b` and
    `from xfuser.prof import ...` outside the plugin itself, and require each name from this package after
    shim.install()."""
    import ast
    import importlib
    import warnings
    import compactfusion_b200.shim as shim
    wanted = {}
    for top, _, files in os.walk("/root/reference"):
        if "/xfuser/compact" in top.replace("\\", "/"):
            continue
        for fn in files:
            if not fn.endswith(".py"):
                continue
            path = os.path.join(top, fn)
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    tree = ast.parse(open(path, encoding="utf-8", errors="ignore").read())
            except SyntaxError:
                continue
            for node in ast.walk(tree):
                if isinstance(node, ast.ImportFrom) and node.module and (
                        node.module.startswith("xfuser.compact") or node.module == "xfuser.prof"):
                    for a in node.names:
                        wanted.setdefault(node.module, {})[a.name] = os.path.relpath(path, "/root/reference")
    assert sum(len(v) for v in wanted.values()) > 40, wanted
    # plot.py is the matplotlib figure code of the stats module: out of scope, no user outside the plugin
    shim.install()
    try:
        missing = []
        for mod, names in sorted(wanted.items()):
            try:
                m = importlib.import_module(mod)
            except ImportError:
                missing += [f"{mod} (module; e.g. {next(iter(names.values()))})"]
                continue
            assert m.__name__.startswith("compactfusion_b200"), mod
            missing += [f"{mod}.{n} ({src})" for n, src in names.items() if n != "*" and not hasattr(m, n)]
        assert not missing, "\n".join(missing)
    finally:
        shim.uninstall()


def test_profiler_api_matches_the_reference_semantics():
    """The profiler calls the reference's own tests make (tests/compact/prof_test.py), on CPU sections:
    accumulation over repeated scopes, (total, avg) read-out, enable / disable, decorator, summary lines."""
    import time
    from compactfusion_b200.prof import Profiler, prof_summary, set_torch_profiler, torch_profiler_step
    p = Profiler.instance()
    assert p is Profiler.instance()
    p.reset()
    assert Profiler().enabled is True           # like the reference's: recording from the start
    p.disable()
    p.start("ignored", cpu=True)
    p.stop("ignored", cpu=True)
    assert p.events == {}                       # disabled: nothing is recorded
    p.enable()
    try:
        with Profiler.scope("no_cuda_here"):    # a CUDA section on a host without CUDA: wall clock, no raise
            time.sleep(0.001)
        assert p.elapsed_time("no_cuda_here")[0] >= 1.0
        p.reset()
        for _ in range(3):
            with Profiler.scope("total", cpu=True):
                with Profiler.scope("inner", cpu=True):
                    time.sleep(0.005)
        total, avg = p.elapsed_time("inner")
        assert 12.0 < total < 5000.0 and abs(avg - total / 3) < 1e-9
        assert p.elapsed_time("inner") == (total, avg)   # idempotent once folded

        @Profiler.prof_func("decorated", cpu=True)
        def work(x):
            time.sleep(0.002)
            return x + 1

        assert work(1) == 2 and work(2) == 3
        totals, avgs = p.get_all_elapsed_times()
        assert set(totals) == {"total", "inner", "decorated"} and totals["total"] >= totals["inner"]
        assert abs(avgs["decorated"] - totals["decorated"] / 2) < 1e-9
        lines = prof_summary(p, rank=0)
        assert isinstance(lines, list) and any("[total]" in ln and "100.00%" in ln for ln in lines)
        assert lines.index(next(ln for ln in lines if "[total]" in ln)) < lines.index(next(ln for ln in lines if "[inner]" in ln))
        with pytest.raises(AssertionError):
            p.stop("inner", cpu=True)           # stop without start
        p.start("inner", cpu=True)
        with pytest.raises(AssertionError):
            p.start("inner", cpu=True)          # nested start of the same section
        p.stop("inner", cpu=True)
        with pytest.raises(ValueError):
            p.elapsed_time("never")
        p.disable()
        work(5)
        assert p.elapsed_time("decorated")[0] == totals["decorated"]
    finally:
        p.enable()
        p.reset()

    class Stepper:
        n = 0

        def step(self):
            self.n += 1

    s = Stepper()
    torch_profiler_step()                       # no profiler installed: no-op
    set_torch_profiler(s)
    torch_profiler_step()
    set_torch_profiler(None)
    torch_profiler_step()
    assert s.n == 1


@pytest.mark.skipif(not os.path.isfile("/root/reference/examples/configs.py"), reason="reference tree not mounted (GPU box)")
def test_reference_example_presets_construct_against_this_package(capsys):
    """Execute the reference's own examples/configs.py (the presets its example scripts and benchmarks use:
    binary, int2, lowrank*, distrifusion, patch, int2patch ...) on top of the shim: every preset must build a
    valid CompactConfig from THIS package, and init / compress-type selection must work as the pipelines use
    them (compact_init -> compress_func(layer, step) -> get_compress_type)."""
    import compactfusion_b200 as cf
    import compactfusion_b200.shim as shim
    shim.install()
    try:
        ns = {"__name__": "ref_configs"}
        exec(compile(open("/root/reference/examples/configs.py").read(), "configs.py", "exec"), ns)
        methods = ["binary", "int2", "lowrank12", "lowrank8", "lowrankq32", "df", "pipe", "ring", "ulysses", "int2patch"]
        built = {}
        for model in ("Flux", "Pixart-alpha", "CogVideoX"):
            for method in methods:
                cfg = ns["get_config"](model, method)
                assert isinstance(cfg, cf.CompactConfig), (model, method)
                built[(model, method)] = cfg
                cf.compact_init(cfg)
                assert cf.compact_config() is cfg and cf.compact_get_step() is None
                if cfg.enabled and cfg.compress_func is not None:
                    warm = 2 if model == "CogVideoX" else 1
                    assert cfg.compress_func(0, 0) == cf.COMPACT_COMPRESS_TYPE.WARMUP
                    assert cfg.compress_func(3, warm) != cf.COMPACT_COMPRESS_TYPE.WARMUP
                    assert isinstance(cfg.get_compress_type(), str)
                cf.compact_reset()
        assert built[("Flux", "binary")].fastpath and built[("Flux", "binary")].compress_func(0, 5) == cf.COMPACT_COMPRESS_TYPE.BINARY
        assert built[("CogVideoX", "lowrankq32")].comp_rank == 32
        assert built[("Flux", "int2patch")].override_with_patch_gather_fwd
        assert built[("Flux", "int2patch")].patch_gather_fwd_config.use_compact
        assert not built[("Flux", "ring")].enabled
        with pytest.raises(ValueError):
            ns["get_config"]("SDXL", "binary")
    finally:
        shim.uninstall()
    capsys.readouterr()


def test_integration_md_stub_binds_against_the_library(built_lib):
    """The ctypes stub INTEGRATION.md shows a reference maintainer is real code: it loads this library, its
    prototypes agree with the binding the package itself uses, and it defines the reference's entry points."""
    from compactfusion_b200 import _native as nv
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", md, flags=re.S)
    stub = next(b for b in blocks if "_compactb200.py" in b)
    stub = stub.replace('ctypes.CDLL("libcompactb200.so")', f'ctypes.CDLL({built_lib!r})')
    ns = {}
    exec(compile(stub, "INTEGRATION.md:_compactb200.py", "exec"), ns)
    for fn in ("binary_quant_fastpath", "binary_dequant_fastpath"):
        assert callable(ns[fn])
    lib = ns["_lib"]
    for name in ("cf_binary_compress", "cf_int2_compress", "cf_binary_decompress", "cf_int2_decompress",
                 "cf_workspace_bytes"):
        res, args = nv.SYMBOLS[name]
        assert list(getattr(lib, name).argtypes) == list(args), name
        assert getattr(lib, name).restype is res, name
    assert ns["_lib"].cf_workspace_bytes(1, 4608, 3072, 0, 1) == nv.workspace_bytes(nv.CODEC_BINARY, 4608, 3072)


def test_quantized_cache_host_logic(monkeypatch):
    """CompactCache(quantize=True): deprecated gate, int8 storage tuple, dequantise-on-read -- with the codec
    calls replaced by the oracle IN THE TEST ONLY."""
    from compactfusion_b200 import compress_quantize as cq
    from compactfusion_b200 import utils
    from oracle import codecs as oc
    monkeypatch.setattr(utils, "ALLOW_DEPRECATED", False)
    with pytest.raises(AssertionError):
        utils.CompactCache(quantize=True)
    with pytest.raises(AssertionError):
        utils.CompactConfig(enabled=True, residual=1, ef=True, quantized_cache=True)
    monkeypatch.setattr(utils, "ALLOW_DEPRECATED", True)
    assert utils.CompactConfig(enabled=True, residual=1, ef=True, quantized_cache=True).quantized_cache
    monkeypatch.setattr(cq, "quantize_int8", lambda t: tuple(oc.int8_quantize(t)))
    monkeypatch.setattr(cq, "dequantize_int8", lambda q, s, z: oc.int8_dequantize(q, s, z))
    g = torch.Generator().manual_seed(1)
    x = torch.randn(64, 128, generator=g).half()
    cache = utils.CompactCache(quantize=True)
    cache.put("0-0-k", x, x * 0)
    stored = cache.base["0-0-k"]
    assert isinstance(stored, tuple) and len(stored) == 4 and stored[3] == x.shape
    got = cache.get_base("0-0-k")
    assert got.shape == x.shape and got.dtype == torch.half
    assert torch.equal(got, oc.int8_dequantize(*oc.int8_quantize(x)))
    assert float((got.float() - x.float()).norm() / x.float().norm()) < 2e-2
    assert cache.get_delta_base("0-0-k") is not None and cache.get_base("nope") is None
    plain = utils.CompactCache()
    plain.put("k", x, None)
    assert plain.get_base("k") is x


# ---------------- dropin.hot: the hooks' one-probe path (dropin.py) ----------------
def test_dropin_hot_caches_per_layer_and_respects_type_and_config(monkeypatch):
    """`hot` = `usable` + `lookup` once per (hook, group, layer, shape); afterwards one dictionary probe that still
    refuses a compress type the engines do not serve and re-decides for a new configuration object."""
    from compactfusion_b200 import dropin
    from compactfusion_b200.utils import COMPACT_COMPRESS_TYPE as T
    dropin.shutdown()
    calls = {"usable": 0, "lookup": 0}

    class Cfg:
        comp_rank = -1

    def usable(cfg, ctype, k):
        calls["usable"] += 1
        return ctype in dropin._ENGINE_TYPES

    def lookup(kind, group, k, mod_idx, comp_rank=None):
        calls["lookup"] += 1
        return ("engine", 7, None)

    class K:   # what `hot` reads of a tensor
        shape, dtype, is_cuda = (1, 576, 24, 128), torch.half, True

    monkeypatch.setattr(dropin, "usable", usable)
    monkeypatch.setattr(dropin, "lookup", lookup)
    monkeypatch.setattr(dropin, "_config_ok", lambda cfg: dropin._ENGINE_TYPES)
    cfg = Cfg()
    assert dropin.hot(cfg, "patch", None, K, 3, T.SPARSE) is None and calls == {"usable": 1, "lookup": 0}
    ent = dropin.hot(cfg, "patch", None, K, 3, T.BINARY)
    assert ent[:3] == ("engine", 7, None) and calls == {"usable": 2, "lookup": 1}
    for ct in (T.BINARY, T.INT2, T.WARMUP):
        assert dropin.hot(cfg, "patch", None, K, 3, ct) is ent
    assert dropin.hot(cfg, "patch", None, K, 3, T.SPARSE) is None   # served types are checked on the cached entry too
    assert calls == {"usable": 2, "lookup": 1}
    assert dropin.hot(cfg, "patch", None, K, 4, T.BINARY)[1] == 7 and calls["lookup"] == 2   # another layer: its own entry
    other = Cfg()
    dropin.hot(other, "patch", None, K, 3, T.BINARY)                                          # new configuration object
    assert calls["usable"] == 4 and calls["lookup"] == 3
    dropin.shutdown()
    assert not dropin._hot and not dropin._fast


# ---------------- the INT4 encode's exact-quotient shortcut (csrc/cf_minmax_codecs.cu, int4_codes2) ----------------
def test_int4_bracketed_quotient_never_accepts_a_wrong_fp16_value():
    """RN16 of both ends of [t (1 - 2^-21), t (1 + 2^-21)], t = RN32(a * RN32(1 / s)): whenever they agree they equal
    the reference's fp16(a / s) -- every finite fp16 numerator against 60 scales here (tools/check_int4_bracket.py runs
    3000), and the division is needed for well under 1 % of the elements."""
    import numpy as np
    a = np.arange(0, 0x7C00, dtype=np.uint16).view(np.float16).astype(np.float32)
    rng = np.random.default_rng(1)
    s_bits = rng.integers(1, 0x7C00, size=60, dtype=np.uint16)
    s_bits[:6] = [1, 0x03FF, 0x0400, 0x3C00, 0x7BFF, 0x2E66]
    k_lo, k_hi = np.float32(1.0) - np.float32(2.0 ** -21), np.float32(1.0) + np.float32(2.0 ** -21)
    wrong = fallback = 0
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        for sb in s_bits:
            s = np.array([sb], dtype=np.uint16).view(np.float16).astype(np.float32)[0]
            ref = (a / s).astype(np.float16).view(np.uint16)
            t = a * (np.float32(1.0) / s)
            lo, hi = (t * k_lo).astype(np.float16).view(np.uint16), (t * k_hi).astype(np.float16).view(np.uint16)
            same = lo == hi
            wrong += int((lo[same] != ref[same]).sum())
            fallback += int((~same).sum())
    assert wrong == 0
    assert fallback < 0.01 * a.size * len(s_bits)
