"""Execute the SOURCE of the pipelined (bulk-async / mbarrier) BINARY and INT2 kernels -- the default hot path:
k_delta_stats_tma, k_int2_encode_tma, k_apply_codes_tma of csrc/cf_sign_tma.cuh, with k_finalize_scales between
them -- on the CPU and hold the result to the GPU parity bars against the oracle.

On top of tests/cuda_emulation.py's thread / fp16 shim, cf_pipe.cuh's inline PTX is replaced by stand-ins with
the same names: an mbarrier is a (pending arrivals, pending bytes, phase) record, `cp.async.bulk` is a memcpy
that then completes its byte count on the barrier (and asserts the 16-byte alignment / size rules), the
shared-window loads index the emulated shared-memory image, the named compute barrier and __syncwarp are
pthread barriers.  The producer lane, the stage ring with its full / empty barriers and phase parities, the
ragged last tiles, the per-row scale staging and the flattened tile schedule all execute as written; what a
CPU cannot tell is how fast it runs or whether the async-proxy fences are sufficient on hardware.
"""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import cuda_emulation as emu
from oracle import codecs as oc
from test_codec_kernel_emulation import _inputs, _same_bits, _ulp

RUNNER = r'''
template <int MODE, int G, int OCC>
static int run(const __half* x, const __half* base, int N, int C, int n_cta_stats, int n_cta_apply) {
  using namespace cf;
  constexpr int per_code = (MODE == MODE_BINARY) ? 8 : 4;
  std::vector<uint8_t> packed((size_t)N * C / per_code, 0xAA);
  std::vector<__half> rowmean(N), U(N), V(C), new_base((size_t)N * C), recon((size_t)N * C);
  // ---- pass 1: stage = [x tile | base tile] ----
  const PipeGeom g1 = make_pipe_geom(C, 2, 0, false);
  if (!g1.ok || g1.G != G || g1.ctas_per_sm != OCC) { fprintf(stderr, "stats geometry: ok=%d G=%d occ=%d\n", g1.ok, g1.G, g1.ctas_per_sm); return 3; }
  int64_t rpc = (N + n_cta_stats - 1) / n_cta_stats;
  rpc = (rpc + g1.R - 1) / g1.R * g1.R;
  const int B = (int)((N + rpc - 1) / rpc);
  std::vector<float> tokpart(B), colpart((size_t)B * C);
  auto args = [](const PipeGeom& g, int rows_per_cta) {
    PipeArgs a{};
    a.TX = g.TX; a.TY = g.TY; a.R = g.R; a.stages = g.stages; a.chunk_rows = g.chunk_rows; a.u_cap = g.u_cap;
    a.tile_bytes = g.tile_bytes; a.stage_bytes = g.stage_bytes; a.rows_per_cta = rows_per_cta; a.l2_hints = 1; a.early_load = 1;
    return a;
  };
  auto sched = [](const PipeGeom& g, int N, int want_ctas, int* n_cta) {
    TileSched ts{};
    ts.tiles_per_tensor = (N + g.R - 1) / g.R;
    ts.total_tiles = ts.tiles_per_tensor;
    int ctas = want_ctas > ts.total_tiles ? ts.total_tiles : want_ctas;
    ts.tiles_per_cta = (ts.total_tiles + ctas - 1) / ctas;
    const int cap = g.u_cap / g.R > 0 ? g.u_cap / g.R : 1;
    if (ts.tiles_per_cta > cap) ts.tiles_per_cta = cap;
    *n_cta = (ts.total_tiles + ts.tiles_per_cta - 1) / ts.tiles_per_cta;
    return ts;
  };
  StatsParams sp{};
  sp.x[0] = x; sp.base[0] = base; sp.packed[0] = packed.data(); sp.rowmean[0] = rowmean.data();
  sp.tokpart[0] = tokpart.data(); sp.colpart[0] = colpart.data(); sp.N = N; sp.C = C; sp.rows_per_cta = (int)rpc;
  const PipeArgs a1 = args(g1, (int)rpc);
  emu_ncompute = g1.TX * g1.TY;
  launch(B, 1, g1.TX * g1.TY + 32, 1, [&] { k_delta_stats_tma<MODE, G, OCC, false>(sp, a1, FanOut{}); });
  FinalizeParams fp{};
  fp.rowmean[0] = rowmean.data(); fp.tokpart[0] = tokpart.data(); fp.colpart[0] = colpart.data();
  fp.scale_u[0] = U.data(); fp.scale_v[0] = V.data(); fp.N = N; fp.C = C; fp.B = B;
  launch((C + 31) / 32, 1, 1024, 1, [&] { k_finalize_scales<MODE, false>(fp, FanOut{}, 0); });
  if (MODE == MODE_INT2) {   // second pass over x / base: codes from the final scales (+ error-feedback base)
    const PipeGeom g2 = make_pipe_geom(C, 2, 0, true);
    if (!g2.ok || g2.G != G || g2.ctas_per_sm != OCC) { fprintf(stderr, "encode geometry\n"); return 3; }
    int n_cta = 0;
    const TileSched ts = sched(g2, N, n_cta_apply, &n_cta);
    Int2EncodeParams ep{};
    ep.x[0] = x; ep.base[0] = base; ep.scale_u[0] = U.data(); ep.scale_v[0] = V.data(); ep.packed[0] = packed.data();
    ep.new_base[0] = new_base.data(); ep.N = N; ep.C = C;
    const PipeArgs a2 = args(g2, 0);
    emu_ncompute = g2.TX * g2.TY;
    launch(n_cta, 1, g2.TX * g2.TY + 32, 1, [&] { k_int2_encode_tma<G, OCC, false>(ep, a2, ts, FanOut{}); });
  }
  // ---- apply: stage = [base tile | code tile] ----
  const PipeGeom g3 = make_pipe_geom(C, 1, C / per_code, true);
  if (!g3.ok || g3.G != G || g3.ctas_per_sm != OCC) { fprintf(stderr, "apply geometry\n"); return 3; }
  ApplyParams ap{};
  ap.packed[0] = packed.data(); ap.scale_u[0] = U.data(); ap.scale_v[0] = V.data(); ap.base[0] = base;
  ap.N = N; ap.C = C; ap.K = 1;
  const PipeArgs a3 = args(g3, 0);
  emu_ncompute = g3.TX * g3.TY;
  if (MODE == MODE_BINARY) {
    int n_cta = 0;
    const TileSched ts = sched(g3, N, n_cta_apply, &n_cta);
    ap.recon[0] = new_base.data();
    launch(n_cta, 1, g3.TX * g3.TY + 32, 1, [&] { k_apply_codes_tma<MODE, G, OCC>(ap, a3, ts); });
  }
  int n_cta = 0;
  const TileSched ts = sched(g3, N, n_cta_apply + 1, &n_cta);
  ap.recon[0] = recon.data();
  launch(n_cta, 1, g3.TX * g3.TY + 32, 1, [&] { k_apply_codes_tma<MODE, G, OCC>(ap, a3, ts); });
  fprintf(stderr, "R=%d stages=%d TX=%d TY=%d B=%d apply_ctas=%d tiles/cta=%d\n", g3.R, g3.stages, g3.TX, g3.TY, B, n_cta, ts.tiles_per_cta);
  fwrite(packed.data(), 1, packed.size(), stdout);
  fwrite(U.data(), 2, U.size(), stdout);
  fwrite(V.data(), 2, V.size(), stdout);
  fwrite(new_base.data(), 2, new_base.size(), stdout);
  fwrite(recon.data(), 2, recon.size(), stdout);
  return 0;
}

int main(int argc, char** argv) {  // <binary|int2> x.bin base.bin N C stats_ctas apply_ctas
  const std::string mode = argv[1];
  auto x = slurp(argv[2]), b = slurp(argv[3]);
  const int N = atoi(argv[4]), C = atoi(argv[5]), sc = atoi(argv[6]), ac = atoi(argv[7]);
  const __half* xh = reinterpret_cast<const __half*>(x.data());
  const __half* bh = reinterpret_cast<const __half*>(b.data());
  const cf::PipeGeom g = cf::make_pipe_geom(C, 2, 0, false);
  if (!g.ok) { fprintf(stderr, "no pipelined geometry for C=%d\n", C); return 3; }
  if (g.G == 2) {   // C > 4096: two column groups per thread, one CTA per SM
    if (g.ctas_per_sm != 1) { fprintf(stderr, "unexpected G = 2 geometry\n"); return 3; }
    return mode == "binary" ? run<cf::MODE_BINARY, 2, 1>(xh, bh, N, C, sc, ac) : run<cf::MODE_INT2, 2, 1>(xh, bh, N, C, sc, ac);
  }
  if (mode == "binary") return g.ctas_per_sm == 2 ? run<cf::MODE_BINARY, 1, 2>(xh, bh, N, C, sc, ac) : run<cf::MODE_BINARY, 1, 1>(xh, bh, N, C, sc, ac);
  return g.ctas_per_sm == 2 ? run<cf::MODE_INT2, 1, 2>(xh, bh, N, C, sc, ac) : run<cf::MODE_INT2, 1, 1>(xh, bh, N, C, sc, ac);
}
'''


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    src = open(os.path.join(emu.CSRC, "cf_sign_codecs.cu")).read()
    kern = re.search(r"(namespace cf \{.*?\n\}  // namespace cf\n)", src, flags=re.S).group(1)
    kern = emu.strip_asm(kern).replace("extern __shared__ float smem[];", "float* smem = emu_smem;")
    tma = open(os.path.join(emu.CSRC, "cf_sign_tma.cuh")).read()
    tma = tma.replace('#pragma once', '').replace('#include "cf_pipe.cuh"', '')
    tma = emu.strip_asm(tma).replace("extern __shared__ __align__(128) unsigned char pipe_smem_raw[];",
                                     "unsigned char* pipe_smem_raw = emu_smem_bytes;")
    assert "<<<" not in tma and "asm" not in tma.replace("fastpath", "")
    hook = "#define EMU_LAUNCH_HOOK if (emu_ncompute > 0) pthread_barrier_init(&compute_bar, nullptr, emu_ncompute);\n"
    fwd = "static pthread_barrier_t compute_bar; static int emu_ncompute;\n"
    pipe_shim = emu.PIPE_SHIM.replace("static pthread_barrier_t compute_bar;\n", "").replace(
        "static int emu_ncompute = 0;   // set by the runner before a launch of a pipelined kernel\n", "")
    d = tmp_path_factory.mktemp("pipe_emu")
    text = ("#include <pthread.h>\n" + fwd + hook + emu.SHIM_HEAD + pipe_shim + emu.common_source() + kern + tma
            + emu.SLURP + RUNNER)
    return emu.build(d, text), d


def _run(emulator, mode, x, base, stats_ctas, apply_ctas):
    exe, d = emulator
    n, c = x.shape
    (d / "x.bin").write_bytes(x.numpy().tobytes())
    (d / "b.bin").write_bytes(base.numpy().tobytes())
    r = subprocess.run([exe, mode, str(d / "x.bin"), str(d / "b.bin"), str(n), str(c), str(stats_ctas), str(apply_ctas)],
                       capture_output=True, timeout=1200)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    print(mode, tuple(x.shape), r.stderr.decode().strip())  # pipeline geometry of the run (pytest -s)
    buf, per = r.stdout, (8 if mode == "binary" else 4)
    sizes = [n * c // per, 2 * n, 2 * c, 2 * n * c, 2 * n * c]
    assert len(buf) == sum(sizes), r.stderr.decode()[-500:]
    parts, o = [], 0
    for s in sizes:
        parts.append(buf[o:o + s])
        o += s
    half = lambda b, shape: torch.from_numpy(np.frombuffer(b, dtype=np.int16).copy()).view(torch.half).view(shape)  # noqa: E731
    packed = np.frombuffer(parts[0], dtype=np.uint8).reshape(n, c // per)
    return packed, half(parts[1], (n, 1)), half(parts[2], (c, 1)), half(parts[3], (n, c)), half(parts[4], (n, c))


# shapes: ragged last tiles; one CTA streaming 7 tiles through a 2-3 stage ring (stage reuse, both barrier
# parities); C = 1152 (4.5 warps of column groups: inactive lanes) next to powers of two
@pytest.mark.parametrize("n,c,stats_ctas,apply_ctas", [(300, 256, 2, 2), (77, 1152, 3, 2), (150, 512, 1, 4), (400, 256, 1, 1)])
def test_pipelined_binary_kernel_source_matches_the_oracle(emulator, n, c, stats_ctas, apply_ctas):
    x, base = _inputs(n, c, seed=7 * n)
    packed, u, v, new_base, recon = _run(emulator, "binary", x, base, stats_ctas, apply_ctas)
    o_packed, o_u, o_v, _ = oc.binary_quant(x, base, False)
    assert np.array_equal(packed, o_packed), "sign bits differ from the oracle"
    assert _ulp(u, o_u) <= 1 and _ulp(v, o_v) <= 1
    assert _same_bits(recon, new_base), "sender's error-feedback base != receiver's reconstruction"
    assert _same_bits(recon, oc.binary_dequant(o_packed, u, v, base)), "reconstruction differs given identical scales"


@pytest.mark.parametrize("n,c,stats_ctas,apply_ctas", [(300, 256, 2, 2), (77, 1152, 3, 2), (400, 256, 1, 1)])
def test_pipelined_int2_kernel_source_matches_the_oracle(emulator, n, c, stats_ctas, apply_ctas):
    x, base = _inputs(n, c, seed=11 * n)
    packed, tok, chan, new_base, recon = _run(emulator, "int2", x, base, stats_ctas, apply_ctas)
    _, o_tok, o_chan, _ = oc.int2_quant(x, base, True)
    assert _ulp(tok, o_tok) <= 1 and _ulp(chan, o_chan) <= 1
    s_packed, _, _, s_nb = oc.int2_quant(x, base, True, scales=(tok, chan))
    assert np.array_equal(packed, s_packed), "INT2 codes differ given identical scales"
    assert _same_bits(new_base, s_nb), "INT2 error-feedback base differs given identical scales"
    assert _same_bits(recon, new_base), "receiver != sender"


# edge shapes: fewer rows than one tile / one quad, a single row, the smallest C (8 column groups: 24 idle
# lanes per warp), the widest single-group C, row counts one past a tile boundary, and C > 4096 (two column
# groups per thread, one CTA per SM; 6144 leaves the second group partly idle); C = 3072 is the FLUX / CogVideoX
# geometry itself (384 compute threads, 4-row tiles, 2 stages, 2 CTAs per SM)
@pytest.mark.parametrize("mode", ["binary", "int2"])
@pytest.mark.parametrize("n,c,stats_ctas,apply_ctas", [(3, 256, 1, 1), (1, 64, 1, 1), (129, 64, 2, 3), (5, 4096, 1, 2),
                                                       (65, 128, 4, 4), (64, 256, 1, 1), (6, 6144, 1, 1),
                                                       (37, 3072, 2, 3)])
def test_pipelined_kernel_source_edge_shapes(emulator, mode, n, c, stats_ctas, apply_ctas):
    if (c // (8 if mode == "binary" else 4)) % 16:
        pytest.skip("the host dispatch (launch_apply) sends code rows that are not a multiple of 16 bytes to the "
                    "register-staged kernel: cp.async.bulk needs 16-byte sizes")
    x, base = _inputs(max(n, 2), c, seed=n * c)
    x, base = x[:n].contiguous(), base[:n].contiguous()
    packed, u, v, new_base, recon = _run(emulator, mode, x, base, stats_ctas, apply_ctas)
    if mode == "binary":
        o_packed, o_u, o_v, _ = oc.binary_quant(x, base, False)
        assert np.array_equal(packed, o_packed) and _ulp(u, o_u) <= 1 and _ulp(v, o_v) <= 1
        assert _same_bits(recon, new_base) and _same_bits(recon, oc.binary_dequant(o_packed, u, v, base))
    else:
        _, o_tok, o_chan, _ = oc.int2_quant(x, base, True)
        s_packed, _, _, s_nb = oc.int2_quant(x, base, True, scales=(u, v))
        assert _ulp(u, o_tok) <= 1 and _ulp(v, o_chan) <= 1
        assert np.array_equal(packed, s_packed) and _same_bits(new_base, s_nb) and _same_bits(recon, new_base)
