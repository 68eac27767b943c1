"""Execute the SOURCE of the min/max codec kernels (k_minmax_stats, k_minmax_finalize, k_int4_codec of
csrc/cf_minmax_codecs.cu) on the CPU (tests/cuda_emulation.py) against what the REFERENCE produced on the
committed inputs (tests/golden/codecs.npz): packed INT4 codes, scale, min, dequantised values and sim_int4 --
and the 4-level instantiation (sim_int2_minmax) that was added after the round's GPU budget was spent -- all
bit-exact, including the reciprocal-quotient fast path with its tie check (quot_for_rn16)."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import cuda_emulation as emu
from conftest import GOLDEN, bits16, h16

RUNNER = r'''
template <int MODE>
static int run(const __half* x, int N, int C, int rows_per_cta) {
  using namespace cf;
  constexpr int LEVELS = MmLevels<MODE>::value;
  const RowGeom g = make_row_geom(C);
  if (g.G != 1) { fprintf(stderr, "this runner instantiates G = 1 only\n"); return 3; }
  const int B = (N + rows_per_cta - 1) / rows_per_cta;
  std::vector<__half> pmin((size_t)B * C), pmax((size_t)B * C), scale(C), minv(C), recon((size_t)N * C), deq((size_t)N * C);
  std::vector<uint8_t> packed((size_t)(N / 2) * C, 0xEE);
  launch(B, 1, g.TX, g.TY, [&] { k_minmax_stats<1>(x, nullptr, pmin.data(), pmax.data(), N, C, rows_per_cta); });
  launch((C + 31) / 32, 1, 256, 1, [&] { k_minmax_finalize<MODE>(pmin.data(), pmax.data(), B, C, scale.data(), minv.data(), nullptr); });
  launch(2, 1, g.TX, g.TY, [&] { k_int4_codec<1, true, LEVELS>(x, nullptr, scale.data(), minv.data(), packed.data(), recon.data(), N, C); });
  launch(3, 1, g.TX, g.TY, [&] { k_int4_codec<1, false, LEVELS>(nullptr, nullptr, scale.data(), minv.data(), packed.data(), deq.data(), N, C); });
  fwrite(packed.data(), 1, packed.size(), stdout);
  fwrite(scale.data(), 2, C, stdout);
  fwrite(minv.data(), 2, C, stdout);
  fwrite(recon.data(), 2, recon.size(), stdout);
  fwrite(deq.data(), 2, deq.size(), stdout);
  return 0;
}

static int run_int8(const __half* x, int N, int C, int rows_per_cta) {
  using namespace cf;
  const RowGeom g = make_row_geom(C);
  if (g.G != 1) return 3;
  const int B = (N + rows_per_cta - 1) / rows_per_cta;
  std::vector<__half> pmin((size_t)B * C), pmax((size_t)B * C), scale(C), min_ws(C), recon((size_t)N * C), deq((size_t)N * C);
  std::vector<int16_t> zp(C);
  std::vector<int8_t> q((size_t)N * C, 77);
  launch(B, 1, g.TX, g.TY, [&] { k_minmax_stats<1>(x, nullptr, pmin.data(), pmax.data(), N, C, rows_per_cta); });
  launch((C + 31) / 32, 1, 256, 1, [&] { k_minmax_finalize<MODE_INT8>(pmin.data(), pmax.data(), B, C, scale.data(), zp.data(), min_ws.data()); });
  launch(2, 1, g.TX, g.TY, [&] { k_int8_codec<1, true>(x, nullptr, scale.data(), zp.data(), q.data(), recon.data(), N, C); });
  launch(3, 1, g.TX, g.TY, [&] { k_int8_codec<1, false>(nullptr, nullptr, scale.data(), zp.data(), q.data(), deq.data(), N, C); });
  fwrite(q.data(), 1, q.size(), stdout);
  fwrite(scale.data(), 2, C, stdout);
  fwrite(zp.data(), 2, C, stdout);
  fwrite(recon.data(), 2, recon.size(), stdout);
  fwrite(deq.data(), 2, deq.size(), stdout);
  return 0;
}

// the generic path (C % 8 != 0: the tall-skinny low-rank factors of LOW_RANK_Q, slowpath.py:69-70)
static int run_int4_generic(const __half* x, int N, int C) {
  using namespace cf;
  std::vector<__half> scale(C), minv(C), recon((size_t)N * C), deq((size_t)N * C);
  std::vector<uint8_t> packed((size_t)(N / 2) * C, 0xEE);
  launch(C, 1, 256, 1, [&] { k_minmax_column_generic<MODE_INT4>(x, nullptr, N, C, scale.data(), minv.data()); });
  launch(2, 1, 256, 1, [&] { k_int4_codec_generic<true>(x, nullptr, scale.data(), minv.data(), packed.data(), recon.data(), N, C); });
  launch(3, 1, 256, 1, [&] { k_int4_codec_generic<false>(nullptr, nullptr, scale.data(), minv.data(), packed.data(), deq.data(), N, C); });
  fwrite(packed.data(), 1, packed.size(), stdout);
  fwrite(scale.data(), 2, C, stdout);
  fwrite(minv.data(), 2, C, stdout);
  fwrite(recon.data(), 2, recon.size(), stdout);
  fwrite(deq.data(), 2, deq.size(), stdout);
  return 0;
}

int main(int argc, char** argv) {  // <int4|int2mm|int8|int4g> x.bin N C rows_per_cta
  const std::string mode = argv[1];
  auto x = slurp(argv[2]);
  const int N = atoi(argv[3]), C = atoi(argv[4]), rpc = atoi(argv[5]);
  const __half* xh = reinterpret_cast<const __half*>(x.data());
  if (mode == "int8") return run_int8(xh, N, C, rpc);
  if (mode == "int4g") return run_int4_generic(xh, N, C);
  return mode == "int4" ? run<cf::MODE_INT4>(xh, N, C, rpc) : run<cf::MODE_INT2MM>(xh, N, C, rpc);
}
'''


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    src = open(os.path.join(emu.CSRC, "cf_minmax_codecs.cu")).read()
    kern = re.search(r"(namespace cf \{.*?)struct MinMaxPlan", src, flags=re.S).group(1) + "}  // namespace cf\n"
    assert "k_int4_codec" in kern and "<<<" not in kern and "asm" not in kern
    kern = kern.replace("extern __shared__ uint32_t sm_u32[];", "uint32_t* sm_u32 = reinterpret_cast<uint32_t*>(emu_smem);")
    d = tmp_path_factory.mktemp("minmax_emu")
    return emu.build(d, emu.SHIM_HEAD + emu.common_source() + kern + emu.SLURP + RUNNER), d


def _run(emulator, mode, d16, rows_per_cta):
    exe, d = emulator
    n, c = d16.shape
    (d / "x.bin").write_bytes(d16.numpy().tobytes())
    r = subprocess.run([exe, mode, str(d / "x.bin"), str(n), str(c), str(rows_per_cta)], capture_output=True, timeout=900)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    sizes = [(n if mode == "int8" else n // 2) * c, 2 * c, 2 * c, 2 * n * c, 2 * n * c]
    assert len(r.stdout) == sum(sizes)
    parts, o = [], 0
    for s in sizes:
        parts.append(r.stdout[o:o + s])
        o += s
    u16 = lambda b, shape: np.frombuffer(b, dtype=np.uint16).reshape(shape)  # noqa: E731
    return (np.frombuffer(parts[0], dtype=np.uint8).reshape(-1, c), u16(parts[1], (1, c)), u16(parts[2], (1, c)),
            u16(parts[3], (n, c)), u16(parts[4], (n, c)))


@pytest.mark.parametrize("name,rows_per_cta", [("rand_64x256", 20), ("rand_48x1152", 48), ("rand_130x64", 33),
                                               ("flux_k_96x512", 40)])
def test_int4_and_int2_minmax_kernel_source_match_the_reference_goldens(emulator, name, rows_per_cta):
    g = np.load(os.path.join(GOLDEN, "codecs.npz"))
    d16 = (h16(g[f"{name}/x"]) - h16(g[f"{name}/base"])).contiguous()
    packed, scale, mn, recon, deq = _run(emulator, "int4", d16, rows_per_cta)
    assert np.array_equal(packed, g[f"{name}/int4_packed"]), "INT4 codes differ from the reference"
    assert np.array_equal(scale, bits16(h16(g[f"{name}/int4_scale"])).reshape(scale.shape))
    assert np.array_equal(mn, bits16(h16(g[f"{name}/int4_min"])).reshape(mn.shape))
    assert np.array_equal(deq, bits16(h16(g[f"{name}/int4_deq"])).reshape(deq.shape)), "dequantised values differ"
    assert np.array_equal(recon, bits16(h16(g[f"{name}/sim_int4_d0"])).reshape(recon.shape)), "sim_int4 differs"
    # the 4-level instantiation: the reference's sim_int2_minmax
    packed2, _, mn2, recon2, deq2 = _run(emulator, "int2mm", d16, rows_per_cta)
    assert np.array_equal(recon2, bits16(h16(g[f"{name}/sim_int2_minmax"])).reshape(recon2.shape)), "sim_int2_minmax differs"
    assert np.array_equal(recon2, deq2) and np.array_equal(mn2, mn)
    assert int((packed2 & 0x0F).max()) <= 3 and int((packed2 >> 4).max()) <= 3
    # INT8 (per-channel affine, int16 zero point)
    q8, s8, zp8, recon8, deq8 = _run(emulator, "int8", d16, rows_per_cta)
    assert np.array_equal(q8.view(np.int8), g[f"{name}/int8_q"]), "INT8 codes differ from the reference"
    assert np.array_equal(s8, bits16(h16(g[f"{name}/int8_scale"])).reshape(s8.shape))
    assert np.array_equal(zp8.view(np.int16), g[f"{name}/int8_zp"].reshape(zp8.shape))
    assert np.array_equal(deq8, bits16(h16(g[f"{name}/int8_deq"])).reshape(deq8.shape)) and np.array_equal(recon8, deq8)


@pytest.mark.parametrize("n,c", [(64, 12), (130, 20), (96, 33)])
def test_generic_int4_kernel_source_matches_the_oracle(emulator, n, c):
    """C % 8 != 0 (LOW_RANK_Q quantises its (N, r) / (C, r) factors): the scalar kernels, bit-exact against the
    oracle's quantize_int4 / dequantize_int4 restatement (itself pinned to the reference's goldens)."""
    from oracle import codecs as oc
    g = torch.Generator().manual_seed(n + c)
    x = (torch.randn(n, c, generator=g) * torch.rand(1, c, generator=g) * 3).half()
    packed, scale, mn, recon, deq = _run(emulator, "int4g", x, 1)
    o_packed, o_scale, o_min = oc.int4_quantize(x)
    assert np.array_equal(packed, o_packed), "codes differ"
    assert np.array_equal(scale, bits16(o_scale).reshape(scale.shape)) and np.array_equal(mn, bits16(o_min).reshape(mn.shape))
    want = bits16(oc.int4_dequantize(o_packed, o_scale, o_min))
    assert np.array_equal(deq, want.reshape(deq.shape)) and np.array_equal(recon, deq)
