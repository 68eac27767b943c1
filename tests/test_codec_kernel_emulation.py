"""Execute the SOURCE of the register-staged BINARY / INT2 codec kernels on the CPU (TEST ONLY) and hold the
result to the same bars as the GPU parity tests: sign bits and INT2 codes bit-exact, scales within 1 fp16 ulp,
reconstruction / error-feedback base bit-exact given the kernels' own scales, sender == receiver.

k_delta_stats, k_finalize_scales, k_int2_encode and k_apply_codes (csrc/cf_sign_codecs.cu) are cut out of the
.cu file together with the device helpers of cf_common.cuh and compiled with g++ against a shim: one OS thread
per CUDA thread (2-D blocks, CTAs one after another), pthread barriers for __syncthreads, a per-warp exchange
buffer for __shfl_xor_sync, and software fp16 for the ~15 half-precision intrinsics the kernels use (every op
= exact fp32 op + one round-to-nearest-even, which is what the hardware instructions do).  The pipelined TMA
kernels share these arithmetic helpers (binary_apply8, int2_apply8, the sign / threshold masks) but not the
staging code, which only a GPU can run.

This keeps the kernels' arithmetic and packing under test in the CPU-only suite, and makes the oracle and the
kernel source check each other without a GPU in the loop.
"""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import cuda_emulation as emu
from oracle import codecs as oc


RUNNER = r'''
template <int MODE, int G>
static int run(const __half* x, const __half* base, int N, int C, int rows_per_cta) {
  using namespace cf;
  const RowGeom g = make_row_geom(C);
  if (g.G != G) { fprintf(stderr, "geometry G=%d, instantiated %d\n", g.G, G); return 3; }
  const int B = (N + rows_per_cta - 1) / rows_per_cta;
  const int per_code = (MODE == MODE_BINARY) ? 8 : 4;
  std::vector<uint8_t> packed((size_t)N * C / per_code, 0xAA);
  std::vector<__half> rowmean(N), U(N), V(C), new_base((size_t)N * C), recon((size_t)N * C);
  std::vector<float> tokpart(B), colpart((size_t)B * C);
  StatsParams sp{};
  sp.x[0] = x; sp.base[0] = base; sp.packed[0] = packed.data(); sp.rowmean[0] = rowmean.data();
  sp.tokpart[0] = tokpart.data(); sp.colpart[0] = colpart.data(); sp.N = N; sp.C = C; sp.rows_per_cta = rows_per_cta;
  launch(B, 1, g.TX, g.TY, [&] { k_delta_stats<MODE, G>(sp); });
  FinalizeParams fp{};
  fp.rowmean[0] = rowmean.data(); fp.tokpart[0] = tokpart.data(); fp.colpart[0] = colpart.data();
  fp.scale_u[0] = U.data(); fp.scale_v[0] = V.data(); fp.N = N; fp.C = C; fp.B = B;
  launch((C + 31) / 32, 1, 1024, 1, [&] { k_finalize_scales<MODE, false>(fp, FanOut{}, 0); });
  const int grid_rows = 3;  // a few row blocks: exercises the grid-stride row loop
  if (MODE == MODE_INT2) {
    Int2EncodeParams ep{};
    ep.x[0] = x; ep.base[0] = base; ep.scale_u[0] = U.data(); ep.scale_v[0] = V.data(); ep.packed[0] = packed.data();
    ep.new_base[0] = new_base.data(); ep.N = N; ep.C = C;
    launch(grid_rows, 1, g.TX, g.TY, [&] { k_int2_encode<G>(ep); });
  }
  ApplyParams ap{};
  ap.packed[0] = packed.data(); ap.scale_u[0] = U.data(); ap.scale_v[0] = V.data(); ap.base[0] = base;
  ap.N = N; ap.C = C; ap.K = 1;
  if (MODE == MODE_BINARY) {  // the sender's error-feedback update runs the receiver's kernel on its own payload
    ap.recon[0] = new_base.data();
    launch(grid_rows, 1, g.TX, g.TY, [&] { k_apply_codes<MODE, G>(ap); });
  }
  ap.recon[0] = recon.data();
  launch(grid_rows + 1, 1, g.TX, g.TY, [&] { k_apply_codes<MODE, G>(ap); });
  fwrite(packed.data(), 1, packed.size(), stdout);
  fwrite(U.data(), 2, U.size(), stdout);
  fwrite(V.data(), 2, V.size(), stdout);
  fwrite(new_base.data(), 2, new_base.size(), stdout);
  fwrite(recon.data(), 2, recon.size(), stdout);
  return 0;
}

int main(int argc, char** argv) {  // <binary|int2> x.bin base.bin N C rows_per_cta
  const std::string mode = argv[1];
  auto x = slurp(argv[2]), b = slurp(argv[3]);
  const int N = atoi(argv[4]), C = atoi(argv[5]), rpc = atoi(argv[6]);
  const __half* xh = reinterpret_cast<const __half*>(x.data());
  const __half* bh = reinterpret_cast<const __half*>(b.data());
  const int G = cf::make_row_geom(C).G;
  if (mode == "binary") return G == 1 ? run<cf::MODE_BINARY, 1>(xh, bh, N, C, rpc) : run<cf::MODE_BINARY, 2>(xh, bh, N, C, rpc);
  return G == 1 ? run<cf::MODE_INT2, 1>(xh, bh, N, C, rpc) : run<cf::MODE_INT2, 2>(xh, bh, N, C, rpc);
}
'''


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    src = open(os.path.join(emu.CSRC, "cf_sign_codecs.cu")).read()
    kern = re.search(r"(namespace cf \{.*?\n\}  // namespace cf\n)", src, flags=re.S).group(1)
    assert "k_int2_encode" in kern and "<<<" not in kern and "cf_sign_tma" not in kern
    kern = emu.strip_asm(kern).replace("extern __shared__ float smem[];", "float* smem = emu_smem;")
    d = tmp_path_factory.mktemp("codec_emu")
    return emu.build(d, emu.SHIM_HEAD + emu.common_source() + kern + emu.SLURP + RUNNER), d


def _inputs(n, c, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, c, generator=g).half()
    base = (0.97 * x.float() + 0.2 * torch.randn(n, c, generator=g)).half()
    base[0, :8] = x[0, :8]       # zero deltas (sign bit 1)
    return x, base


def _run(emulator, mode, x, base, rows_per_cta):
    exe, d = emulator
    n, c = x.shape
    (d / "x.bin").write_bytes(x.numpy().tobytes())
    (d / "b.bin").write_bytes(base.numpy().tobytes())
    r = subprocess.run([exe, mode, str(d / "x.bin"), str(d / "b.bin"), str(n), str(c), str(rows_per_cta)],
                       capture_output=True, timeout=900)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    buf, per = r.stdout, (8 if mode == "binary" else 4)
    sizes = [n * c // per, 2 * n, 2 * c, 2 * n * c, 2 * n * c]
    assert len(buf) == sum(sizes)
    parts, o = [], 0
    for s in sizes:
        parts.append(buf[o:o + s])
        o += s
    half = lambda b, shape: torch.from_numpy(np.frombuffer(b, dtype=np.int16).copy()).view(torch.half).view(shape)  # noqa: E731
    packed = np.frombuffer(parts[0], dtype=np.uint8).reshape(n, c // per)
    return packed, half(parts[1], (n, 1)), half(parts[2], (c, 1)), half(parts[3], (n, c)), half(parts[4], (n, c))


def _ulp(a, b):
    return int((a.view(torch.int16).int() - b.view(torch.int16).int()).abs().max())


def _same_bits(a, b):
    return torch.equal(a.view(torch.int16), b.view(torch.int16))


@pytest.mark.parametrize("n,c,rows_per_cta", [(70, 256, 24), (33, 1152, 40), (6, 8192, 4)])
def test_binary_kernel_source_matches_the_oracle(emulator, n, c, rows_per_cta):
    x, base = _inputs(n, c, seed=n)
    packed, u, v, new_base, recon = _run(emulator, "binary", x, base, rows_per_cta)
    o_packed, o_u, o_v, _ = oc.binary_quant(x, base, False)
    assert np.array_equal(packed, o_packed), "sign bits differ from the oracle"
    assert _ulp(u, o_u) <= 1 and _ulp(v, o_v) <= 1
    assert _same_bits(recon, new_base), "sender's error-feedback base != receiver's reconstruction"
    assert _same_bits(recon, oc.binary_dequant(o_packed, u, v, base)), "reconstruction differs given identical scales"


@pytest.mark.parametrize("n,c,rows_per_cta", [(70, 256, 24), (33, 1152, 40)])
def test_int2_kernel_source_matches_the_oracle(emulator, n, c, rows_per_cta):
    x, base = _inputs(n, c, seed=100 + n)
    packed, tok, chan, new_base, recon = _run(emulator, "int2", x, base, rows_per_cta)
    _, o_tok, o_chan, _ = oc.int2_quant(x, base, True)
    assert _ulp(tok, o_tok) <= 1 and _ulp(chan, o_chan) <= 1
    s_packed, _, _, s_nb = oc.int2_quant(x, base, True, scales=(tok, chan))
    assert np.array_equal(packed, s_packed), "INT2 codes differ given identical scales"
    assert _same_bits(new_base, s_nb), "INT2 error-feedback base differs given identical scales"
    assert _same_bits(recon, new_base), "receiver != sender"
