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

from oracle import codecs as oc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "compactfusion_b200", "csrc")

SHIM = r'''
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <pthread.h>
#include <string>
#include <thread>
#include <vector>
#include <vector_types.h>
#include <vector_functions.h>
#include "compactb200.h"

struct Idx { unsigned x = 0, y = 0, z = 0; };
static thread_local Idx threadIdx, blockIdx;
static Idx blockDim, gridDim;
static pthread_barrier_t cta_bar, warp_bar[32];
static unsigned char shfl_buf[32][32][8];
static float emu_smem[1 << 18];

#undef __global__
#undef __device__
#undef __host__
#undef __forceinline__
#undef __launch_bounds__
#undef __shared__
#undef __restrict__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __restrict__
static inline int min(int a, int b) { return a < b ? a : b; }
static inline void __syncthreads() { pthread_barrier_wait(&cta_bar); }
static inline void __threadfence() { __sync_synchronize(); }
static inline void __threadfence_system() { __sync_synchronize(); }
static inline void pdl_wait() {}
static inline void pdl_launch_dependents() {}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int o) {
  const unsigned lin = threadIdx.y * blockDim.x + threadIdx.x;
  const int lane = lin & 31, warp = lin >> 5;
  static_assert(sizeof(T) <= 8, "shuffle payload");
  memcpy(shfl_buf[warp][lane], &v, sizeof(T));
  pthread_barrier_wait(&warp_bar[warp]);
  T r;
  memcpy(&r, shfl_buf[warp][lane ^ o], sizeof(T));
  pthread_barrier_wait(&warp_bar[warp]);
  return r;
}
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }

// ---- software fp16: exact fp32 arithmetic + one round-to-nearest-even ------------------------------
struct __half { uint16_t v; };
struct __half2 { __half x, y; };
static inline float h2f(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu;
  uint32_t bits;
  if (exp == 0) { float f = (float)man * (1.0f / 16777216.0f); memcpy(&bits, &f, 4); bits |= sign; }
  else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
  else bits = sign | ((exp + 112u) << 23) | (man << 13);
  float out; memcpy(&out, &bits, 4); return out;
}
static inline uint16_t f2h(float f) {
  uint32_t x; memcpy(&x, &f, 4);
  const uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
  const uint32_t absx = x & 0x7FFFFFFFu;
  if (absx >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | (absx > 0x7F800000u ? 0x200u : 0u));
  if (absx >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);
  if (absx < 0x33000001u) return sign;
  const int32_t e = (int32_t)(absx >> 23) - 127;
  const uint32_t m = (absx & 0x7FFFFFu) | 0x800000u;
  const int shift = (e < -14) ? (13 + (-14 - e)) : 13;
  uint32_t kept = m >> shift;
  const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
  if (rem > half || (rem == half && (kept & 1u))) kept += 1u;
  if (e < -14) return (uint16_t)(sign | kept);
  uint32_t he = (uint32_t)(e + 15);
  if (kept & 0x800u) { kept >>= 1; he += 1; }
  if (he >= 31) return (uint16_t)(sign | 0x7C00u);
  return (uint16_t)(sign | (he << 10) | (kept & 0x3FFu));
}
static inline __half __float2half_rn(float f) { return __half{f2h(f)}; }
static inline float __half2float(__half h) { return h2f(h.v); }
static inline uint16_t __half_as_ushort(__half h) { return h.v; }
static inline __half2 __half2half2(__half h) { return __half2{h, h}; }
static inline __half2 __float2half2_rn(float f) { return __half2{__float2half_rn(f), __float2half_rn(f)}; }
static inline float2 __half22float2(__half2 h) { return make_float2(h2f(h.x.v), h2f(h.y.v)); }
static inline __half __hadd_rn(__half a, __half b) { return __float2half_rn(h2f(a.v) + h2f(b.v)); }
static inline __half __hneg(__half a) { return __half{(uint16_t)(a.v ^ 0x8000u)}; }
#define EMU_H2_OP(name, expr) \
  static inline __half2 name(__half2 a, __half2 b) { \
    const float ax = h2f(a.x.v), ay = h2f(a.y.v), bx = h2f(b.x.v), by = h2f(b.y.v); \
    (void)ax; (void)ay; (void)bx; (void)by; return __half2{__float2half_rn(expr(ax, bx)), __float2half_rn(expr(ay, by))}; }
#define EMU_ADD(p, q) ((p) + (q))
#define EMU_SUB(p, q) ((p) - (q))
#define EMU_MUL(p, q) ((p) * (q))
EMU_H2_OP(__hadd2_rn, EMU_ADD)
EMU_H2_OP(__hsub2_rn, EMU_SUB)
EMU_H2_OP(__hmul2_rn, EMU_MUL)
static inline __half2 __habs2(__half2 a) { return __half2{__half{(uint16_t)(a.x.v & 0x7FFFu)}, __half{(uint16_t)(a.y.v & 0x7FFFu)}}; }
static inline uint32_t __hge2_mask(__half2 a, __half2 b) {
  return (h2f(a.x.v) >= h2f(b.x.v) ? 0xFFFFu : 0u) | (h2f(a.y.v) >= h2f(b.y.v) ? 0xFFFF0000u : 0u);
}
static inline uint32_t __hgt2_mask(__half2 a, __half2 b) {
  return (h2f(a.x.v) > h2f(b.x.v) ? 0xFFFFu : 0u) | (h2f(a.y.v) > h2f(b.y.v) ? 0xFFFF0000u : 0u);
}
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s) {
  const uint64_t src = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) {
    const uint32_t sel = (s >> (4 * i)) & 0xFu;
    uint32_t byte = (uint32_t)(src >> (8 * (sel & 7u))) & 0xFFu;
    if (sel & 8u) byte = (byte & 0x80u) ? 0xFFu : 0u;   // sign-replicate mode
    r |= byte << (8 * i);
  }
  return r;
}
static inline uint4 ldg_stream(const void* p) { return *static_cast<const uint4*>(p); }
static inline void stg_stream(void* p, const uint4& v) { *static_cast<uint4*>(p) = v; }
static inline void stg_stream_pol(void* p, const uint4& v, uint64_t) { *static_cast<uint4*>(p) = v; }

template <class F> static void launch(unsigned gx, unsigned gy, unsigned bx, unsigned by, F body) {
  gridDim.x = gx; gridDim.y = gy; blockDim.x = bx; blockDim.y = by;
  const unsigned nthreads = bx * by;
  for (unsigned cy = 0; cy < gy; ++cy)
    for (unsigned cx = 0; cx < gx; ++cx) {
      pthread_barrier_init(&cta_bar, nullptr, nthreads);
      for (unsigned w = 0; w < (nthreads + 31) / 32; ++w) pthread_barrier_init(&warp_bar[w], nullptr, 32);
      std::vector<std::thread> ts;
      for (unsigned ty = 0; ty < by; ++ty)
        for (unsigned tx = 0; tx < bx; ++tx)
          ts.emplace_back([=] { threadIdx.x = tx; threadIdx.y = ty; blockIdx.x = cx; blockIdx.y = cy; body(); });
      for (auto& t : ts) t.join();
    }
}

namespace cf {
// ---- host + device helpers, verbatim from csrc/cf_common.cuh ----
COMMON_SOURCE
}  // namespace cf

// ---- kernels, verbatim from csrc/cf_sign_codecs.cu ----
KERNEL_SOURCE

static std::vector<unsigned char> slurp(const char* path) {
  FILE* f = fopen(path, "rb"); if (!f) { perror(path); exit(2); }
  fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<unsigned char> v(n); if (fread(v.data(), 1, n, f) != (size_t)n) exit(2); fclose(f); return v;
}

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


def _strip_asm(text):
    """Replace every `asm volatile(...);` / `asm(...);` statement by `;` (paren-balanced)."""
    out, i = [], 0
    while True:
        m = re.search(r"\basm\b(\s+volatile)?\s*\(", text[i:])
        if not m:
            out.append(text[i:])
            return "".join(out)
        out.append(text[i:i + m.start()])
        j, depth = i + m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(text[j], 0)
            j += 1
        i = text.index(";", j) + 1
        out.append(";")


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    common = open(os.path.join(CSRC, "cf_common.cuh")).read()
    geom = re.search(r"(struct RowGeom \{.*?\n\}\n)", common, flags=re.S).group(1)
    geom = re.search(r"(// Geometry shared by all row-streaming kernels.*?return g;\n\})", common, flags=re.S).group(1)
    dev = re.search(r"#ifdef __CUDACC__\n(.*?)#endif  // __CUDACC__", common, flags=re.S).group(1)
    # the three streaming load / store helpers are PTX: the shim provides them
    dev = re.sub(r"(// [^\n]*\n)*__device__ __forceinline__ (uint4|void) (ldg_stream|stg_stream|stg_stream_pol)\(.*?\n\}\n",
                 "", dev, flags=re.S)
    assert "asm" not in dev, "an inline-PTX helper of cf_common.cuh is not covered by the shim"
    src = open(os.path.join(CSRC, "cf_sign_codecs.cu")).read()
    kern = re.search(r"(namespace cf \{.*?\n\}  // namespace cf\n)", src, flags=re.S).group(1)
    assert "k_int2_encode" in kern and "<<<" not in kern and "cf_sign_tma" not in kern
    kern = _strip_asm(kern).replace("extern __shared__ float smem[];", "float* smem = emu_smem;")
    d = tmp_path_factory.mktemp("codec_emu")
    cpp = d / "emu.cpp"
    cpp.write_text(SHIM.replace("COMMON_SOURCE", geom + "\n" + dev).replace("KERNEL_SOURCE", kern))
    exe = str(d / "emu")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-pthread", "-w", "-I", "/usr/local/cuda/include", "-I",
                        os.path.join(ROOT, "include"), str(cpp), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return exe, d


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
