"""Execute the SOURCE of the tensor-core low-rank kernels (csrc/cf_lowrank_mma.cuh: k_lr_gemm with TF32
mma.sync + ldmatrix + cp.async, the CholeskyQR2 kernels, k_lr_reconstruct_mma with fp16 mma.sync) on the CPU.

On top of tests/cuda_emulation.py, the warp-collective PTX wrappers get host implementations with the PTX
fragment layouts: ldmatrix(.trans).x4 and mma.sync.m16n8k8.tf32 / m16n8k16.f16 exchange the lanes' registers
through a per-warp buffer between two warp barriers and every lane computes its own accumulator elements;
cp.async is a synchronous 16-byte copy (zero fill when predicated off); cvt.rna.tf32 is done on the bit
pattern.  The projector's launch sequence (lr_mma_project, cf_lowrank.cu) is restated in the runner.

Checked against the oracle started from the same Q0 at the GPU tests' bars (only U V is comparable: QR sign /
basis conventions are arbitrary), U orthonormal, and the fused reconstruct against base + fp16(U V).
"""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import cuda_emulation as emu
from oracle import codecs as oc

MMA_SHIM = r'''
// ---- warp-collective PTX wrappers of cf_lowrank_mma.cuh, host implementations (PTX fragment layouts) ----
static uint32_t emu_frag[32][32][8];
static unsigned char emu_rows[32][32][16];
static inline void emu_lw(int& lane, int& warp) { const unsigned lin = threadIdx.y * blockDim.x + threadIdx.x; lane = lin & 31; warp = lin >> 5; }
static inline float emu_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __uint_as_float(uint32_t u) { return emu_u2f(u); }
static inline void cp_async16(void* dst, const void* src, bool valid) { if (valid) memcpy(dst, src, 16); else memset(dst, 0, 16); }
static inline void cp_async_commit() {}
template <int N> static inline void cp_async_wait() {}
static inline void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  int L, W; emu_lw(L, W);
  memcpy(emu_rows[W][L], p, 16);
  pthread_barrier_wait(&warp_bar[W]);
  for (int j = 0; j < 4; ++j) memcpy(&r[j], emu_rows[W][j * 8 + (L >> 2)] + 4 * (L & 3), 4);
  pthread_barrier_wait(&warp_bar[W]);
}
static inline void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  int L, W; emu_lw(L, W);
  memcpy(emu_rows[W][L], p, 16);
  pthread_barrier_wait(&warp_bar[W]);
  for (int j = 0; j < 4; ++j) {
    uint16_t lo, hi;
    memcpy(&lo, emu_rows[W][j * 8 + 2 * (L & 3)] + 2 * (L >> 2), 2);
    memcpy(&hi, emu_rows[W][j * 8 + 2 * (L & 3) + 1] + 2 * (L >> 2), 2);
    r[j] = (uint32_t)lo | ((uint32_t)hi << 16);
  }
  pthread_barrier_wait(&warp_bar[W]);
}
static inline void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  int L, W; emu_lw(L, W);
  for (int i = 0; i < 4; ++i) emu_frag[W][L][i] = a[i];
  emu_frag[W][L][4] = b0; emu_frag[W][L][5] = b1;
  pthread_barrier_wait(&warp_bar[W]);
  const int g = L >> 2, t = L & 3;
  const int rows[4] = {g, g, g + 8, g + 8}, cols[4] = {2 * t, 2 * t + 1, 2 * t, 2 * t + 1};
  for (int e = 0; e < 4; ++e) {
    float s = d[e];
    for (int k = 0; k < 8; ++k) {   // A: a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  B: b0 (k=t, n=g) b1 (k=t+4, n=g)
      const float av = emu_u2f(emu_frag[W][(rows[e] & 7) * 4 + (k & 3)][(rows[e] >= 8 ? 1 : 0) + (k >= 4 ? 2 : 0)]);
      const float bv = emu_u2f(emu_frag[W][cols[e] * 4 + (k & 3)][4 + (k >= 4 ? 1 : 0)]);
      s += av * bv;
    }
    d[e] = s;
  }
  pthread_barrier_wait(&warp_bar[W]);
}
static inline void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  int L, W; emu_lw(L, W);
  for (int i = 0; i < 4; ++i) emu_frag[W][L][i] = a[i];
  emu_frag[W][L][4] = b0; emu_frag[W][L][5] = b1;
  pthread_barrier_wait(&warp_bar[W]);
  const int g = L >> 2, t = L & 3;
  const int rows[4] = {g, g, g + 8, g + 8}, cols[4] = {2 * t, 2 * t + 1, 2 * t, 2 * t + 1};
  for (int e = 0; e < 4; ++e) {
    float s = d[e];
    for (int k = 0; k < 16; ++k) {  // A: a0 (g, 2t..) a1 (g+8, 2t..) a2 (g, 2t+8..) a3 (g+8, 2t+8..);  B: b0 (k=2t.., n=g) b1 (k=2t+8.., n=g)
      const uint32_t aw = emu_frag[W][(rows[e] & 7) * 4 + ((k & 7) >> 1)][(rows[e] >= 8 ? 1 : 0) + (k >= 8 ? 2 : 0)];
      const uint32_t bw = emu_frag[W][cols[e] * 4 + ((k & 7) >> 1)][4 + (k >= 8 ? 1 : 0)];
      s += h2f((uint16_t)((k & 1) ? (aw >> 16) : (aw & 0xFFFFu))) * h2f((uint16_t)((k & 1) ? (bw >> 16) : (bw & 0xFFFFu)));
    }
    d[e] = s;
  }
  pthread_barrier_wait(&warp_bar[W]);
}
static inline uint32_t emu_rna_tf32(float v) {   // cvt.rna.tf32.f32: nearest, ties away, 10 mantissa bits
  uint32_t u; memcpy(&u, &v, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return u;
  return (u + 0x1000u) & 0xFFFFE000u;
}
static inline float2 split_tf32(float v) {
  const uint32_t hi = emu_rna_tf32(v);
  const float rest = v - emu_u2f(hi);
  return make_float2(emu_u2f(hi), emu_u2f(emu_rna_tf32(rest)));
}
static inline unsigned atomicMax(unsigned* p, unsigned v) {
  unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
static inline double rsqrt(double d) { return 1.0 / sqrt(d); }
static inline float rsqrtf(float d) { return 1.0f / sqrtf(d); }
'''

RUNNER = r'''
using namespace cf;
template <int RP, bool HB = false>
static int project(const __half* x, const __half* base, const float* q0, int n, int c, int r, int iters,
                   std::vector<__half>& U, std::vector<__half>& V) {
  // HB: the bases Q and the orthonormal U as fp16 planes, k_lr_gemm<.., 64, 2, true> (lr_mma_project's default for RP >= 16)
  const int aq_splits = 2, aty_splits = 3;
  const int aq_kper = ((c + aq_splits - 1) / aq_splits + kLrBK - 1) / kLrBK * kLrBK;
  const int aty_kper = ((n + aty_splits - 1) / aty_splits + kLrBK - 1) / kLrBK * kLrBK;
  const int aq_s = (c + aq_kper - 1) / aq_kper, aty_s = (n + aty_kper - 1) / aty_kper;
  const size_t maxm = n > c ? n : c;
  std::vector<float2> Q2((size_t)c * RP), Y2((size_t)n * RP);
  std::vector<float> Xsum(maxm * RP), part((size_t)(aq_s > aty_s ? aq_s : aty_s) * maxm * RP), rfac(RP * RP), rdinv(RP);
  std::vector<double> gpart(16 * (size_t)r * r);
  unsigned ticket = 0;
  U.assign((size_t)n * r, __half{0}); V.assign((size_t)r * c, __half{0});
  launch(2, 1, 256, 1, [&] { k_lr_pad_split(q0, Q2.data(), c, r, RP, HB ? 1 : 0); });
  unsigned amax_slots[64] = {0};
  auto gemm_AQ = [&](unsigned* amax = nullptr) {
    launch((n + kLrBM - 1) / kLrBM, aq_s, kLrThreads, 1, [&] {
      if (HB) k_lr_gemm<RP, false, 64, 2, HB>(x, base, Q2.data(), part.data(), n, c, aq_kper, amax);
      else k_lr_gemm<RP, false>(x, base, Q2.data(), part.data(), n, c, aq_kper, nullptr);
    });
  };
  auto gemm_AtY = [&](bool planes = false) {
    launch((c + kLrBM - 1) / kLrBM, aty_s, kLrThreads, 1, [&] {
      if (planes) k_lr_gemm<RP, true, 64, 2, HB>(x, base, Y2.data(), part.data(), n, c, aty_kper, nullptr);
      else k_lr_gemm<RP, true>(x, base, Y2.data(), part.data(), n, c, aty_kper, nullptr);
    });
  };
  auto orth = [&](int S, int M, float2* out2, __half* out16) {
    launch(2, 1, 256, 1, [&] { k_lr_sum_split(part.data(), S, (size_t)M * RP, nullptr, Xsum.data(), (size_t)M * RP, nullptr); });
    const int ctas = 3, rows = ((M + ctas - 1) / ctas + 31) / 32 * 32;
    GramParams g{};
    g.xpart = Xsum.data(); g.S = 1; g.part_stride = (size_t)M * RP; g.X = Xsum.data(); g.M = M; g.r = r;
    g.rows_per_cta = rows; g.gpart = gpart.data(); g.ticket = &ticket; g.r_out = rfac.data(); g.rdinv_out = rdinv.data();
    const int nct = (M + rows - 1) / rows;
    for (int pass = 0; pass < 2; ++pass) {   // CholeskyQR2
      launch(nct, 1, 256, 1, [&] { k_lr_gram_chol<RP>(g); });
      const bool last = pass == 1;
      launch((M + 127) / 128, 1, 128, 1, [&] { k_lr_solve_out<RP>(Xsum.data(), rfac.data(), rdinv.data(), M, r, last ? out2 : nullptr, last ? out16 : nullptr, nullptr, HB ? 1 : 0); });
    }
  };
  for (int it = 0; it < iters; ++it) {
    unsigned* amax = HB ? &amax_slots[it] : nullptr;   // HB: Y scaled by a power of two into fp16 planes
    gemm_AQ(amax);
    launch(2, 1, 256, 1, [&] { k_lr_sum_split(part.data(), aq_s, (size_t)n * RP, Y2.data(), nullptr, (size_t)n * RP, amax); });
    gemm_AtY(HB);
    orth(aty_s, c, Q2.data(), nullptr);
  }
  gemm_AQ();
  orth(aq_s, n, Y2.data(), U.data());
  gemm_AtY(HB);
  launch(2, 1, 256, 1, [&] { k_lr_store_v_sum(part.data(), aty_s, (size_t)c * RP, V.data(), c, RP, r); });
  return ticket == 0 ? 0 : 9;
}

int main(int argc, char** argv) {  // x.bin base.bin q0.bin N C r iters [hb]
  auto x = slurp(argv[1]), b = slurp(argv[2]), q = slurp(argv[3]);
  const int N = atoi(argv[4]), C = atoi(argv[5]), r = atoi(argv[6]), iters = atoi(argv[7]);
  const __half* xh = reinterpret_cast<const __half*>(x.data());
  const __half* bh = reinterpret_cast<const __half*>(b.data());
  const float* q0 = reinterpret_cast<const float*>(q.data());
  std::vector<__half> U, V, recon((size_t)N * C);
  const bool hb = argc > 8 && atoi(argv[8]) != 0 && r > 8;
  int rc;
  if (hb)
    rc = r <= 16 ? project<16, true>(xh, bh, q0, N, C, r, iters, U, V) : project<32, true>(xh, bh, q0, N, C, r, iters, U, V);
  else
    rc = r <= 8 ? project<8>(xh, bh, q0, N, C, r, iters, U, V) : (r <= 16 ? project<16>(xh, bh, q0, N, C, r, iters, U, V)
                                                                          : project<32>(xh, bh, q0, N, C, r, iters, U, V));
  if (rc) return rc;
  const int KS = (r + 15) / 16;
  auto rec = [&](auto ks) {
    constexpr int K = decltype(ks)::value;
    launch((C + 255) / 256, (N + 63) / 64, 128, 1, [&] { k_lr_reconstruct_mma<K>(U.data(), V.data(), bh, recon.data(), N, C, r); });
  };
  if (KS == 1) rec(std::integral_constant<int, 1>{}); else rec(std::integral_constant<int, 2>{});
  fwrite(U.data(), 2, U.size(), stdout);
  fwrite(V.data(), 2, V.size(), stdout);
  fwrite(recon.data(), 2, recon.size(), stdout);
  return 0;
}
'''

_WRAPPERS = ["cp_async16", "cp_async_commit", "cp_async_wait", "ldmatrix_x4", "ldmatrix_x4_trans", "mma_tf32",
             "split_tf32", "mma_f16"]


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    hdr = open(os.path.join(emu.CSRC, "cf_lowrank_mma.cuh")).read()
    hdr = hdr.replace("#pragma once", "").replace('#include "cf_common.cuh"', "")
    for name in _WRAPPERS:   # the PTX wrappers are replaced as whole functions
        pat = r"(template <int N>\n)?__device__ __forceinline__ [\w ]+? " + name + r"\(.*?\n\}\n" if name not in (
            "cp_async_commit", "cp_async_wait") else r"(template <int N>\n)?__device__ __forceinline__ void " + name + r"\(\)[^\n]*\n"
        hdr, n = re.subn(pat, "", hdr, count=1, flags=re.S)
        assert n == 1, f"wrapper {name} not found"
    assert "asm" not in hdr, "an inline-PTX wrapper of cf_lowrank_mma.cuh is not covered by the shim"
    hdr = hdr.replace("extern __shared__ __align__(128) unsigned char lr_smem_raw[];",
                      "unsigned char* lr_smem_raw = reinterpret_cast<unsigned char*>(emu_smem);")
    d = tmp_path_factory.mktemp("lr_emu")
    text = ("#define __align__(n) alignas(n)\n" + emu.SHIM_HEAD + "#include <type_traits>\n"
            "static inline void __syncwarp() { const unsigned lin = threadIdx.y * blockDim.x + threadIdx.x; pthread_barrier_wait(&warp_bar[lin >> 5]); }\n"
            + MMA_SHIM + emu.common_source() + hdr + emu.SLURP + RUNNER)
    return emu.build(d, text), d


@pytest.mark.parametrize("n,c,r,iters,amp", [(160, 256, 8, 2, 1.0), (200, 320, 12, 2, 1.0), (130, 192, 20, 1, 1.0),
                                             (144, 256, 16, 2, 3000.0), (144, 256, 16, 2, 2e-3)])
def test_lowrank_kernel_source_tracks_the_oracle(emulator, n, c, r, iters, amp):
    """amp != 1: activations near the top / bottom of the fp16 range -- Y = A Q leaves it (|Y| ~ 1e5) or sits far
    below 1, which is what the power-of-two scaling of the fp16 planes is for."""
    exe, d = emulator
    g = torch.Generator().manual_seed(n + r)
    low = torch.randn(n, r, generator=g) @ torch.randn(r, c, generator=g) / r ** 0.5
    x = (amp * (low + 0.05 * torch.randn(n, c, generator=g))).half()
    base = (amp * 0.1 * torch.randn(n, c, generator=g)).half()
    assert bool(torch.isfinite(x).all())
    q0, _ = torch.linalg.qr(torch.randn(c, r, generator=g))
    q0 = q0.contiguous().float()
    (d / "x.bin").write_bytes(x.numpy().tobytes())
    (d / "b.bin").write_bytes(base.numpy().tobytes())
    (d / "q.bin").write_bytes(q0.numpy().tobytes())
    half = lambda b, shape: torch.from_numpy(np.frombuffer(b, dtype=np.int16).copy()).view(torch.half).view(shape)  # noqa: E731

    def run(hb):
        res = subprocess.run([exe, str(d / "x.bin"), str(d / "b.bin"), str(d / "q.bin"), str(n), str(c), str(r), str(iters),
                              "1" if hb else "0"], capture_output=True, timeout=1500)
        assert res.returncode == 0, res.stderr.decode()[-2000:]
        buf = res.stdout
        return (half(buf[:2 * n * r], (n, r)), half(buf[2 * n * r:2 * n * r + 2 * r * c], (r, c)),
                half(buf[2 * n * r + 2 * r * c:], (n, c)))
    u, v, recon = run(False)
    if r > 8:
        # the fp16-plane products (bases and the orthonormal U as hi + lo * 2^-11, k_lr_gemm<.., true>) carry the
        # same ~21 bits of the skinny operand as the TF32 pairs: U V agrees to the fp16 rounding of the outputs
        uh, vh, _ = run(True)
        a, b = u.float() @ v.float(), uh.float() @ vh.float()
        assert float((a - b).norm() / a.norm()) < 5e-4
        assert torch.allclose(uh.float().t() @ uh.float(), torch.eye(r), atol=5e-3)
    delta = (x - base)
    ou, ov, _ = oc.subspace_iter(delta, r, iters, init_q=q0)
    got, want = u.float() @ v.float(), ou.float() @ ov.float()
    assert float((got - want).norm() / want.norm()) < 1e-2                       # tests/test_gpu_lowrank.py's bar
    assert torch.allclose(u.float().t() @ u.float(), torch.eye(r), atol=5e-3)    # U orthonormal
    assert float((got - delta.float()).norm() / delta.float().norm()) < 0.35     # it does approximate the residual
    ref = (base.float() + (u.float() @ v.float()).half().float()).half()
    assert float((recon.float() - ref.float()).norm() / ref.float().norm()) < 1e-3
    assert float((recon.float() - ref.float()).abs().max()) <= 2e-2 * max(amp, 1.0)   # a few fp16 ulp of the product
