// SPARSE 1:m codec ("top-k" in the reference: per-m-block argmax |v|) for sm_100a.
//
// Reference: xfuser/compact/compress_topk.py (topk_compress :11-104, topk_decompress
// :108-163, topk_sparsify :165-219, sim_topk :221-236).  The reference launches one Triton
// program per 1024-element row; here every thread owns 32 consecutive elements (four 128-bit
// loads), i.e. 32/m blocks and 16/m index bytes, so the grid scales with the tensor and all
// accesses are coalesced.  The residual subtract (v = x - base) and the error-feedback update
// (new_base = base + sparsified v) are fused in.  Ties: lowest index wins (Triton argmax,
// SURVEY.md App-B.7).
#include <stdlib.h>

#include "cf_common.cuh"

namespace cf {

template <int M>
__global__ void __launch_bounds__(256) k_topk_compress(const __half* __restrict__ x, const __half* __restrict__ base,
                                                       __half* __restrict__ new_base, __half* __restrict__ val,
                                                       uint8_t* __restrict__ idx, int64_t nthreads_total) {
  constexpr int NB = 32 / M;  // blocks per thread
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nthreads_total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const size_t off = static_cast<size_t>(i) * 32;
    __align__(16) __half v[32];
    __align__(16) __half b[32];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 xv = ldg_stream(x + off + 8 * q);
      uint4 bv = make_uint4(0, 0, 0, 0);
      if (base != nullptr) bv = ldg_stream(base + off + 8 * q);
      const H8 d = h8_sub(as_h8(xv), as_h8(bv));
      *reinterpret_cast<uint4*>(v + 8 * q) = as_u4(d);
      *reinterpret_cast<uint4*>(b + 8 * q) = bv;
    }
    __align__(16) __half outv[NB];
    uint32_t sel[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      int best = 0;
      float bm = fabsf(__half2float(v[k * M]));
#pragma unroll
      for (int e = 1; e < M; ++e) {
        const float a = fabsf(__half2float(v[k * M + e]));
        if (a > bm) { bm = a; best = e; }  // strict: lowest index wins ties
      }
      sel[k] = best;
      __half pick = v[k * M];
#pragma unroll
      for (int e = 1; e < M; ++e) if (e == best) pick = v[k * M + e];
      outv[k] = pick;
    }
    // values: NB halfs per thread, contiguous
    if (NB >= 8) {
#pragma unroll
      for (int q = 0; q < NB / 8; ++q)
        *reinterpret_cast<uint4*>(val + static_cast<size_t>(i) * NB + 8 * q) = *reinterpret_cast<uint4*>(outv + 8 * q);
    } else if (NB == 4) {
      *reinterpret_cast<uint2*>(val + static_cast<size_t>(i) * NB) = *reinterpret_cast<uint2*>(outv);
    } else {
      *reinterpret_cast<uint32_t*>(val + static_cast<size_t>(i) * NB) = *reinterpret_cast<uint32_t*>(outv);
    }
    // indices: one byte per block pair, first block in the high nibble (compress_topk.py:100)
#pragma unroll
    for (int k = 0; k < NB / 2; ++k)
      idx[static_cast<size_t>(i) * (NB / 2) + k] = static_cast<uint8_t>((sel[2 * k] << 4) | sel[2 * k + 1]);
    if (new_base != nullptr) {
      const __half zero = __float2half_rn(0.f);
#pragma unroll
      for (int k = 0; k < NB; ++k)
#pragma unroll
        for (int e = 0; e < M; ++e) {
          const __half add = (e == static_cast<int>(sel[k])) ? v[k * M + e] : zero;
          b[k * M + e] = (base != nullptr) ? __hadd_rn(b[k * M + e], add) : add;
        }
#pragma unroll
      for (int q = 0; q < 4; ++q) stg_stream(new_base + off + 8 * q, *reinterpret_cast<uint4*>(b + 8 * q));
    }
  }
}

template <int M>
__global__ void __launch_bounds__(256) k_topk_decompress(const __half* __restrict__ val,
                                                         const uint8_t* __restrict__ idx,
                                                         const __half* __restrict__ base, __half* __restrict__ recon,
                                                         int64_t nthreads_total) {
  constexpr int NB = 32 / M;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nthreads_total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const size_t off = static_cast<size_t>(i) * 32;
    __align__(16) __half b[32];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 bv = make_uint4(0, 0, 0, 0);
      if (base != nullptr) bv = ldg_stream(base + off + 8 * q);
      *reinterpret_cast<uint4*>(b + 8 * q) = bv;
    }
    const __half zero = __float2half_rn(0.f);
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const uint8_t byte = idx[static_cast<size_t>(i) * (NB / 2) + (k >> 1)];
      const int sel = (k & 1) ? (byte & 0xF) : (byte >> 4);
      const __half pv = val[static_cast<size_t>(i) * NB + k];
#pragma unroll
      for (int e = 0; e < M; ++e) {
        const __half add = (e == sel) ? pv : zero;
        b[k * M + e] = (base != nullptr) ? __hadd_rn(b[k * M + e], add) : add;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) stg_stream(recon + off + 8 * q, *reinterpret_cast<uint4*>(b + 8 * q));
  }
}

static int topk_check(int64_t numel, int m) {
  CF_CHECK_ARG(m == 2 || m == 4 || m == 8 || m == 16, "sparse ratio m=%d must be 2, 4, 8 or 16", m);
  CF_CHECK_ARG(numel >= 1024 && numel % 1024 == 0, "numel=%lld must be a positive multiple of 1024", (long long)numel);
  return CF_OK;
}
static int topk_grid(int64_t nthreads) {
  int64_t blocks = (nthreads + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks);
}
// second version (cf_topk_v2.cuh): one 16-byte group per thread, four in flight; CF_LEGACY_KERNELS=1 keeps the first
static bool topk_v2() {
  const char* e = getenv("CF_LEGACY_KERNELS");
  return !(e && e[0] == '1');
}
static int topk_grid_v2(int64_t ngroups) {
  int64_t blocks = (ngroups + 4 * 256 - 1) / (4 * 256);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}
#include "cf_topk_v2.cuh"

}  // namespace cf

extern "C" {
int cf_topk_compress(const void* x, const void* base, void* new_base, void* val, void* idx, int64_t numel, int m,
                     cf_stream_t stream) {
  if (int rc = cf::topk_check(numel, m)) return rc;
  CF_CHECK_ARG(x && val && idx, "null pointer");
  CF_CHECK_ARG(cf::aligned16(x) && (!base || cf::aligned16(base)) && (!new_base || cf::aligned16(new_base)) &&
                   cf::aligned16(val),
               "x/base/new_base/val must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t nt = numel / 32;
  const int grid = cf::topk_grid(nt);
  const __half* xh = static_cast<const __half*>(x);
  const __half* bh = static_cast<const __half*>(base);
  __half* nb = static_cast<__half*>(new_base);
  __half* vh = static_cast<__half*>(val);
  uint8_t* ih = static_cast<uint8_t*>(idx);
  if (cf::topk_v2() && (reinterpret_cast<uintptr_t>(idx) & 1u) == 0) {
    const int64_t ng = numel / 8;
    const int g2 = cf::topk_grid_v2(ng);
    switch (m) {
      case 2: cf::k_topk_compress_v2<2><<<g2, 256, 0, st>>>(xh, bh, nb, vh, ih, ng); break;
      case 4: cf::k_topk_compress_v2<4><<<g2, 256, 0, st>>>(xh, bh, nb, vh, ih, ng); break;
      case 8: cf::k_topk_compress_v2<8><<<g2, 256, 0, st>>>(xh, bh, nb, vh, ih, ng); break;
      case 16: cf::k_topk_compress_v2<16><<<g2, 256, 0, st>>>(xh, bh, nb, vh, ih, ng); break;
    }
    CF_CHECK_LAUNCH();
    return CF_OK;
  }
  switch (m) {
    case 2: cf::k_topk_compress<2><<<grid, 256, 0, st>>>(xh, bh, nb, vh, ih, nt); break;
    case 4: cf::k_topk_compress<4><<<grid, 256, 0, st>>>(xh, bh, nb, vh, ih, nt); break;
    case 8: cf::k_topk_compress<8><<<grid, 256, 0, st>>>(xh, bh, nb, vh, ih, nt); break;
    case 16: cf::k_topk_compress<16><<<grid, 256, 0, st>>>(xh, bh, nb, vh, ih, nt); break;
  }
  CF_CHECK_LAUNCH();
  return CF_OK;
}

int cf_topk_decompress(const void* val, const void* idx, const void* base, void* recon, int64_t numel, int m,
                       cf_stream_t stream) {
  if (int rc = cf::topk_check(numel, m)) return rc;
  CF_CHECK_ARG(val && idx && recon, "null pointer");
  CF_CHECK_ARG(cf::aligned16(recon) && (!base || cf::aligned16(base)) && cf::aligned2(val),
               "base/recon must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t nt = numel / 32;
  const int grid = cf::topk_grid(nt);
  const __half* vh = static_cast<const __half*>(val);
  const uint8_t* ih = static_cast<const uint8_t*>(idx);
  const __half* bh = static_cast<const __half*>(base);
  __half* rh = static_cast<__half*>(recon);
  if (cf::topk_v2() && cf::aligned16(val) && (reinterpret_cast<uintptr_t>(idx) & 1u) == 0) {
    const int64_t ng = numel / 8;
    const int g2 = cf::topk_grid_v2(ng);
    switch (m) {
      case 2: cf::k_topk_decompress_v2<2><<<g2, 256, 0, st>>>(vh, ih, bh, rh, ng); break;
      case 4: cf::k_topk_decompress_v2<4><<<g2, 256, 0, st>>>(vh, ih, bh, rh, ng); break;
      case 8: cf::k_topk_decompress_v2<8><<<g2, 256, 0, st>>>(vh, ih, bh, rh, ng); break;
      case 16: cf::k_topk_decompress_v2<16><<<g2, 256, 0, st>>>(vh, ih, bh, rh, ng); break;
    }
    CF_CHECK_LAUNCH();
    return CF_OK;
  }
  switch (m) {
    case 2: cf::k_topk_decompress<2><<<grid, 256, 0, st>>>(vh, ih, bh, rh, nt); break;
    case 4: cf::k_topk_decompress<4><<<grid, 256, 0, st>>>(vh, ih, bh, rh, nt); break;
    case 8: cf::k_topk_decompress<8><<<grid, 256, 0, st>>>(vh, ih, bh, rh, nt); break;
    case 16: cf::k_topk_decompress<16><<<grid, 256, 0, st>>>(vh, ih, bh, rh, nt); break;
  }
  CF_CHECK_LAUNCH();
  return CF_OK;
}
}
