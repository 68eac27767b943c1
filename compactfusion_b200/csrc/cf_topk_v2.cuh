// SPARSE 1:m kernels, second version (included by cf_topk.cu inside namespace cf).
//
// k_topk_compress / k_topk_decompress give every thread 32 CONSECUTIVE elements: a warp's 128-bit load touches 32
// separate 64-byte segments and uses a quarter of every sector it pulls (0.50-0.67 of the HBM roofline in round
// 1's sweep, LDG-bound).  Here a thread owns ONE 16-byte group of 8 elements, so every load / store instruction of
// a warp covers 512 contiguous bytes, and four such groups per thread are in flight (grid-stride x 4).  A block
// of m <= 8 elements lives inside one thread; for m = 16 the two lanes of a pair exchange their local maxima
// (the even lane holds the lower indices and wins ties).  Index bytes hold two blocks (first block in the high
// nibble, compress_topk.py:100): for m = 8 / 16 the lanes sharing a byte combine their nibbles with a shuffle.
// Same values, indices and tie rule as the first version, bit for bit.
#pragma once

template <int M>
__global__ void __launch_bounds__(256) k_topk_compress_v2(const __half* __restrict__ x, const __half* __restrict__ base,
                                                          __half* __restrict__ new_base, __half* __restrict__ val,
                                                          uint8_t* __restrict__ idx, int64_t ngroups) {
  constexpr int NB = (M <= 8) ? 8 / M : 1;   // blocks (or half-blocks for M = 16) per thread
  constexpr int kU = 4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int lane = threadIdx.x & 31;
  // ngroups is a multiple of 128 (numel % 1024 == 0) and the stride a multiple of 32: a warp is either entirely
  // inside or entirely outside, so the shuffles below always see full warps
  for (int64_t i0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i0 < ngroups; i0 += kU * stride) {
    uint4 xv[kU], bv[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t i = i0 + u * stride;
      xv[u] = bv[u] = make_uint4(0, 0, 0, 0);
      if (i < ngroups) {
        xv[u] = ldg_stream(x + i * 8);
        if (base != nullptr) bv[u] = ldg_stream(base + i * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= ngroups) continue;   // warp-uniform
      const H8 d8 = h8_sub(as_h8(xv[u]), as_h8(bv[u]));
      __align__(16) __half v[8];
      *reinterpret_cast<uint4*>(v) = as_u4(d8);
      uint32_t sel[NB];
      __half pick[NB];
      float best[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        constexpr int W = (M <= 8) ? M : 8;
        int bsel = 0;
        float bm = fabsf(__half2float(v[k * W]));
#pragma unroll
        for (int e = 1; e < W; ++e) {
          const float a = fabsf(__half2float(v[k * W + e]));
          if (a > bm) { bm = a; bsel = e; }   // strict: lowest index wins ties
        }
        sel[k] = bsel;
        best[k] = bm;
        __half p = v[k * W];
#pragma unroll
        for (int e = 1; e < W; ++e) if (e == bsel) p = v[k * W + e];
        pick[k] = p;
      }
      bool mine = true;   // M = 16: does the block's maximum sit in this thread's half?
      if (M == 16) {
        const float ob = __shfl_xor_sync(0xffffffffu, best[0], 1);
        const uint32_t os = __shfl_xor_sync(0xffffffffu, sel[0], 1);
        const uint32_t op = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(__half_as_ushort(pick[0])), 1);
        const bool even = (lane & 1) == 0;
        mine = even ? !(ob > best[0]) : (best[0] > ob);   // ties: the even lane (lower indices)
        const uint32_t gsel = mine ? (sel[0] + (even ? 0u : 8u)) : (os + (even ? 8u : 0u));
        if (!mine) pick[0] = __ushort_as_half(static_cast<unsigned short>(op));
        sel[0] = gsel;   // 0..15 within the 16-element block, known to both lanes
      }
      // ---- values ----
      if (M == 2) {
        __align__(8) __half o4[4] = {pick[0], pick[1], pick[2], pick[3]};
        *reinterpret_cast<uint2*>(val + i * 4) = *reinterpret_cast<uint2*>(o4);
      } else if (M == 4) {
        *reinterpret_cast<__half2*>(val + i * 2) = __halves2half2(pick[0], pick[1]);
      } else if (M == 8) {
        val[i] = pick[0];
      } else if ((lane & 1) == 0) {
        val[i >> 1] = pick[0];
      }
      // ---- indices: one byte per block pair, first block in the high nibble ----
      if (M == 2) {
        const uint32_t w = ((sel[0] << 4) | sel[1]) | (((sel[2] << 4) | sel[3]) << 8);
        *reinterpret_cast<uint16_t*>(idx + i * 2) = static_cast<uint16_t>(w);
      } else if (M == 4) {
        idx[i] = static_cast<uint8_t>((sel[0] << 4) | sel[1]);
      } else if (M == 8) {
        const uint32_t other = __shfl_xor_sync(0xffffffffu, sel[0], 1);
        if ((lane & 1) == 0) idx[i >> 1] = static_cast<uint8_t>((sel[0] << 4) | other);
      } else {
        const uint32_t other = __shfl_xor_sync(0xffffffffu, sel[0], 2);
        if ((lane & 3) == 0) idx[i >> 2] = static_cast<uint8_t>((sel[0] << 4) | other);
      }
      // ---- error feedback: new_base = base + sparsified delta ----
      if (new_base != nullptr) {
        __align__(16) __half b[8];
        *reinterpret_cast<uint4*>(b) = bv[u];
        const __half zero = __float2half_rn(0.f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          bool hit;
          if (M <= 8) hit = static_cast<uint32_t>(e % M) == sel[e / M];
          else hit = mine && static_cast<uint32_t>(e + ((lane & 1) ? 8 : 0)) == sel[0];
          const __half add = hit ? v[e] : zero;
          b[e] = (base != nullptr) ? __hadd_rn(b[e], add) : add;
        }
        stg_stream(new_base + i * 8, *reinterpret_cast<uint4*>(b));
      }
    }
  }
}

template <int M>
__global__ void __launch_bounds__(256) k_topk_decompress_v2(const __half* __restrict__ val, const uint8_t* __restrict__ idx,
                                                            const __half* __restrict__ base, __half* __restrict__ recon,
                                                            int64_t ngroups) {
  constexpr int kU = 4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i0 < ngroups; i0 += kU * stride) {
    uint4 bv[kU];
    uint32_t vals[kU][2], bytes[kU];   // up to 4 fp16 values and 2 index bytes per group
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t i = i0 + u * stride;
      bv[u] = make_uint4(0, 0, 0, 0);
      vals[u][0] = vals[u][1] = bytes[u] = 0;
      if (i >= ngroups) continue;
      if (base != nullptr) bv[u] = ldg_stream(base + i * 8);
      if (M == 2) {
        const uint2 w = *reinterpret_cast<const uint2*>(val + i * 4);
        vals[u][0] = w.x; vals[u][1] = w.y;
        bytes[u] = *reinterpret_cast<const uint16_t*>(idx + i * 2);
      } else if (M == 4) {
        vals[u][0] = *reinterpret_cast<const uint32_t*>(val + i * 2);
        bytes[u] = idx[i];
      } else if (M == 8) {
        vals[u][0] = __half_as_ushort(val[i]);
        bytes[u] = idx[i >> 1];
      } else {
        vals[u][0] = __half_as_ushort(val[i >> 1]);
        bytes[u] = idx[i >> 2];
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= ngroups) continue;
      __align__(16) __half b[8];
      *reinterpret_cast<uint4*>(b) = bv[u];
      const __half zero = __float2half_rn(0.f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        uint32_t sel, pv;
        bool hit;
        if (M == 2) {
          const int k = e / 2;   // block k of the group: byte k / 2, high nibble first
          sel = (bytes[u] >> (8 * (k >> 1) + ((k & 1) ? 0 : 4))) & 0xFu;
          pv = (vals[u][k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
          hit = static_cast<uint32_t>(e % 2) == sel;
        } else if (M == 4) {
          const int k = e / 4;
          sel = (bytes[u] >> ((k & 1) ? 0 : 4)) & 0xFu;
          pv = (vals[u][0] >> (16 * k)) & 0xFFFFu;
          hit = static_cast<uint32_t>(e % 4) == sel;
        } else if (M == 8) {
          sel = (bytes[u] >> ((i & 1) ? 0 : 4)) & 0xFu;
          pv = vals[u][0];
          hit = static_cast<uint32_t>(e) == sel;
        } else {
          sel = (bytes[u] >> (((i >> 1) & 1) ? 0 : 4)) & 0xFu;
          pv = vals[u][0];
          hit = static_cast<uint32_t>(e + ((i & 1) ? 8 : 0)) == sel;
        }
        const __half add = hit ? __ushort_as_half(static_cast<unsigned short>(pv)) : zero;
        b[e] = (base != nullptr) ? __hadd_rn(b[e], add) : add;
      }
      stg_stream(recon + i * 8, *reinterpret_cast<uint4*>(b));
    }
  }
}
