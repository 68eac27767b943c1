// Bulk-async (TMA 1-D) pipelined versions of the INT4 / INT2-minmax streaming kernels (included by
// cf_minmax_codecs.cu after the arithmetic helpers int4_codes2 / int4_values2 / finalize_column).
//
// Round 1's register-staged kernels reached 0.29 of the HBM roofline for `int4.compress_ef` at 4608 x 3072
// (ncu, profiles/r2_ncu_int4_topk_before.md: 12.7 us statistics + 11.4 us finalize + 25.6 us encode): one
// 16-byte load per thread in flight, a finalize whose every thread walked 37 dependent 2-byte loads, and an
// encode pass that re-read x and base from HBM.  Same skeleton as the BINARY / INT2 kernels (cf_sign_tma.cuh)
// here: one producer lane streams row tiles (a tile of R rows of a row-major (N, C) tensor is ONE contiguous
// span) into a shared-memory ring with cp.async.bulk + mbarriers, compute warps read 16 bytes per thread per
// row from shared memory and keep their per-column state (running min / max, or scale / min / reciprocal
// fragments) in registers; kernels are chained with programmatic dependent launch; pass 1 loads x and base with
// an L2 evict_last policy when both fit, so the encode pass re-reads them from L2.
// Arithmetic is the register-staged kernels' (same helpers): min / max are exact and order-free, so scales and
// codes stay bit-identical to the reference whatever the geometry.
#pragma once

#include <stdlib.h>

#include <unordered_map>

#include "cf_pipe.cuh"

namespace cf {

constexpr int kMmThreads2 = 384 + 32;   // two CTAs per SM
constexpr int kMmThreads1 = 512 + 32;   // one CTA per SM
constexpr size_t kMmSmemBudget = 216 * 1024;

struct MmPipe {
  int TX, TY, G;
  int R;                 // rows per stage (even, multiple of 2 * TY)
  int stages, ctas_per_sm;
  uint32_t tile_bytes;   // R * C * 2
  uint32_t code_tile;    // decode: bytes of codes staged per tile, rounded up to 128
  uint32_t stage_bytes;
  size_t smem_bytes;
  bool ok;
};

static int mm_env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}

// nfull: full-size fp16 operands per stage (2: x + base, 1: x or base alone); code_bytes_per_pair: bytes of codes
// per ROW PAIR staged alongside (decode only)
static MmPipe make_mm_pipe(int64_t C, int nfull, int code_bytes_per_pair) {
  MmPipe g{};
  g.ok = false;
  if (C % 8 != 0 || C < 64 || C > 8192) return g;
  const int groups = static_cast<int>(C / 8);
  const size_t stage_target = static_cast<size_t>(mm_env_int("CF_PIPE_STAGE_KB", 48)) * 1024;
  for (g.ctas_per_sm = 2; g.ctas_per_sm >= 1; --g.ctas_per_sm) {
    const int cap = g.ctas_per_sm == 2 ? kMmThreads2 - 32 : kMmThreads1 - 32;
    g.G = (groups + cap - 1) / cap;
    if (g.G == 3) g.G = 4;
    if (g.G > 4) continue;
    if (g.G > 1 && g.ctas_per_sm == 2) continue;   // several column groups per thread: 72 registers are not enough
    const int per = (groups + g.G - 1) / g.G;
    g.TX = (per + 31) / 32 * 32;
    if (g.TX > cap) continue;
    g.TY = cap / g.TX;
    if (g.TY < 1) g.TY = 1;
    if (g.TY > 8) g.TY = 8;
    const size_t budget = kMmSmemBudget / g.ctas_per_sm;
    for (int k = 4; k >= 1; k /= 2) {   // row pairs per thread per stage
      g.R = 2 * g.TY * k;
      const size_t row_bytes = static_cast<size_t>(C) * 2 * nfull;
      if (k > 1 && static_cast<size_t>(g.R) * row_bytes > stage_target) continue;
      g.tile_bytes = static_cast<uint32_t>(static_cast<size_t>(g.R) * C * 2);
      g.code_tile = static_cast<uint32_t>((static_cast<size_t>(g.R / 2) * code_bytes_per_pair + 127) / 128 * 128);
      g.stage_bytes = g.tile_bytes * nfull + g.code_tile;
      const size_t fixed = 256;
      if (budget < fixed + 2 * static_cast<size_t>(g.stage_bytes)) continue;
      int stages = static_cast<int>((budget - fixed) / g.stage_bytes);
      if (stages > 4) stages = 4;
      g.stages = stages;
      g.smem_bytes = static_cast<size_t>(stages) * g.stage_bytes + fixed;
      g.ok = true;
      return g;
    }
  }
  return g;
}

struct MmArgs {
  int TX, TY, R, stages;
  uint32_t tile_bytes, stage_bytes;
  int rows_per_cta;     // stats: contiguous row range per CTA (multiple of R)
  int tiles_per_cta;    // codec: contiguous tile range per CTA
  int l2_keep;          // stats: load with evict_last (the encode pass re-reads from L2)
};

__device__ __forceinline__ void mm_pipe_init(uint64_t* full, uint64_t* empty, int stages, int ncompute) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], ncompute / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------
// pass 1: per-CTA column min / max of delta = x - base          grid (B), block TX*TY + 32
// partial layout: pmin[b][C], pmax[b][C] (fp16), like k_minmax_stats
// ---------------------------------------------------------------------------------------
template <int G, int OCC, bool HAS_BASE>
__global__ void __launch_bounds__(OCC == 2 ? kMmThreads2 : kMmThreads1, OCC)
    k_minmax_stats_tma(const __half* __restrict__ x, const __half* __restrict__ base, __half* __restrict__ pmin,
                       __half* __restrict__ pmax, int N, int C, const MmArgs a) {
  extern __shared__ __align__(128) unsigned char mm_smem_raw[];
  const int TX = a.TX, TY = a.TY, ncompute = TX * TY;
  unsigned char* stage0 = mm_smem_raw;
  uint64_t* full = reinterpret_cast<uint64_t*>(mm_smem_raw + static_cast<size_t>(a.stages) * a.stage_bytes);
  uint64_t* empty = full + 8;
  const int tid = threadIdx.x;
  const int groups = C >> 3;
  const int r_begin = blockIdx.x * a.rows_per_cta;
  const int r_end = min(N, r_begin + a.rows_per_cta);
  const uint32_t row_bytes = static_cast<uint32_t>(C) * 2u;
  mm_pipe_init(full, empty, a.stages, ncompute);
  pdl_wait();
  pdl_launch_dependents();

  if (tid >= ncompute) {  // producer warp
    if (tid == ncompute) {
      const uint64_t pol = a.l2_keep ? make_policy_evict_last() : 0ull;
      int it = 0;
      for (int r0 = r_begin; r0 < r_end; r0 += a.R, ++it) {
        const int st = it % a.stages, k = it / a.stages;
        if (k > 0) mbar_wait(&empty[st], (k - 1) & 1);
        const uint32_t bytes = static_cast<uint32_t>(min(a.R, r_end - r0)) * row_bytes;
        mbar_arrive_expect_tx(&full[st], HAS_BASE ? 2 * bytes : bytes);
        unsigned char* dst = stage0 + static_cast<size_t>(st) * a.stage_bytes;
        bulk_g2s_pol(dst, reinterpret_cast<const unsigned char*>(x) + static_cast<size_t>(r0) * row_bytes, bytes,
                     &full[st], pol);
        if (HAS_BASE)
          bulk_g2s_pol(dst + a.tile_bytes, reinterpret_cast<const unsigned char*>(base) + static_cast<size_t>(r0) * row_bytes,
                       bytes, &full[st], pol);
      }
    }
    return;
  }

  const int tx = tid % TX, ty = tid / TX, lane = tid & 31;
  const __half2 pinf = __half2half2(__ushort_as_half(0x7C00)), ninf = __half2half2(__ushort_as_half(0xFC00));
  __half2 mn[G][4], mx[G][4];
#pragma unroll
  for (int j = 0; j < G; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) { mn[j][i] = pinf; mx[j][i] = ninf; }
  const uint32_t thr_a = smem_addr(stage0) + static_cast<uint32_t>(tx) * 16u;
  bool act[G];
#pragma unroll
  for (int j = 0; j < G; ++j) act[j] = tx + j * TX < groups;

  int it = 0;
  for (int r0 = r_begin; r0 < r_end; r0 += a.R, ++it) {
    const int st = it % a.stages, k = it / a.stages;
    mbar_wait(&full[st], k & 1);
    const uint32_t xs_a = thr_a + static_cast<uint32_t>(st) * a.stage_bytes;
    const int rows = min(a.R, r_end - r0);
    constexpr int kFly = (G == 1) ? 4 : 2;
    int rl = ty;
    for (; rl + (kFly - 1) * TY < rows; rl += kFly * TY) {  // kFly rows of loads in flight
      H8 d[kFly][G];
#pragma unroll
      for (int f = 0; f < kFly; ++f)
#pragma unroll
        for (int j = 0; j < G; ++j) {
          if (!act[j]) continue;
          const uint32_t o = xs_a + static_cast<uint32_t>(rl + f * TY) * row_bytes + static_cast<uint32_t>(j * TX) * 16u;
          const H8 xv = as_h8(lds128a(o));
          d[f][j] = HAS_BASE ? h8_sub(xv, as_h8(lds128a(o + a.tile_bytes))) : xv;
        }
#pragma unroll
      for (int f = 0; f < kFly; ++f)
#pragma unroll
        for (int j = 0; j < G; ++j) {
          if (!act[j]) continue;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            mn[j][i] = __hmin2(mn[j][i], u2h2(d[f][j].w[i]));
            mx[j][i] = __hmax2(mx[j][i], u2h2(d[f][j].w[i]));
          }
        }
    }
    for (; rl < rows; rl += TY) {
#pragma unroll
      for (int j = 0; j < G; ++j) {
        if (!act[j]) continue;
        const uint32_t o = xs_a + static_cast<uint32_t>(rl) * row_bytes + static_cast<uint32_t>(j * TX) * 16u;
        const H8 xv = as_h8(lds128a(o));
        const H8 d = HAS_BASE ? h8_sub(xv, as_h8(lds128a(o + a.tile_bytes))) : xv;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          mn[j][i] = __hmin2(mn[j][i], u2h2(d.w[i]));
          mx[j][i] = __hmax2(mx[j][i], u2h2(d.w[i]));
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
  }

  // reduce over ty through the drained stage memory: [ty][j][tx][8 words: 4 min, 4 max]
  uint32_t* red = reinterpret_cast<uint32_t*>(stage0);
  if (TY > 1) {
    compute_sync(ncompute);
#pragma unroll
    for (int j = 0; j < G; ++j) {
      uint32_t* dst = red + ((static_cast<size_t>(ty) * G + j) * TX + tx) * 8;
#pragma unroll
      for (int i = 0; i < 4; ++i) { dst[i] = h22u(mn[j][i]); dst[4 + i] = h22u(mx[j][i]); }
    }
    compute_sync(ncompute);
    if (ty == 0) {
      for (int yy = 1; yy < TY; ++yy)
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const uint32_t* src = red + ((static_cast<size_t>(yy) * G + j) * TX + tx) * 8;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            mn[j][i] = __hmin2(mn[j][i], u2h2(src[i]));
            mx[j][i] = __hmax2(mx[j][i], u2h2(src[4 + i]));
          }
        }
    }
  }
  if (ty == 0) {
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const int g = tx + j * TX;
      if (g < groups) {
        const size_t off = static_cast<size_t>(blockIdx.x) * C + 8 * g;
        *reinterpret_cast<uint4*>(pmin + off) = make_uint4(h22u(mn[j][0]), h22u(mn[j][1]), h22u(mn[j][2]), h22u(mn[j][3]));
        *reinterpret_cast<uint4*>(pmax + off) = make_uint4(h22u(mx[j][0]), h22u(mx[j][1]), h22u(mx[j][2]), h22u(mx[j][3]));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// finalize, 8 columns per thread: block 256 = 2 column groups x 128 partial lanes, grid ceil(C / 16).  Every
// thread folds ceil(B / 128) partial rows (<= 3 for B = 296) with independent 16-byte loads, all in flight at
// once; lanes are then combined with xor-shuffles inside a warp and through shared memory across the 8 warps.
// (The 2-byte version is a chain of B / 8 dependent L2 round trips per thread: 11.4 us for 3.6 MB; a first
// 8-column version with only C / 64 = 48 CTAs still took 8.1 us.)
// ---------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) k_minmax_finalize_v8(const __half* __restrict__ pmin, const __half* __restrict__ pmax,
                                                           int B, int C, __half* __restrict__ scale_out,
                                                           void* __restrict__ second_out, __half* __restrict__ min_ws) {
  __shared__ uint4 smn[8][2], smx[8][2];
  const int gx = threadIdx.x & 1, py = threadIdx.x >> 1;
  const int warp = threadIdx.x >> 5;
  const int g = blockIdx.x * 2 + gx;   // column group of 8
  const bool in = 8 * g < C;
  pdl_wait();
  pdl_launch_dependents();
  const __half2 pinf = __half2half2(__ushort_as_half(0x7C00)), ninf = __half2half2(__ushort_as_half(0xFC00));
  __half2 mn[4] = {pinf, pinf, pinf, pinf}, mx[4] = {ninf, ninf, ninf, ninf};
  if (in) {
#pragma unroll 3
    for (int b = py; b < B; b += 128) {
      const uint4 a = *reinterpret_cast<const uint4*>(pmin + static_cast<size_t>(b) * C + 8 * g);
      const uint4 c = *reinterpret_cast<const uint4*>(pmax + static_cast<size_t>(b) * C + 8 * g);
      mn[0] = __hmin2(mn[0], u2h2(a.x)); mn[1] = __hmin2(mn[1], u2h2(a.y));
      mn[2] = __hmin2(mn[2], u2h2(a.z)); mn[3] = __hmin2(mn[3], u2h2(a.w));
      mx[0] = __hmax2(mx[0], u2h2(c.x)); mx[1] = __hmax2(mx[1], u2h2(c.y));
      mx[2] = __hmax2(mx[2], u2h2(c.z)); mx[3] = __hmax2(mx[3], u2h2(c.w));
    }
  }
  // lanes of a warp: 16 partial lanes x 2 column groups (bit 0 of the lane): fold over bits 1..4
#pragma unroll
  for (int o = 2; o <= 16; o <<= 1)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      mn[i] = __hmin2(mn[i], u2h2(__shfl_xor_sync(0xffffffffu, h22u(mn[i]), o)));
      mx[i] = __hmax2(mx[i], u2h2(__shfl_xor_sync(0xffffffffu, h22u(mx[i]), o)));
    }
  if ((threadIdx.x & 31) < 2) {
    smn[warp][gx] = make_uint4(h22u(mn[0]), h22u(mn[1]), h22u(mn[2]), h22u(mn[3]));
    smx[warp][gx] = make_uint4(h22u(mx[0]), h22u(mx[1]), h22u(mx[2]), h22u(mx[3]));
  }
  __syncthreads();
  // 16 threads finish the 16 columns of this CTA: thread t -> column group t / 8, element t % 8
  if (threadIdx.x < 16) {
    const int cg = threadIdx.x >> 3, e = threadIdx.x & 7;
    const int c = (blockIdx.x * 2 + cg) * 8 + e;
    if (c < C) {
      __half m = __ushort_as_half(0x7C00), M = __ushort_as_half(0xFC00);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        m = __hmin(m, reinterpret_cast<const __half*>(&smn[k][cg])[e]);
        M = __hmax(M, reinterpret_cast<const __half*>(&smx[k][cg])[e]);
      }
      finalize_column<MODE>(m, M, c, scale_out, second_out, min_ws);
    }
  }
}

// ---------------------------------------------------------------------------------------
// INT4 / INT2-minmax pass 2 (ENCODE: x, base -> packed (+ out = base + deq)) and decode
// (!ENCODE: packed, base -> out = base + deq).  Flattened tile schedule: CTA b owns tiles
// [b * tiles_per_cta, ...); a tile is R rows = R / 2 row pairs; thread (tx, ty) takes pairs ty, ty + TY, ...
// stage = ENCODE ? [x tile | base tile] : [base tile | code tile (R / 2 rows of C bytes)]
// ---------------------------------------------------------------------------------------
template <int G, int OCC, bool ENCODE, bool HAS_BASE, int LEVELS>
__global__ void __launch_bounds__(OCC == 2 ? kMmThreads2 : kMmThreads1, OCC)
    k_int4_codec_tma(const __half* __restrict__ x, const __half* __restrict__ base, const __half* __restrict__ scale,
                     const __half* __restrict__ minv, uint8_t* __restrict__ packed, __half* __restrict__ out, int N,
                     int C, const MmArgs a, const int l2_hints) {
  extern __shared__ __align__(128) unsigned char mm_smem_raw[];
  const int TX = a.TX, TY = a.TY, ncompute = TX * TY;
  unsigned char* stage0 = mm_smem_raw;
  uint64_t* full = reinterpret_cast<uint64_t*>(mm_smem_raw + static_cast<size_t>(a.stages) * a.stage_bytes);
  uint64_t* empty = full + 8;
  const int tid = threadIdx.x;
  const int groups = C >> 3;
  const int total_tiles = (N + a.R - 1) / a.R;
  const int T0 = blockIdx.x * a.tiles_per_cta;
  const int T1 = min(total_tiles, T0 + a.tiles_per_cta);
  const uint32_t row_bytes = static_cast<uint32_t>(C) * 2u;
  // operand A of a stage: x (encode) or base (decode); operand B: base (encode) or the code rows (decode)
  constexpr bool kTwoFull = ENCODE && HAS_BASE;
  mm_pipe_init(full, empty, a.stages, ncompute);
  pdl_wait();
  pdl_launch_dependents();

  if (tid >= ncompute) {
    if (tid == ncompute) {
      const uint64_t pol = l2_hints ? make_policy_evict_first() : 0ull;   // last use of these lines
      for (int tile = T0, it = 0; tile < T1; ++tile, ++it) {
        const int st = it % a.stages, k = it / a.stages;
        if (k > 0) mbar_wait(&empty[st], (k - 1) & 1);
        const int r0 = tile * a.R;
        const uint32_t rows = static_cast<uint32_t>(min(a.R, N - r0));
        const uint32_t bytes = rows * row_bytes;
        unsigned char* dst = stage0 + static_cast<size_t>(st) * a.stage_bytes;
        if (ENCODE) {
          mbar_arrive_expect_tx(&full[st], kTwoFull ? 2 * bytes : bytes);
          bulk_g2s_pol(dst, reinterpret_cast<const unsigned char*>(x) + static_cast<size_t>(r0) * row_bytes, bytes,
                       &full[st], pol);
          if (HAS_BASE)
            bulk_g2s_pol(dst + a.tile_bytes, reinterpret_cast<const unsigned char*>(base) + static_cast<size_t>(r0) * row_bytes,
                         bytes, &full[st], pol);
        } else {
          const uint32_t cbytes = (rows / 2) * static_cast<uint32_t>(C);
          mbar_arrive_expect_tx(&full[st], (HAS_BASE ? bytes : 0u) + cbytes);
          if (HAS_BASE)
            bulk_g2s_pol(dst, reinterpret_cast<const unsigned char*>(base) + static_cast<size_t>(r0) * row_bytes, bytes,
                         &full[st], pol);
          bulk_g2s_pol(dst + (HAS_BASE ? a.tile_bytes : 0u), packed + static_cast<size_t>(r0 / 2) * C, cbytes, &full[st], pol);
        }
      }
    }
    return;
  }

  const int tx = tid % TX, ty = tid / TX, lane = tid & 31;
  uint32_t sfrag[G][4], mfrag[G][4];
  float rfrag[G][ENCODE ? 8 : 1];
#pragma unroll
  for (int j = 0; j < G; ++j) {
    const int g = tx + j * TX;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sfrag[j][i] = (g < groups) ? ld_h2(scale, 8 * g + 2 * i) : 0x3C003C00u;
      mfrag[j][i] = (g < groups) ? ld_h2(minv, 8 * g + 2 * i) : 0u;
      if (ENCODE) {
        rfrag[j][2 * i] = __frcp_rn(__half2float(lo_h(sfrag[j][i])));
        rfrag[j][2 * i + 1] = __frcp_rn(__half2float(hi_h(sfrag[j][i])));
      }
    }
  }
  const uint32_t stage_a = smem_addr(stage0) + static_cast<uint32_t>(tx) * 16u;
  const uint64_t st_pol = l2_hints ? make_policy_evict_first() : 0ull;

  for (int tile = T0, it = 0; tile < T1; ++tile, ++it) {
    const int st = it % a.stages, k = it / a.stages;
    const int r0 = tile * a.R;
    const int pairs = min(a.R, N - r0) >> 1;
    mbar_wait(&full[st], k & 1);
    const uint32_t sa = stage_a + static_cast<uint32_t>(st) * a.stage_bytes;
    for (int pi = ty; pi < pairs; pi += TY) {
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const int g = tx + j * TX;
        if (g >= groups) continue;
        const uint32_t o0 = sa + static_cast<uint32_t>(2 * pi) * row_bytes + static_cast<uint32_t>(j * TX) * 16u;
        const uint32_t o1 = o0 + row_bytes;
        const uint32_t base_off = ENCODE ? a.tile_bytes : 0u;
        H8 b0h, b1h;
        if (HAS_BASE) {
          b0h = as_h8(lds128a(o0 + base_off));
          b1h = as_h8(lds128a(o1 + base_off));
        } else {
          b0h = as_h8(make_uint4(0, 0, 0, 0));
          b1h = b0h;
        }
        uint32_t r0c[4], r1c[4];  // 1024 + code, two columns per word
        const size_t row0 = static_cast<size_t>(r0 + 2 * pi);
        if (ENCODE) {
          const H8 x0 = as_h8(lds128a(o0)), x1 = as_h8(lds128a(o1));
          const H8 d0 = HAS_BASE ? h8_sub(x0, b0h) : x0, d1 = HAS_BASE ? h8_sub(x1, b1h) : x1;
          uint32_t m[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            r0c[i] = int4_codes2<LEVELS>(d0.w[i], mfrag[j][i], sfrag[j][i], rfrag[j][2 * i], rfrag[j][2 * i + 1]);
            r1c[i] = int4_codes2<LEVELS>(d1.w[i], mfrag[j][i], sfrag[j][i], rfrag[j][2 * i], rfrag[j][2 * i + 1]);
            m[i] = (r0c[i] & 0x000F000Fu) | ((r1c[i] & 0x000F000Fu) << 4);  // low nibble = even row  :573
          }
          *reinterpret_cast<uint2*>(packed + (row0 >> 1) * C + 8 * g) =
              make_uint2(__byte_perm(m[0], m[1], 0x6420), __byte_perm(m[2], m[3], 0x6420));
        } else {
          // code rows of the tile: pair pi at byte offset pi * C behind the base tile
          const uint32_t ca = smem_addr(stage0) + static_cast<uint32_t>(st) * a.stage_bytes + (HAS_BASE ? a.tile_bytes : 0u) +
                              static_cast<uint32_t>(pi) * static_cast<uint32_t>(C) + static_cast<uint32_t>(g) * 8u;
          uint2 pv;
          asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(pv.x), "=r"(pv.y) : "r"(ca));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t w = __byte_perm(i < 2 ? pv.x : pv.y, 0u, (i & 1) ? 0x4342u : 0x4140u);  // [b(2i), 0, b(2i+1), 0]
            r0c[i] = (w & 0x000F000Fu) | 0x64006400u;
            r1c[i] = ((w >> 4) & 0x000F000Fu) | 0x64006400u;
          }
        }
        if (out != nullptr) {
          H8 q0, q1;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const __half2 v0 = int4_values2(r0c[i], mfrag[j][i], sfrag[j][i]);
            const __half2 v1 = int4_values2(r1c[i], mfrag[j][i], sfrag[j][i]);
            q0.w[i] = h22u(HAS_BASE ? __hadd2_rn(u2h2(b0h.w[i]), v0) : v0);
            q1.w[i] = h22u(HAS_BASE ? __hadd2_rn(u2h2(b1h.w[i]), v1) : v1);
          }
          stg_stream_pol(out + row0 * C + 8 * g, as_u4(q0), st_pol);
          stg_stream_pol(out + (row0 + 1) * C + 8 * g, as_u4(q1), st_pol);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static bool mm_tma_enabled() {
  const char* e = getenv("CF_LEGACY_KERNELS");
  return !(e && e[0] == '1');
}
static bool mm_pdl_enabled() {
  const char* e = getenv("CF_PDL");
  return !(e && e[0] == '0');
}

template <typename... KArgs, typename... Args>
static cudaError_t mm_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mm_pdl_enabled() ? 1 : 0;
  if (smem > 48 * 1024) {
    static thread_local std::unordered_map<const void*, size_t> granted;
    size_t& g = granted[reinterpret_cast<const void*>(kern)];
    if (smem > g) {
      cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(kern),
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return e;
      g = smem;
    }
  }
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

struct MmStatsPlan {
  MmPipe pipe;
  int B, rows_per_cta;
};
// the statistics pass of the pipelined path; B <= 2 * SMs, so the register-staged plan's workspace covers it
static MmStatsPlan make_mm_stats_plan(int64_t N, int64_t C, bool has_base) {
  MmStatsPlan pl{};
  pl.pipe = make_mm_pipe(C, has_base ? 2 : 1, 0);
  if (!pl.pipe.ok) return pl;
  int B = sm_count() * pl.pipe.ctas_per_sm;
  int64_t rpc = (N + B - 1) / B;
  rpc = (rpc + pl.pipe.R - 1) / pl.pipe.R * pl.pipe.R;
  pl.rows_per_cta = static_cast<int>(rpc);
  pl.B = static_cast<int>((N + rpc - 1) / rpc);
  return pl;
}

static MmArgs mm_args(const MmPipe& g) {
  MmArgs a{};
  a.TX = g.TX; a.TY = g.TY; a.R = g.R; a.stages = g.stages;
  a.tile_bytes = g.tile_bytes; a.stage_bytes = g.stage_bytes;
  return a;
}

// x and base of this call fit in L2 together: pass 1 keeps them (evict_last), pass 2 re-reads them from L2
static int mm_l2_keep(int64_t N, int64_t C, bool has_base) {
  if (mm_env_int("CF_L2_HINTS", 1) == 0) return 0;
  const int64_t cap = static_cast<int64_t>(mm_env_int("CF_L2_KEEP_MB", 72)) << 20;
  return N * C * 2 * (has_base ? 2 : 1) <= cap ? 1 : 0;
}

template <bool HAS_BASE>
static int launch_mm_stats_tma(const MmStatsPlan& pl, const __half* x, const __half* base, __half* pmin, __half* pmax,
                               int n, int c, cudaStream_t st) {
  MmArgs a = mm_args(pl.pipe);
  a.rows_per_cta = pl.rows_per_cta;
  a.l2_keep = mm_l2_keep(n, c, HAS_BASE);
  dim3 grid(pl.B), block(pl.pipe.TX * pl.pipe.TY + 32);
  const int variant = (pl.pipe.G == 1 ? 0 : (pl.pipe.G == 2 ? 2 : 4)) + (pl.pipe.ctas_per_sm == 2 ? 1 : 0);
#define CF_MM_ST(GG, OO) CF_CHECK_CUDA(mm_launch(k_minmax_stats_tma<GG, OO, HAS_BASE>, grid, block, pl.pipe.smem_bytes, st, x, base, pmin, pmax, n, c, a))
  switch (variant) {
    case 0: CF_MM_ST(1, 1); break;
    case 1: CF_MM_ST(1, 2); break;
    case 2: CF_MM_ST(2, 1); break;
    case 3: CF_MM_ST(2, 2); break;
    case 4: CF_MM_ST(4, 1); break;
    default: CF_MM_ST(4, 2); break;
  }
#undef CF_MM_ST
  return CF_OK;
}

template <bool ENCODE, bool HAS_BASE, int LEVELS>
static int launch_int4_codec_tma(const MmPipe& pg, const __half* x, const __half* base, const __half* scale,
                                 const __half* minv, uint8_t* packed, __half* out, int n, int c, cudaStream_t st) {
  MmArgs a = mm_args(pg);
  const int total_tiles = (n + pg.R - 1) / pg.R;
  int ctas = sm_count() * pg.ctas_per_sm;
  if (ctas > total_tiles) ctas = total_tiles;
  a.tiles_per_cta = (total_tiles + ctas - 1) / ctas;
  const int n_cta = (total_tiles + a.tiles_per_cta - 1) / a.tiles_per_cta;
  const int hints = mm_env_int("CF_L2_HINTS", 1) != 0 ? 1 : 0;
  dim3 grid(n_cta), block(pg.TX * pg.TY + 32);
  const int variant = (pg.G == 1 ? 0 : (pg.G == 2 ? 2 : 4)) + (pg.ctas_per_sm == 2 ? 1 : 0);
#define CF_MM_CD(GG, OO) CF_CHECK_CUDA(mm_launch(k_int4_codec_tma<GG, OO, ENCODE, HAS_BASE, LEVELS>, grid, block, pg.smem_bytes, st, x, base, scale, minv, packed, out, n, c, a, hints))
  switch (variant) {
    case 0: CF_MM_CD(1, 1); break;
    case 1: CF_MM_CD(1, 2); break;
    case 2: CF_MM_CD(2, 1); break;
    case 3: CF_MM_CD(2, 2); break;
    case 4: CF_MM_CD(4, 1); break;
    default: CF_MM_CD(4, 2); break;
  }
#undef CF_MM_CD
  return CF_OK;
}

}  // namespace cf
