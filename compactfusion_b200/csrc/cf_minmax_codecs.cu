// INT4 and INT8 per-channel min/max affine codecs for sm_100a.
//
// Reference semantics (eager == the reference's ground truth):
//   INT4  quantize_int4 / dequantize_int4 / sim_int4(dim=0)   compress_quantize.py:487-640
//   INT8  quantize_int8 / dequantize_int8                     compress_quantize.py:428-484
// The reference only *simulates* INT4 on residuals (slowpath.py:205-206) and uses INT8 for a
// deprecated cache; here both are real wire codecs with the residual subtract and the
// error-feedback update fused in.  min/max are exact, so scales and codes are bit-exact.
//
// Two streaming passes: (1) per-column min/max of delta = x - base, (2) encode (+ optional
// new_base = base + dequant).  All fp16 arithmetic rounds once per reference op; the
// divisions are IEEE fp32 divisions followed by one rounding to fp16 (what eager torch does).
#include "cf_common.cuh"

namespace cf {

// MODE_INT2MM: the reference's simulation-only 4-level min/max quantiser (sim_int2_minmax,
// compress_quantize.py:386-426) -- the INT4 arithmetic with qmax = 3 instead of 15; codes travel in nibbles
enum { MODE_INT4 = 0, MODE_INT8 = 1, MODE_INT2MM = 2 };
template <int MODE> struct MmLevels { static constexpr int value = (MODE == MODE_INT2MM) ? 3 : 15; };
__host__ __device__ constexpr int mm_unroll_for(int G) { return G == 1 ? 4 : (G == 2 ? 2 : 1); }

__device__ __forceinline__ __half hdiv_exact(__half a, __half b) {
  return __float2half_rn(__fdiv_rn(__half2float(a), __half2float(b)));
}

// ---------------------------------------------------------------------------------------
// pass 1: per-CTA column min / max of delta.   grid (B), block (TX, TY)
// partial layout: pmin[b][C], pmax[b][C] (fp16)
// ---------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(512) k_minmax_stats(const __half* __restrict__ x,
                                                      const __half* __restrict__ base,
                                                      __half* __restrict__ pmin, __half* __restrict__ pmax,
                                                      int N, int C, int rows_per_cta) {
  extern __shared__ uint32_t sm_u32[];
  constexpr int kMMUnroll = mm_unroll_for(G);
  const int groups = C >> 3;
  const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(N, r_begin + rows_per_cta);
  const __half2 pinf = __half2half2(__ushort_as_half(0x7C00)), ninf = __half2half2(__ushort_as_half(0xFC00));
  __half2 mn[G][4], mx[G][4];
#pragma unroll
  for (int j = 0; j < G; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) { mn[j][i] = pinf; mx[j][i] = ninf; }

  for (int r = r_begin + ty; r < r_end; r += TY * kMMUnroll) {
    uint4 xv[kMMUnroll][G], bv[kMMUnroll][G];
#pragma unroll
    for (int u = 0; u < kMMUnroll; ++u) {
      const int rr = r + u * TY;
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const int g = tx + j * TX;
        xv[u][j] = make_uint4(0, 0, 0, 0);
        bv[u][j] = make_uint4(0, 0, 0, 0);
        if (rr < r_end && g < groups) {
          const size_t off = static_cast<size_t>(rr) * C + 8 * g;
          xv[u][j] = ldg_stream(x + off);
          if (base != nullptr) bv[u][j] = ldg_stream(base + off);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kMMUnroll; ++u) {
      const int rr = r + u * TY;
      if (rr < r_end) {
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const H8 d = h8_sub(as_h8(xv[u][j]), as_h8(bv[u][j]));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            mn[j][i] = __hmin2(mn[j][i], u2h2(d.w[i]));
            mx[j][i] = __hmax2(mx[j][i], u2h2(d.w[i]));
          }
        }
      }
    }
  }
  // reduce over ty through smem: layout [ty][j][tx][8 words: 4 min, 4 max]
  if (TY > 1) {
#pragma unroll
    for (int j = 0; j < G; ++j) {
      uint32_t* dst = sm_u32 + ((static_cast<size_t>(ty) * G + j) * TX + tx) * 8;
#pragma unroll
      for (int i = 0; i < 4; ++i) { dst[i] = h22u(mn[j][i]); dst[4 + i] = h22u(mx[j][i]); }
    }
    __syncthreads();
    if (ty == 0) {
      for (int yy = 1; yy < TY; ++yy) {
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const uint32_t* src = sm_u32 + ((static_cast<size_t>(yy) * G + j) * TX + tx) * 8;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            mn[j][i] = __hmin2(mn[j][i], u2h2(src[i]));
            mx[j][i] = __hmax2(mx[j][i], u2h2(src[4 + i]));
          }
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
// finalize: min/max over partials -> scale, min (INT4) or scale, zero_point (INT8)
// ---------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void finalize_column(__half mn, __half mx, int c, __half* __restrict__ scale_out,
                                                void* __restrict__ second_out, __half* __restrict__ min_ws) {
  const __half diff = __hsub_rn(mx, mn);  // (max_val - min_val) in fp16
  if (MODE == MODE_INT4 || MODE == MODE_INT2MM) {
    // scale = (max - min) / (15 + 1e-6): fp16 tensor / python scalar = fp32 divide by float(15.000001)
    // (INT2MM: qmax - qmin + 1e-6 = 3.000001, compress_quantize.py:411)
    const __half s = __float2half_rn(__fdiv_rn(__half2float(diff), MODE == MODE_INT4 ? 15.000001f : 3.000001f));  // compress_quantize.py:556
    scale_out[c] = s;
    static_cast<__half*>(second_out)[c] = mn;
  } else {
    const __half s = __float2half_rn(__fdiv_rn(__half2float(diff), 255.000001f));  // compress_quantize.py:455
    scale_out[c] = s;
    // zero_point = clamp(qmin - round(min / scale), qmin, qmax).to(int16)   (:460-463)
    const float t1 = __half2float(hdiv_exact(mn, s));
    const float t2 = __half2float(__float2half_rn(rintf(t1)));
    float zp = __half2float(__float2half_rn(-128.f - t2));
    zp = fminf(fmaxf(zp, -128.f), 127.f);
    if (t1 != t1) zp = 0.f;  // NaN (zero scale): the reference's cast is undefined, we define 0
    static_cast<int16_t*>(second_out)[c] = static_cast<int16_t>(zp);
    if (min_ws) min_ws[c] = mn;
  }
}

// grid ceil(C/32), block 256 = 32 columns x 8 partial lanes: every thread folds ceil(B/8) partials with
// independent loads (the old one-thread-per-column loop was a chain of B dependent L2 round trips)
template <int MODE>
__global__ void __launch_bounds__(256) k_minmax_finalize(const __half* __restrict__ pmin,
                                                         const __half* __restrict__ pmax, int B, int C,
                                                         __half* __restrict__ scale_out,
                                                         void* __restrict__ second_out,
                                                         __half* __restrict__ min_ws) {
  __shared__ __half smn[8][33], smx[8][33];
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  __half mn = __ushort_as_half(0x7C00), mx = __ushort_as_half(0xFC00);
  if (c < C) {
#pragma unroll 4
    for (int b = py; b < B; b += 8) {
      mn = __hmin(mn, pmin[static_cast<size_t>(b) * C + c]);
      mx = __hmax(mx, pmax[static_cast<size_t>(b) * C + c]);
    }
  }
  smn[py][cx] = mn;
  smx[py][cx] = mx;
  __syncthreads();
  if (py == 0 && c < C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      mn = __hmin(mn, smn[k][cx]);
      mx = __hmax(mx, smx[k][cx]);
    }
    finalize_column<MODE>(mn, mx, c, scale_out, second_out, min_ws);
  }
}

// ---- generic path for C % 8 != 0 (tall-skinny low-rank factors, LOW_RANK_Q: slowpath.py:69-70) ----
// one CTA per column: block min/max over the N rows, then the same finalize arithmetic
template <int MODE>
__global__ void __launch_bounds__(256) k_minmax_column_generic(const __half* __restrict__ x,
                                                               const __half* __restrict__ base, int N, int C,
                                                               __half* __restrict__ scale_out,
                                                               void* __restrict__ second_out) {
  __shared__ __half smn[256], smx[256];
  const int c = blockIdx.x, t = threadIdx.x;
  __half mn = __ushort_as_half(0x7C00), mx = __ushort_as_half(0xFC00);
  for (int n = t; n < N; n += 256) {
    const size_t i = static_cast<size_t>(n) * C + c;
    const __half d = base ? __hsub_rn(x[i], base[i]) : x[i];
    mn = __hmin(mn, d);
    mx = __hmax(mx, d);
  }
  smn[t] = mn; smx[t] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) { smn[t] = __hmin(smn[t], smn[t + o]); smx[t] = __hmax(smx[t], smx[t + o]); }
    __syncthreads();
  }
  if (t == 0) finalize_column<MODE>(smn[0], smx[0], c, scale_out, second_out, nullptr);
}

// ---------------------------------------------------------------------------------------
// Exact fp16 quotient without a division per element.
//
// Eager torch computes fp16 a / s as RN16(RN32(float(a) / float(s))).  With a per-column reciprocal
// rcp = RN32(1 / s), t = RN32(a * rcp) is within 2 fp32 ulps of the true quotient, so RN16(t) can differ
// from the reference only when t lies within a few ulps of a rounding boundary of the fp16 grid (the 13
// discarded mantissa bits within +-4 of 0x1000), or when the result is subnormal in fp16, zero, infinite
// or NaN.  Those elements (< 0.1 % on activations) take the IEEE division; every other element costs one
// multiply and three integer instructions.  (Checked against the division for all fp16 numerators x 3000
// scales and 1.6e8 random pairs: tools-free numpy experiment recorded in DESIGN.md section 4.)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float quot_for_rn16(float a, float rcp, __half s_h) {
  float t = a * rcp;
  const uint32_t b = __float_as_uint(t);
  const uint32_t e = (b >> 23) & 0xFFu;
  if (((b - 0x0FFCu) & 0x1FFFu) <= 8u || e - 113u > 29u)  // near a tie | outside the normal fp16 range
    t = __fdiv_rn(a, __half2float(s_h));
  return t;
}

__device__ __forceinline__ uint32_t ld_h2(const __half* v, int c) {  // 2-byte aligned pair load
  return static_cast<uint32_t>(__half_as_ushort(v[c])) | (static_cast<uint32_t>(__half_as_ushort(v[c + 1])) << 16);
}
__device__ __forceinline__ __half lo_h(uint32_t w) { return __ushort_as_half(static_cast<unsigned short>(w & 0xFFFFu)); }
__device__ __forceinline__ __half hi_h(uint32_t w) { return __ushort_as_half(static_cast<unsigned short>(w >> 16)); }

// ---------------------------------------------------------------------------------------
// INT4 pass 2 / decode.  A thread handles the row pair (2i, 2i+1) of its column groups; all
// fp16 arithmetic is packed (two columns per instruction).
//   code  = clamp(rne(fp16((d - min) / scale)), 0, 15)        compress_quantize.py:561-564
//   value = fp16(fp16(code * scale) + min)                    compress_quantize.py:636
// Rounding to integer: clamp first (monotone, integer bounds), then add 1024.0 in fp16 -- the sum lies in
// [1024, 2048) where the fp16 spacing is 1, so the addition itself rounds half-to-even and the code is the
// low mantissa bits.  NaN -> 0 (hmax2 returns the non-NaN operand).
// ---------------------------------------------------------------------------------------
template <int LEVELS = 15>
__device__ __forceinline__ uint32_t int4_codes2(uint32_t d2, uint32_t mn2, uint32_t s2, float rcp0, float rcp1) {
  const float2 a = __half22float2(__hsub2_rn(u2h2(d2), u2h2(mn2)));  // (input - min_val)  :561
  // fp16(a / s) = RN16(RN32(a / s)) without the division: t = RN32(a * RN32(1 / s)) and RN32(a / s) both lie within
  // 2^-22 (relative) of the true quotient, so both lie inside [t (1 - 2^-21), t (1 + 2^-21)]; RN16 is monotone, so
  // when the two ends round to the same fp16 number that number is the answer (99.9 % of the elements; two packed
  // conversions and one compare per PAIR instead of seven integer instructions per element).  Otherwise divide.
  // Infinite / NaN t (zero scale) gives identical ends as well and falls through to the clamp like the division
  // would.  (tools/check_int4_bracket.py: every fp16 numerator x 3000 scales, 0 accepted-but-wrong.)
  const float t0 = a.x * rcp0, t1 = a.y * rcp1;
  constexpr float kLo = 1.f - 0x1p-21f, kHi = 1.f + 0x1p-21f;
  __half2 h = __floats2half2_rn(t0 * kLo, t1 * kLo);                  // fp16(. / scale)
  const __half2 h_hi = __floats2half2_rn(t0 * kHi, t1 * kHi);
  if (h22u(h) != h22u(h_hi))
    h = __floats2half2_rn(__fdiv_rn(a.x, __half2float(lo_h(s2))), __fdiv_rn(a.y, __half2float(hi_h(s2))));
  h = __hmin2(__hmax2(h, __float2half2_rn(0.f)), __float2half2_rn(static_cast<float>(LEVELS)));
  return h22u(__hadd2_rn(h, __float2half2_rn(1024.f)));              // 0x6400 + code per half
}
// biased (1024 + code) pair -> fp16(code * scale) + min
__device__ __forceinline__ __half2 int4_values2(uint32_t r2, uint32_t mn2, uint32_t s2) {
  const __half2 q = __hsub2_rn(u2h2(r2), __float2half2_rn(1024.f));  // exact
  return __hadd2_rn(__hmul2_rn(q, u2h2(s2)), u2h2(mn2));
}

template <int G, bool ENCODE, int LEVELS = 15>
__global__ void __launch_bounds__(512) k_int4_codec(const __half* __restrict__ x, const __half* __restrict__ base,
                                                    const __half* __restrict__ scale, const __half* __restrict__ minv,
                                                    uint8_t* __restrict__ packed, __half* __restrict__ out,
                                                    int N, int C) {
  // ENCODE: x, base -> packed (+ out = base + deq if out != null)
  // !ENCODE: packed, base -> out = base + deq
  const int groups = C >> 3;
  const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
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
  const int pairs = N >> 1;
  const int pair_stride = gridDim.x * TY;
  for (int pi = blockIdx.x * TY + ty; pi < pairs; pi += pair_stride) {
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const int g = tx + j * TX;
      if (g >= groups) continue;
      const size_t off0 = static_cast<size_t>(2 * pi) * C + 8 * g, off1 = off0 + C;
      uint4 b0 = make_uint4(0, 0, 0, 0), b1 = b0;
      if (base != nullptr) { b0 = ldg_stream(base + off0); b1 = ldg_stream(base + off1); }
      const H8 b0h = as_h8(b0), b1h = as_h8(b1);
      uint32_t r0[4], r1[4];  // 1024 + code, two columns per word
      uint8_t* pk = packed + static_cast<size_t>(pi) * C + 8 * g;
      if (ENCODE) {
        const uint4 x0 = ldg_stream(x + off0), x1 = ldg_stream(x + off1);
        const H8 d0 = h8_sub(as_h8(x0), b0h), d1 = h8_sub(as_h8(x1), b1h);
        uint32_t m[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          r0[i] = int4_codes2<LEVELS>(d0.w[i], mfrag[j][i], sfrag[j][i], rfrag[j][2 * i], rfrag[j][2 * i + 1]);
          r1[i] = int4_codes2<LEVELS>(d1.w[i], mfrag[j][i], sfrag[j][i], rfrag[j][2 * i], rfrag[j][2 * i + 1]);
          m[i] = (r0[i] & 0x000F000Fu) | ((r1[i] & 0x000F000Fu) << 4);  // low nibble = even row  :573
        }
        // byte of column 2i sits in bits 0-7 of m[i], column 2i+1 in bits 16-23
        *reinterpret_cast<uint2*>(pk) = make_uint2(__byte_perm(m[0], m[1], 0x6420), __byte_perm(m[2], m[3], 0x6420));
      } else {
        const uint2 pv = *reinterpret_cast<const uint2*>(pk);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t w = __byte_perm(i < 2 ? pv.x : pv.y, 0u, (i & 1) ? 0x4342u : 0x4140u);  // [b(2i), 0, b(2i+1), 0]
          r0[i] = (w & 0x000F000Fu) | 0x64006400u;
          r1[i] = ((w >> 4) & 0x000F000Fu) | 0x64006400u;
        }
      }
      if (out != nullptr) {
        H8 o0, o1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const __half2 v0 = int4_values2(r0[i], mfrag[j][i], sfrag[j][i]);
          const __half2 v1 = int4_values2(r1[i], mfrag[j][i], sfrag[j][i]);
          o0.w[i] = h22u((base != nullptr) ? __hadd2_rn(u2h2(b0h.w[i]), v0) : v0);
          o1.w[i] = h22u((base != nullptr) ? __hadd2_rn(u2h2(b1h.w[i]), v1) : v1);
        }
        stg_stream(out + off0, as_u4(o0));
        stg_stream(out + off1, as_u4(o1));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// INT8 pass 2 / decode (packed fp16, same exact-quotient and magic-add rounding; the bias is 1536 so
// that 1536 + q stays inside [1024, 2048) for q in [-128, 127]; the code byte is the low mantissa byte)
//   q     = clamp(rne(fp16(fp16(x / scale) + zero_point)), -128, 127)     compress_quantize.py:465-467
//   value = fp16((q - zero_point) * scale)                                compress_quantize.py:482
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t int8_codes2(uint32_t d2, uint32_t s2, uint32_t zp2, float rcp0, float rcp1) {
  const float2 a = __half22float2(u2h2(d2));
  float t0 = quot_for_rn16(a.x, rcp0, lo_h(s2));
  float t1 = quot_for_rn16(a.y, rcp1, hi_h(s2));
  // NaN (zero scale with a zero numerator, or NaN input): the reference's integer cast is undefined,
  // we define code 0: fp16(-zp) + zp == 0
  if (t0 != t0) t0 = -__half2float(lo_h(zp2));
  if (t1 != t1) t1 = -__half2float(hi_h(zp2));
  __half2 h = __hadd2_rn(__floats2half2_rn(t0, t1), u2h2(zp2));
  h = __hmin2(__hmax2(h, __float2half2_rn(-128.f)), __float2half2_rn(127.f));
  return h22u(__hadd2_rn(h, __float2half2_rn(1536.f)));  // 0x6400 + 512 + q per half
}
__device__ __forceinline__ __half2 int8_values2(uint32_t r2, uint32_t s2, uint32_t zp2) {
  const __half2 q = __hsub2_rn(u2h2(r2), __float2half2_rn(1536.f));  // exact
  return __hmul2_rn(__hsub2_rn(q, u2h2(zp2)), u2h2(s2));
}

template <int G, bool ENCODE>
__global__ void __launch_bounds__(512) k_int8_codec(const __half* __restrict__ x, const __half* __restrict__ base,
                                                    const __half* __restrict__ scale,
                                                    const int16_t* __restrict__ zpv, int8_t* __restrict__ qout,
                                                    __half* __restrict__ out, int N, int C) {
  const int groups = C >> 3;
  const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
  uint32_t sfrag[G][4], zfrag[G][4];
  float rfrag[G][ENCODE ? 8 : 1];
#pragma unroll
  for (int j = 0; j < G; ++j) {
    const int g = tx + j * TX;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sfrag[j][i] = (g < groups) ? ld_h2(scale, 8 * g + 2 * i) : 0x3C003C00u;
      zfrag[j][i] = (g < groups) ? h22u(__halves2half2(__short2half_rn(zpv[8 * g + 2 * i]),
                                                        __short2half_rn(zpv[8 * g + 2 * i + 1])))
                                 : 0u;
      if (ENCODE) {
        rfrag[j][2 * i] = __frcp_rn(__half2float(lo_h(sfrag[j][i])));
        rfrag[j][2 * i + 1] = __frcp_rn(__half2float(hi_h(sfrag[j][i])));
      }
    }
  }
  const int row_stride = gridDim.x * TY;
  for (int r = blockIdx.x * TY + ty; r < N; r += row_stride) {
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const int g = tx + j * TX;
      if (g >= groups) continue;
      const size_t off = static_cast<size_t>(r) * C + 8 * g;
      uint4 b = make_uint4(0, 0, 0, 0);
      if (base != nullptr) b = ldg_stream(base + off);
      const H8 bh = as_h8(b);
      uint32_t rq[4];  // 1536 + q, two columns per word
      int8_t* qp = qout + off;
      if (ENCODE) {
        const H8 d = h8_sub(as_h8(ldg_stream(x + off)), bh);
        uint32_t m[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          rq[i] = int8_codes2(d.w[i], sfrag[j][i], zfrag[j][i], rfrag[j][2 * i], rfrag[j][2 * i + 1]);
          m[i] = rq[i] & 0x00FF00FFu;  // two's-complement byte of q: 512 is a multiple of 256
        }
        *reinterpret_cast<uint2*>(qp) = make_uint2(__byte_perm(m[0], m[1], 0x6420), __byte_perm(m[2], m[3], 0x6420));
      } else {
        const uint2 pv = *reinterpret_cast<const uint2*>(qp);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t w = __byte_perm(i < 2 ? pv.x : pv.y, 0u, (i & 1) ? 0x4342u : 0x4140u);  // [b(2i), 0, b(2i+1), 0]
          // mantissa of 1536 + q is 512 + q: the byte, with bits 9:8 = 10 for q >= 0 and 01 for q < 0
          const uint32_t neg = (w >> 7) & 0x00010001u;
          rq[i] = ((w | 0x02000200u) ^ (neg * 0x300u)) | 0x64006400u;
        }
      }
      if (out != nullptr) {
        H8 o;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const __half2 v = int8_values2(rq[i], sfrag[j][i], zfrag[j][i]);
          o.w[i] = h22u((base != nullptr) ? __hadd2_rn(u2h2(bh.w[i]), v) : v);
        }
        stg_stream(out + off, as_u4(o));
      }
    }
  }
}

// scalar forms for the generic (C % 8 != 0) path
template <int LEVELS = 15>
__device__ __forceinline__ uint32_t int4_code(__half d, __half mn, __half s) {
  const __half a = __hsub_rn(d, mn);                     // (input - min_val)          :561
  const float q = rintf(__half2float(hdiv_exact(a, s))); // round(. / scale), half-even :561
  return static_cast<uint32_t>(fminf(fmaxf(q, 0.f), static_cast<float>(LEVELS)));  // clamp; NaN -> 0 :564
}
__device__ __forceinline__ __half int4_value(uint32_t q, __half mn, __half s) {
  return __hadd_rn(__hmul_rn(__ushort2half_rn(static_cast<unsigned short>(q)), s), mn);  // q*scale + min :636
}
__device__ __forceinline__ int int8_code(__half d, __half s, __half zp_h) {
  // q = clamp(round(x / scale + zero_point), -128, 127)     compress_quantize.py:465-467
  const __half t = __hadd_rn(hdiv_exact(d, s), zp_h);
  const float q = rintf(__half2float(t));
  if (q != q) return 0;
  return static_cast<int>(fminf(fmaxf(q, -128.f), 127.f));
}
__device__ __forceinline__ __half int8_value(int q, __half s, __half zp_h) {
  // (q.half() - zero_point.half()) * scale                  compress_quantize.py:482
  return __hmul_rn(__hsub_rn(__short2half_rn(static_cast<short>(q)), zp_h), s);
}

// ---- generic element-per-thread codecs for C % 8 != 0 --------------------------------------
template <bool ENCODE, int LEVELS = 15>
__global__ void __launch_bounds__(256) k_int4_codec_generic(const __half* __restrict__ x, const __half* __restrict__ base,
                                                            const __half* __restrict__ scale, const __half* __restrict__ minv,
                                                            uint8_t* __restrict__ packed, __half* __restrict__ out,
                                                            int N, int C) {
  const size_t total = static_cast<size_t>(N / 2) * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int pi = static_cast<int>(i / C), c = static_cast<int>(i % C);
    const size_t o0 = static_cast<size_t>(2 * pi) * C + c, o1 = o0 + C;
    const __half s = scale[c], mn = minv[c];
    const __half zero = __float2half_rn(0.f);
    const __half b0 = base ? base[o0] : zero, b1 = base ? base[o1] : zero;
    uint32_t q0, q1;
    if (ENCODE) {
      q0 = int4_code<LEVELS>(base ? __hsub_rn(x[o0], b0) : x[o0], mn, s);
      q1 = int4_code<LEVELS>(base ? __hsub_rn(x[o1], b1) : x[o1], mn, s);
      packed[i] = static_cast<uint8_t>(q0 | (q1 << 4));
    } else {
      const uint32_t byte = packed[i];
      q0 = byte & 0xFu;
      q1 = byte >> 4;
    }
    if (out != nullptr) {
      const __half v0 = int4_value(q0, mn, s), v1 = int4_value(q1, mn, s);
      out[o0] = base ? __hadd_rn(b0, v0) : v0;
      out[o1] = base ? __hadd_rn(b1, v1) : v1;
    }
  }
}

// ---- LOW_RANK_Q wire packing (slowpath.py:62-75): int4 per column of U (N, r) and per column of V^T --------
// The per-call path runs quantize_int4(U), V.t().contiguous(), quantize_int4(V^T) and six copies into the payload
// (~14 launches per tensor).  V^T is never formed here: column k of V^T is row k of V (r, C), and the row pairs
// (2i, 2i + 1) of V^T that share a byte are the column pairs of V.  Same finalize / code arithmetic as above.
__global__ void __launch_bounds__(256) k_lrq_minmax(const __half* __restrict__ U, const __half* __restrict__ V, int N, int C,
                                                    int r, __half* __restrict__ sU, __half* __restrict__ mU,
                                                    __half* __restrict__ sV, __half* __restrict__ mV) {
  __shared__ __half smn[256], smx[256];
  const int b = blockIdx.x, t = threadIdx.x;
  const bool is_u = b < r;
  const int k = is_u ? b : b - r;
  __half mn = __ushort_as_half(0x7C00), mx = __ushort_as_half(0xFC00);
  if (is_u) {
    for (int n = t; n < N; n += 256) {
      const __half d = U[static_cast<size_t>(n) * r + k];
      mn = __hmin(mn, d);
      mx = __hmax(mx, d);
    }
  } else {
    for (int c = t; c < C; c += 256) {
      const __half d = V[static_cast<size_t>(k) * C + c];
      mn = __hmin(mn, d);
      mx = __hmax(mx, d);
    }
  }
  smn[t] = mn; smx[t] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) { smn[t] = __hmin(smn[t], smn[t + o]); smx[t] = __hmax(smx[t], smx[t + o]); }
    __syncthreads();
  }
  if (t == 0) finalize_column<MODE_INT4>(smn[0], smx[0], k, is_u ? sU : sV, is_u ? mU : mV, nullptr);
}

__global__ void __launch_bounds__(256) k_lrq_encode(const __half* __restrict__ U, const __half* __restrict__ V, int N, int C,
                                                    int r, const __half* __restrict__ sU, const __half* __restrict__ mU,
                                                    const __half* __restrict__ sV, const __half* __restrict__ mV,
                                                    uint8_t* __restrict__ qU, uint8_t* __restrict__ qVt) {
  const size_t nu = static_cast<size_t>(N / 2) * r, nv = static_cast<size_t>(C / 2) * r;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < nu + nv;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    if (i < nu) {
      const int pi = static_cast<int>(i / r), k = static_cast<int>(i % r);
      const __half s = sU[k], mn = mU[k];
      const uint32_t q0 = int4_code<15>(U[static_cast<size_t>(2 * pi) * r + k], mn, s);
      const uint32_t q1 = int4_code<15>(U[static_cast<size_t>(2 * pi + 1) * r + k], mn, s);
      qU[i] = static_cast<uint8_t>(q0 | (q1 << 4));
    } else {
      // consecutive threads take consecutive column pairs of V (coalesced 4-byte reads); the byte lands at
      // qV^T[cp][k]
      const size_t j = i - nu;
      const int k = static_cast<int>(j / (C / 2)), cp = static_cast<int>(j % (C / 2));
      const __half s = sV[k], mn = mV[k];
      const uint32_t d = *reinterpret_cast<const uint32_t*>(V + static_cast<size_t>(k) * C + 2 * cp);
      const uint32_t q0 = int4_code<15>(lo_h(d), mn, s), q1 = int4_code<15>(hi_h(d), mn, s);
      qVt[static_cast<size_t>(cp) * r + k] = static_cast<uint8_t>(q0 | (q1 << 4));
    }
  }
}

template <bool ENCODE>
__global__ void __launch_bounds__(256) k_int8_codec_generic(const __half* __restrict__ x, const __half* __restrict__ base,
                                                            const __half* __restrict__ scale, const int16_t* __restrict__ zpv,
                                                            int8_t* __restrict__ qout, __half* __restrict__ out, int N, int C) {
  const size_t total = static_cast<size_t>(N) * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const __half s = scale[c], zp = __short2half_rn(zpv[c]);
    const __half b = base ? base[i] : __float2half_rn(0.f);
    int q;
    if (ENCODE) {
      q = int8_code(base ? __hsub_rn(x[i], b) : x[i], s, zp);
      qout[i] = static_cast<int8_t>(q);
    } else {
      q = qout[i];
    }
    if (out != nullptr) {
      const __half v = int8_value(q, s, zp);
      out[i] = base ? __hadd_rn(b, v) : v;
    }
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct MinMaxPlan {
  RowGeom geom;
  int B, rows_per_cta;
  size_t part_bytes, total_bytes, smem_bytes;
};
static MinMaxPlan make_minmax_plan(int64_t N, int64_t C) {
  MinMaxPlan pl;
  pl.geom = make_row_geom(C);
  const int threads = pl.geom.TX * pl.geom.TY;
  const int ctas_per_sm = threads >= 512 ? 2 : (1024 / threads);
  int B = sm_count() * ctas_per_sm;
  const int64_t max_b = (N + pl.geom.TY - 1) / pl.geom.TY;
  if (B > max_b) B = static_cast<int>(max_b);
  pl.rows_per_cta = static_cast<int>((N + B - 1) / B);
  pl.B = static_cast<int>((N + pl.rows_per_cta - 1) / pl.rows_per_cta);
  pl.part_bytes = round_up(static_cast<size_t>(pl.B) * C * 2, 256);
  pl.total_bytes = 2 * pl.part_bytes + round_up(static_cast<size_t>(C) * 2, 256);
  pl.smem_bytes = pl.geom.TY > 1 ? static_cast<size_t>(pl.geom.TY) * pl.geom.G * pl.geom.TX * 32 : 0;
  return pl;
}
size_t minmax_codec_workspace_bytes(int64_t N, int64_t C) {
  if (C % 8 != 0) return 256;  // generic path: no scratch
  return make_minmax_plan(N, C).total_bytes;
}

}  // namespace cf

#include "cf_minmax_tma.cuh"

namespace cf {

static int grid_rows(const RowGeom& g, int64_t rows) {
  const int threads = g.TX * g.TY;
  const int ctas_per_sm = threads >= 512 ? 2 : (1024 / threads);
  int64_t bx = static_cast<int64_t>(sm_count()) * ctas_per_sm * 2;
  const int64_t max_b = (rows + g.TY - 1) / g.TY;
  if (bx > max_b) bx = max_b;
  if (bx < 1) bx = 1;
  return static_cast<int>(bx);
}

static int check_mm_shape(int64_t N, int64_t C, bool need_even) {
  CF_CHECK_ARG(N >= 1 && N < (int64_t(1) << 31), "N=%lld out of range", (long long)N);
  CF_CHECK_ARG(!need_even || N % 2 == 0, "INT4 needs an even N, got %lld", (long long)N);
  CF_CHECK_ARG(C >= 1 && C <= 32768, "C=%lld out of range [1, 32768]", (long long)C);
  return CF_OK;
}

static int generic_grid(size_t total) {
  size_t blocks = (total + 255) / 256;
  const size_t cap = static_cast<size_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

template <int MODE>
static int minmax_compress_generic(const __half* xh, const __half* bh, void* new_base, void* codes, void* scale,
                                   void* second, int n, int c, cudaStream_t st) {
  k_minmax_column_generic<MODE><<<c, 256, 0, st>>>(xh, bh, n, c, static_cast<__half*>(scale), second);
  CF_CHECK_LAUNCH();
  if (MODE != MODE_INT8)
    k_int4_codec_generic<true, MmLevels<MODE>::value><<<generic_grid(static_cast<size_t>(n / 2) * c), 256, 0, st>>>(
        xh, bh, static_cast<const __half*>(scale), static_cast<const __half*>(second), static_cast<uint8_t*>(codes),
        static_cast<__half*>(new_base), n, c);
  else
    k_int8_codec_generic<true><<<generic_grid(static_cast<size_t>(n) * c), 256, 0, st>>>(
        xh, bh, static_cast<const __half*>(scale), static_cast<const int16_t*>(second), static_cast<int8_t*>(codes),
        static_cast<__half*>(new_base), n, c);
  CF_CHECK_LAUNCH();
  return CF_OK;
}

template <int MODE>
static int minmax_compress(const void* x, const void* base, void* new_base, void* codes, void* scale,
                           void* second, int64_t N, int64_t C, void* workspace, size_t workspace_bytes,
                           cudaStream_t st) {
  if (int rc = check_mm_shape(N, C, MODE != MODE_INT8)) return rc;
  CF_CHECK_ARG(x && codes && scale && second, "null pointer");
  CF_CHECK_ARG(aligned2(scale) && aligned2(second), "scale vectors must be 2-byte aligned");
  if (C % 8 != 0)
    return minmax_compress_generic<MODE>(static_cast<const __half*>(x), static_cast<const __half*>(base), new_base,
                                         codes, scale, second, static_cast<int>(N), static_cast<int>(C), st);
  CF_CHECK_ARG(aligned16(x) && (!base || aligned16(base)) && (!new_base || aligned16(new_base)),
               "x/base/new_base must be 16-byte aligned");
  CF_CHECK_ARG((reinterpret_cast<uintptr_t>(codes) & 7u) == 0, "codes must be 8-byte aligned");
  MinMaxPlan pl = make_minmax_plan(N, C);
  CF_CHECK_ARG(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0,
               "workspace must be non-null and 256-byte aligned");
  if (pl.total_bytes > workspace_bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", pl.total_bytes, workspace_bytes);
    return CF_ERR_WORKSPACE;
  }
  __half* pmin = static_cast<__half*>(workspace);
  __half* pmax = reinterpret_cast<__half*>(static_cast<char*>(workspace) + pl.part_bytes);
  __half* min_ws = reinterpret_cast<__half*>(static_cast<char*>(workspace) + 2 * pl.part_bytes);
  const __half* xh = static_cast<const __half*>(x);
  const __half* bh = static_cast<const __half*>(base);
  const int n = static_cast<int>(N), c = static_cast<int>(C);
  if (MODE != MODE_INT8 && mm_tma_enabled()) {
    // bulk-async pipelined path (cf_minmax_tma.cuh): statistics -> finalize -> encode, chained with PDL
    const MmStatsPlan sp = make_mm_stats_plan(N, C, bh != nullptr);
    const MmPipe cp = make_mm_pipe(C, bh != nullptr ? 2 : 1, 0);
    if (sp.pipe.ok && cp.ok && static_cast<size_t>(sp.B) * C * 2 <= pl.part_bytes) {
      int rc = bh ? launch_mm_stats_tma<true>(sp, xh, bh, pmin, pmax, n, c, st)
                  : launch_mm_stats_tma<false>(sp, xh, bh, pmin, pmax, n, c, st);
      if (rc) return rc;
      CF_CHECK_CUDA(mm_launch(k_minmax_finalize_v8<MODE>, dim3((c + 15) / 16), dim3(256), 0, st, pmin, pmax, sp.B, c,
                              static_cast<__half*>(scale), second, min_ws));
      constexpr int L = MmLevels<MODE>::value;
      rc = bh ? launch_int4_codec_tma<true, true, L>(cp, xh, bh, static_cast<const __half*>(scale),
                                                     static_cast<const __half*>(second), static_cast<uint8_t*>(codes),
                                                     static_cast<__half*>(new_base), n, c, st)
              : launch_int4_codec_tma<true, false, L>(cp, xh, bh, static_cast<const __half*>(scale),
                                                      static_cast<const __half*>(second), static_cast<uint8_t*>(codes),
                                                      static_cast<__half*>(new_base), n, c, st);
      return rc;
    }
  }
  dim3 block(pl.geom.TX, pl.geom.TY);
#define CF_MM_STATS(GG)                                                                         \
  case GG:                                                                                      \
    if (pl.smem_bytes > 48 * 1024)                                                              \
      CF_CHECK_CUDA(cudaFuncSetAttribute(k_minmax_stats<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         static_cast<int>(pl.smem_bytes)));                     \
    k_minmax_stats<GG><<<pl.B, block, pl.smem_bytes, st>>>(xh, bh, pmin, pmax, n, c, pl.rows_per_cta); \
    break;
  switch (pl.geom.G) {
    CF_MM_STATS(1) CF_MM_STATS(2) CF_MM_STATS(4) CF_MM_STATS(8)
    default: set_error("unsupported geometry"); return CF_ERR_UNSUPPORTED;
  }
#undef CF_MM_STATS
  CF_CHECK_LAUNCH();
  k_minmax_finalize<MODE><<<(c + 31) / 32, 256, 0, st>>>(pmin, pmax, pl.B, c, static_cast<__half*>(scale), second, min_ws);
  CF_CHECK_LAUNCH();
  if (MODE != MODE_INT8) {
    dim3 grid(grid_rows(pl.geom, N / 2));
#define CF_I4(GG)                                                                               \
  case GG:                                                                                      \
    k_int4_codec<GG, true, MmLevels<MODE>::value><<<grid, block, 0, st>>>(xh, bh, static_cast<const __half*>(scale),   \
                                                   static_cast<const __half*>(second),          \
                                                   static_cast<uint8_t*>(codes),                \
                                                   static_cast<__half*>(new_base), n, c);       \
    break;
    switch (pl.geom.G) { CF_I4(1) CF_I4(2) CF_I4(4) CF_I4(8) }
#undef CF_I4
  } else {
    dim3 grid(grid_rows(pl.geom, N));
#define CF_I8(GG)                                                                               \
  case GG:                                                                                      \
    k_int8_codec<GG, true><<<grid, block, 0, st>>>(xh, bh, static_cast<const __half*>(scale),   \
                                                   static_cast<const int16_t*>(second),         \
                                                   static_cast<int8_t*>(codes),                 \
                                                   static_cast<__half*>(new_base), n, c);       \
    break;
    switch (pl.geom.G) { CF_I8(1) CF_I8(2) CF_I8(4) CF_I8(8) }
#undef CF_I8
  }
  CF_CHECK_LAUNCH();
  return CF_OK;
}

template <int MODE>
static int minmax_decompress(const void* codes, const void* scale, const void* second, const void* base,
                             void* recon, int64_t N, int64_t C, cudaStream_t st) {
  if (int rc = check_mm_shape(N, C, MODE == MODE_INT4)) return rc;
  CF_CHECK_ARG(codes && scale && second && recon, "null pointer");
  if (C % 8 != 0 || (reinterpret_cast<uintptr_t>(codes) & 7u) != 0) {
    const int n = static_cast<int>(N), c = static_cast<int>(C);
    if (MODE == MODE_INT4)
      k_int4_codec_generic<false><<<generic_grid(static_cast<size_t>(n / 2) * c), 256, 0, st>>>(
          nullptr, static_cast<const __half*>(base), static_cast<const __half*>(scale),
          static_cast<const __half*>(second), const_cast<uint8_t*>(static_cast<const uint8_t*>(codes)),
          static_cast<__half*>(recon), n, c);
    else
      k_int8_codec_generic<false><<<generic_grid(static_cast<size_t>(n) * c), 256, 0, st>>>(
          nullptr, static_cast<const __half*>(base), static_cast<const __half*>(scale),
          static_cast<const int16_t*>(second), const_cast<int8_t*>(static_cast<const int8_t*>(codes)),
          static_cast<__half*>(recon), n, c);
    CF_CHECK_LAUNCH();
    return CF_OK;
  }
  CF_CHECK_ARG(aligned16(recon) && (!base || aligned16(base)), "base/recon must be 16-byte aligned");
  const RowGeom g = make_row_geom(C);
  dim3 block(g.TX, g.TY);
  const int n = static_cast<int>(N), c = static_cast<int>(C);
  const __half* bh = static_cast<const __half*>(base);
  if (MODE == MODE_INT4 && mm_tma_enabled() && aligned16(codes) && C % 16 == 0 && aligned2(scale) && aligned2(second)) {
    const MmPipe dp = make_mm_pipe(C, bh != nullptr ? 1 : 0, static_cast<int>(C));
    if (dp.ok && (bh != nullptr || dp.stage_bytes > 0)) {
      uint8_t* pk = const_cast<uint8_t*>(static_cast<const uint8_t*>(codes));
      return bh ? launch_int4_codec_tma<false, true, 15>(dp, nullptr, bh, static_cast<const __half*>(scale),
                                                         static_cast<const __half*>(second), pk,
                                                         static_cast<__half*>(recon), n, c, st)
                : launch_int4_codec_tma<false, false, 15>(dp, nullptr, bh, static_cast<const __half*>(scale),
                                                          static_cast<const __half*>(second), pk,
                                                          static_cast<__half*>(recon), n, c, st);
    }
  }
  if (MODE == MODE_INT4) {
    dim3 grid(grid_rows(g, N / 2));
#define CF_I4D(GG)                                                                              \
  case GG:                                                                                      \
    k_int4_codec<GG, false><<<grid, block, 0, st>>>(nullptr, bh, static_cast<const __half*>(scale), \
                                                    static_cast<const __half*>(second),         \
                                                    const_cast<uint8_t*>(static_cast<const uint8_t*>(codes)), \
                                                    static_cast<__half*>(recon), n, c);         \
    break;
    switch (g.G) { CF_I4D(1) CF_I4D(2) CF_I4D(4) CF_I4D(8) }
#undef CF_I4D
  } else {
    dim3 grid(grid_rows(g, N));
#define CF_I8D(GG)                                                                              \
  case GG:                                                                                      \
    k_int8_codec<GG, false><<<grid, block, 0, st>>>(nullptr, bh, static_cast<const __half*>(scale), \
                                                    static_cast<const int16_t*>(second),        \
                                                    const_cast<int8_t*>(static_cast<const int8_t*>(codes)), \
                                                    static_cast<__half*>(recon), n, c);         \
    break;
    switch (g.G) { CF_I8D(1) CF_I8D(2) CF_I8D(4) CF_I8D(8) }
#undef CF_I8D
  }
  CF_CHECK_LAUNCH();
  return CF_OK;
}

}  // namespace cf

extern "C" {
int cf_int4_compress(const void* x, const void* base, void* new_base, void* packed, void* scale, void* minv,
                     int64_t N, int64_t C, void* workspace, size_t workspace_bytes, cf_stream_t stream) {
  return cf::minmax_compress<cf::MODE_INT4>(x, base, new_base, packed, scale, minv, N, C, workspace,
                                            workspace_bytes, static_cast<cudaStream_t>(stream));
}
int cf_int2mm_compress(const void* x, const void* base, void* new_base, void* packed, void* scale, void* minv,
                       int64_t N, int64_t C, void* workspace, size_t workspace_bytes, cf_stream_t stream) {
  return cf::minmax_compress<cf::MODE_INT2MM>(x, base, new_base, packed, scale, minv, N, C, workspace,
                                              workspace_bytes, static_cast<cudaStream_t>(stream));
}
int cf_lowrank_q_pack(const void* U, const void* V, void* payload, int64_t N, int64_t C, int rank, cf_stream_t stream) {
  using namespace cf;
  CF_CHECK_ARG(U && V && payload, "null pointer");
  CF_CHECK_ARG(rank >= 1 && rank <= 64, "rank %d out of range [1, 64]", rank);
  CF_CHECK_ARG(N >= 2 && N % 2 == 0 && C >= 2 && C % 2 == 0, "N and C must be even");
  CF_CHECK_ARG((N / 2 * rank) % 2 == 0 && (C / 2 * rank) % 2 == 0, "N * rank and C * rank must be multiples of 4 (fp16 payload)");
  CF_CHECK_ARG(reinterpret_cast<uintptr_t>(V) % 4 == 0 && reinterpret_cast<uintptr_t>(payload) % 2 == 0, "V must be 4-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* p = static_cast<uint8_t*>(payload);
  const size_t qu_bytes = static_cast<size_t>(N / 2) * rank, qv_bytes = static_cast<size_t>(C / 2) * rank;
  uint8_t* qU = p;
  __half* sU = reinterpret_cast<__half*>(p + qu_bytes);
  __half* mU = sU + rank;
  uint8_t* qVt = reinterpret_cast<uint8_t*>(mU + rank);
  __half* sV = reinterpret_cast<__half*>(qVt + qv_bytes);
  __half* mV = sV + rank;
  const __half* u = static_cast<const __half*>(U);
  const __half* v = static_cast<const __half*>(V);
  k_lrq_minmax<<<2 * rank, 256, 0, st>>>(u, v, static_cast<int>(N), static_cast<int>(C), rank, sU, mU, sV, mV);
  CF_CHECK_LAUNCH();
  k_lrq_encode<<<generic_grid(qu_bytes + qv_bytes), 256, 0, st>>>(u, v, static_cast<int>(N), static_cast<int>(C), rank, sU, mU, sV,
                                                                 mV, qU, qVt);
  CF_CHECK_LAUNCH();
  return CF_OK;
}

int cf_int4_decompress(const void* packed, const void* scale, const void* minv, const void* base, void* recon,
                       int64_t N, int64_t C, cf_stream_t stream) {
  return cf::minmax_decompress<cf::MODE_INT4>(packed, scale, minv, base, recon, N, C,
                                              static_cast<cudaStream_t>(stream));
}
int cf_int8_compress(const void* x, const void* base, void* new_base, void* q, void* scale, void* zero_point,
                     int64_t N, int64_t C, void* workspace, size_t workspace_bytes, cf_stream_t stream) {
  return cf::minmax_compress<cf::MODE_INT8>(x, base, new_base, q, scale, zero_point, N, C, workspace,
                                            workspace_bytes, static_cast<cudaStream_t>(stream));
}
int cf_int8_decompress(const void* q, const void* scale, const void* zero_point, const void* base, void* recon,
                       int64_t N, int64_t C, cf_stream_t stream) {
  return cf::minmax_decompress<cf::MODE_INT8>(q, scale, zero_point, base, recon, N, C,
                                              static_cast<cudaStream_t>(stream));
}
}
