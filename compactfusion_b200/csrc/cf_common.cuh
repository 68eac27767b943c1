// Shared device/host helpers for libcompactb200 (sm_100a only).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/compactb200.h"

namespace cf {

// ---------------------------------------------------------------------------------------
// host side: error reporting, device info
// ---------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int sm_count();

#define CF_CHECK_ARG(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      ::cf::set_error(__VA_ARGS__);      \
      return CF_ERR_ARG;                 \
    }                                    \
  } while (0)

#define CF_CHECK_CUDA(expr)                                                         \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ::cf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                      __FILE__, __LINE__);                                          \
      return CF_ERR_CUDA;                                                           \
    }                                                                               \
  } while (0)

#define CF_CHECK_LAUNCH() CF_CHECK_CUDA(cudaGetLastError())

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline bool aligned2(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 1u) == 0; }
static inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Geometry shared by all row-streaming kernels: a CTA is TX x TY threads; thread (tx, ty)
// owns the 16-byte column groups g = tx + j*TX (j < G) of the rows it visits, so per-column
// state (scale fragments, column accumulators) lives in registers for the whole kernel.
struct RowGeom {
  int TX, TY, G;  // G in {1,2,4,8}
};
static inline RowGeom make_row_geom(int64_t C, int target_threads = 512) {
  RowGeom g;
  const int groups = static_cast<int>(C / 8);
  int G = 1;
  while ((groups + G - 1) / G > target_threads && G < 8) G *= 2;
  int per = (groups + G - 1) / G;
  g.TX = (per + 31) / 32 * 32;
  if (g.TX < 32) g.TX = 32;
  g.G = G;
  g.TY = target_threads / g.TX;
  if (g.TY < 1) g.TY = 1;
  return g;
}

// ---------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__

// 128-bit streaming load: read-only path, do not allocate in L1 (each byte is used once).
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// same with an L2 eviction-priority policy (createpolicy); policy == 0: plain streaming store
__device__ __forceinline__ void stg_stream_pol(void* p, const uint4& v, uint64_t policy) {
  if (policy != 0)
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy)
                 : "memory");
  else
    stg_stream(p, v);
}

__device__ __forceinline__ __half2 u2h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t h22u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 8 consecutive fp16 values as 4 packed half2 words.
struct H8 {
  uint32_t w[4];
};
__device__ __forceinline__ H8 as_h8(const uint4& v) {
  H8 r;
  r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
  return r;
}
__device__ __forceinline__ uint4 as_u4(const H8& v) { return make_uint4(v.w[0], v.w[1], v.w[2], v.w[3]); }

// delta = x - base, one fp16 rounding per element (fastpath.py:58).
__device__ __forceinline__ H8 h8_sub(const H8& a, const H8& b) {
  H8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.w[i] = h22u(__hsub2_rn(u2h2(a.w[i]), u2h2(b.w[i])));
  return r;
}

// Bit i (i = 0..7) of the result is 1 iff element i satisfies (v >= 0); NaN -> 0, -0 -> 1.
__device__ __forceinline__ uint32_t h8_ge0_bits(const H8& v) {
  const __half2 z = __float2half2_rn(0.f);
  uint32_t bits = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t m = __hge2_mask(u2h2(v.w[i]), z);  // 0xFFFF per true half
    bits |= ((m & 1u) | ((m >> 15) & 2u)) << (2 * i);
  }
  return bits;
}

// Sum of |v| over the 8 elements in index order, also accumulated per column.
__device__ __forceinline__ float h8_abs_accumulate(const H8& v, float* colacc) {
  float rs = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(__habs2(u2h2(v.w[i])));
    colacc[2 * i] += f.x;
    colacc[2 * i + 1] += f.y;
    rs += f.x;
    rs += f.y;
  }
  return rs;
}

#endif  // __CUDACC__

}  // namespace cf
