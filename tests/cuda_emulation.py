"""CPU execution of CUDA kernel SOURCE for the tests (TEST INFRASTRUCTURE ONLY).

`SHIM_HEAD` is a C++ prelude that gives kernel source cut out of the .cu files its host-side meaning: one OS
thread per CUDA thread (2-D blocks; the CTAs of a launch run one after another), pthread barriers for
__syncthreads, a per-warp exchange buffer for __shfl_xor_sync, atomics, no-op PDL / fences, and software fp16
for the half-precision intrinsics the kernels use -- every op is the exact fp32 operation followed by one
round-to-nearest-even, which is what the hardware instructions compute.  Faithful for kernels whose warps
shuffle with a full mask in uniform control flow; says nothing about performance, TMA / mbarrier staging or
inter-CTA memory ordering.
"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "compactfusion_b200", "csrc")

SHIM_HEAD = r'''
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

#ifndef EMU_LAUNCH_HOOK
#define EMU_LAUNCH_HOOK
#endif
template <class F> static void launch(unsigned gx, unsigned gy, unsigned bx, unsigned by, F body) {
  // one set of OS threads per launch; they walk the CTAs together (a barrier between CTAs: shared memory and
  // the __shared__ statics belong to one CTA at a time)
  gridDim.x = gx; gridDim.y = gy; blockDim.x = bx; blockDim.y = by;
  const unsigned nthreads = bx * by;
  static pthread_barrier_t between;
  pthread_barrier_init(&cta_bar, nullptr, nthreads);
  pthread_barrier_init(&between, nullptr, nthreads);
  for (unsigned w = 0; w < (nthreads + 31) / 32; ++w) pthread_barrier_init(&warp_bar[w], nullptr, 32);
  EMU_LAUNCH_HOOK
  std::vector<std::thread> ts;
  for (unsigned ty = 0; ty < by; ++ty)
    for (unsigned tx = 0; tx < bx; ++tx)
      ts.emplace_back([=] {
        threadIdx.x = tx; threadIdx.y = ty;
        for (unsigned cy = 0; cy < gy; ++cy)
          for (unsigned cx = 0; cx < gx; ++cx) {
            blockIdx.x = cx; blockIdx.y = cy;
            body();
            pthread_barrier_wait(&between);
          }
      });
  for (auto& t : ts) t.join();
}

// ---- more fp16 / fp32 intrinsics (min/max codecs) ----
static inline __half __ushort_as_half(unsigned short u) { return __half{(uint16_t)u}; }
static inline __half __hsub_rn(__half a, __half b) { return __float2half_rn(h2f(a.v) - h2f(b.v)); }
static inline __half __hmul_rn(__half a, __half b) { return __float2half_rn(h2f(a.v) * h2f(b.v)); }
static inline bool emu_isnan(__half a) { return (a.v & 0x7FFFu) > 0x7C00u; }
static inline __half __hmin(__half a, __half b) {   // NaN: the other operand; -0 < +0
  if (emu_isnan(a)) return b;
  if (emu_isnan(b)) return a;
  const float fa = h2f(a.v), fb = h2f(b.v);
  if (fa == fb) return (a.v & 0x8000u) ? a : b;
  return fa < fb ? a : b;
}
static inline __half __hmax(__half a, __half b) {
  if (emu_isnan(a)) return b;
  if (emu_isnan(b)) return a;
  const float fa = h2f(a.v), fb = h2f(b.v);
  if (fa == fb) return (a.v & 0x8000u) ? b : a;
  return fa > fb ? a : b;
}
static inline __half2 __hmin2(__half2 a, __half2 b) { return __half2{__hmin(a.x, b.x), __hmin(a.y, b.y)}; }
static inline __half2 __hmax2(__half2 a, __half2 b) { return __half2{__hmax(a.x, b.x), __hmax(a.y, b.y)}; }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline __half2 __floats2half2_rn(float a, float b) { return __half2{__float2half_rn(a), __float2half_rn(b)}; }
static inline __half2 __halves2half2(__half a, __half b) { return __half2{a, b}; }
static inline __half __short2half_rn(short v) { return __float2half_rn((float)v); }
static inline __half __ushort2half_rn(unsigned short v) { return __float2half_rn((float)v); }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
'''


PIPE_SHIM = r'''
// ---- stand-ins for csrc/cf_pipe.cuh (all inline PTX there): mbarriers, bulk copies, shared-window loads ----
#include <atomic>
#include <mutex>
#include <sched.h>
static unsigned char* const emu_smem_bytes = reinterpret_cast<unsigned char*>(emu_smem);
static pthread_barrier_t compute_bar;
static int emu_ncompute = 0;   // set by the runner before a launch of a pipelined kernel
struct EmuMbar { std::mutex m; uint32_t init = 0, pending = 0; int64_t tx = 0; uint32_t phase = 0; };
static EmuMbar emu_mbars[64];
static inline EmuMbar& mbar_of(const uint64_t* bar) {
  // barriers live in the shared-memory image; the side table is indexed by their slot in it
  const size_t off = reinterpret_cast<const unsigned char*>(bar) - emu_smem_bytes;
  return emu_mbars[(off / 8) % 64];
}
static inline void mbar_complete_locked(EmuMbar& b) {
  if (b.pending == 0 && b.tx == 0) { b.phase += 1; b.pending = b.init; }
}
static inline uint32_t smem_addr(const void* p) { return (uint32_t)(static_cast<const unsigned char*>(p) - emu_smem_bytes); }
static inline void mbar_init(uint64_t* bar, uint32_t count) {
  EmuMbar& b = mbar_of(bar); std::lock_guard<std::mutex> g(b.m); b.init = b.pending = count; b.tx = 0; b.phase = 0;
}
static inline void mbar_fence_init() {}
static inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  EmuMbar& b = mbar_of(bar); std::lock_guard<std::mutex> g(b.m); b.tx += bytes; b.pending -= 1; mbar_complete_locked(b);
}
static inline void mbar_arrive(uint64_t* bar) {
  EmuMbar& b = mbar_of(bar); std::lock_guard<std::mutex> g(b.m); b.pending -= 1; mbar_complete_locked(b);
}
static inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  EmuMbar& b = mbar_of(bar);
  for (;;) {
    { std::lock_guard<std::mutex> g(b.m); if ((b.phase & 1u) != parity) return; }
    sched_yield();
  }
}
static inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  // the rules of cp.async.bulk: 16-byte aligned addresses, size a multiple of 16
  if ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | bytes) & 15u) { fprintf(stderr, "misaligned bulk copy\n"); abort(); }
  memcpy(dst, src, bytes);
  EmuMbar& b = mbar_of(bar); std::lock_guard<std::mutex> g(b.m); b.tx -= bytes; mbar_complete_locked(b);
}
static inline void bulk_g2s_hint(void* d, const void* s, uint32_t n, uint64_t* bar, uint64_t) { bulk_g2s(d, s, n, bar); }
static inline void bulk_g2s_pol(void* d, const void* s, uint32_t n, uint64_t* bar, uint64_t) { bulk_g2s(d, s, n, bar); }
static inline void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  if ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | bytes) & 15u) { fprintf(stderr, "misaligned bulk store\n"); abort(); }
  memcpy(dst, src, bytes);
}
static inline void bulk_commit() {}
static inline void bulk_wait_read0() {}
static inline void bulk_wait_all0() {}
static inline void fence_async_smem() {}
static inline void fence_async_all() {}
static inline uint64_t make_policy_evict_first() { return 1; }
static inline uint64_t make_policy_evict_last() { return 2; }
static inline void compute_sync(int) { pthread_barrier_wait(&compute_bar); }
static inline void __syncwarp() {
  const unsigned lin = threadIdx.y * blockDim.x + threadIdx.x;
  pthread_barrier_wait(&warp_bar[lin >> 5]);
}
static inline uint4 lds128a(uint32_t a) { uint4 r; memcpy(&r, emu_smem_bytes + a, 16); return r; }
static inline uint4 lds128(const void* p) { uint4 r; memcpy(&r, p, 16); return r; }
static inline uint32_t lds8a(uint32_t a) { return emu_smem_bytes[a]; }
static inline uint32_t lds16a(uint32_t a) { uint16_t r; memcpy(&r, emu_smem_bytes + a, 2); return r; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static inline long long clock64() { return 0; }
static inline void __nanosleep(unsigned) { sched_yield(); }
static inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
'''

SLURP = r'''
static std::vector<unsigned char> slurp(const char* path) {
  FILE* f = fopen(path, "rb"); if (!f) { perror(path); exit(2); }
  fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<unsigned char> v(n); if (fread(v.data(), 1, n, f) != (size_t)n) exit(2); fclose(f); return v;
}
'''


def strip_asm(text):
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


def common_source():
    """RowGeom / make_row_geom and the device helpers of cf_common.cuh (minus the three PTX load / store
    helpers, which the shim provides), wrapped in namespace cf."""
    common = open(os.path.join(CSRC, "cf_common.cuh")).read()
    geom = re.search(r"(// Geometry shared by all row-streaming kernels.*?return g;\n\})", common, flags=re.S).group(1)
    dev = re.search(r"#ifdef __CUDACC__\n(.*?)#endif  // __CUDACC__", common, flags=re.S).group(1)
    dev = re.sub(r"(// [^\n]*\n)*__device__ __forceinline__ (uint4|void) (ldg_stream|stg_stream|stg_stream_pol)\(.*?\n\}\n",
                 "", dev, flags=re.S)
    assert "asm" not in dev, "an inline-PTX helper of cf_common.cuh is not covered by the shim"
    return "namespace cf {\n" + geom + "\n" + dev + "}  // namespace cf\n"


def build(directory, source_text, name="emu"):
    cpp = directory / f"{name}.cpp"
    cpp.write_text(source_text)
    exe = str(directory / name)
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-pthread", "-w", "-I", "/usr/local/cuda/include", "-I",
                        os.path.join(ROOT, "include"), str(cpp), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return exe
