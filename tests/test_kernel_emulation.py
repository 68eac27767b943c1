"""Execute the SOURCE of two small CUDA kernels on the CPU (TEST ONLY) -- k_error_stats and k_lse_merge of
csrc/cf_consumer.cu, which landed after the round's GPU minutes were spent -- so that their indexing, barrier
/ shuffle / last-ticket logic and arithmetic have run at least once before they meet a GPU.

How: the kernel text is cut out of the .cu file and compiled with g++ against a ~100-line shim that gives
threadIdx / blockIdx, __syncthreads (pthread barrier over the CTA's threads), __shfl_xor_sync (per-warp
exchange buffer between two warp barriers), atomicAdd and the fp16 helpers their host-side meaning; every
CUDA thread is an OS thread, CTAs run one after another.  That is a faithful execution for kernels whose
warps shuffle with a full mask in uniform control flow, which these two are.  It says nothing about
performance or about memory-model subtleties between CTAs (the emulated CTAs are sequential).
"""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

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

struct Idx { unsigned x = 0, y = 0, z = 0; };
static thread_local Idx threadIdx, blockIdx;
static Idx blockDim, gridDim;
static pthread_barrier_t cta_bar, warp_bar[32];
static unsigned char shfl_buf[32][32][8];

#undef __global__
#undef __device__
#undef __forceinline__
#undef __launch_bounds__
#undef __shared__
#undef __restrict__
#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __restrict__
static inline void __syncthreads() { pthread_barrier_wait(&cta_bar); }
static inline void __threadfence() { __sync_synchronize(); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int o) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  static_assert(sizeof(T) <= 8, "shuffle payload");
  memcpy(shfl_buf[warp][lane], &v, sizeof(T));
  pthread_barrier_wait(&warp_bar[warp]);
  T r;
  memcpy(&r, shfl_buf[warp][lane ^ o], sizeof(T));
  pthread_barrier_wait(&warp_bar[warp]);
  return r;
}
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }

// fp16 pairs as the kernels see them
struct __half2 { uint32_t u; };
static inline float h2f(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu;
  uint32_t bits;
  if (exp == 0) { float f = (float)man * (1.0f / 16777216.0f); memcpy(&bits, &f, 4); bits |= sign; }
  else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
  else bits = sign | ((exp + 112u) << 23) | (man << 13);
  float out; memcpy(&out, &bits, 4); return out;
}
static inline __half2 u2h2(uint32_t u) { return __half2{u}; }
static inline float2 __half22float2(__half2 h) { return make_float2(h2f((uint16_t)(h.u & 0xFFFFu)), h2f((uint16_t)(h.u >> 16))); }
struct H8 { uint32_t w[4]; };
static inline H8 as_h8(const uint4& v) { H8 r; r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w; return r; }
static inline uint4 ldg_stream(const void* p) { return *static_cast<const uint4*>(p); }

template <class F> static void launch(unsigned grid, unsigned block, F body) {
  gridDim.x = grid; blockDim.x = block;
  for (unsigned b = 0; b < grid; ++b) {
    pthread_barrier_init(&cta_bar, nullptr, block);
    for (unsigned w = 0; w < (block + 31) / 32; ++w) pthread_barrier_init(&warp_bar[w], nullptr, 32);
    std::vector<std::thread> ts;
    for (unsigned t = 0; t < block; ++t)
      ts.emplace_back([=] { threadIdx.x = t; blockIdx.x = b; body(); });
    for (auto& t : ts) t.join();
  }
}

// ---- kernel source, verbatim from csrc/cf_consumer.cu ----
KERNEL_SOURCE

static std::vector<unsigned char> slurp(const char* path) {
  FILE* f = fopen(path, "rb"); if (!f) { perror(path); exit(2); }
  fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<unsigned char> v(n); if (fread(v.data(), 1, n, f) != (size_t)n) exit(2); fclose(f); return v;
}

int main(int argc, char** argv) {
  const std::string what = argv[1];
  if (what == "stats") {   // stats <a.bin> <b.bin> <grid> <calls>
    auto a = slurp(argv[2]), b = slurp(argv[3]);
    const unsigned grid = atoi(argv[4]); const int calls = atoi(argv[5]);
    const int64_t n8 = (int64_t)a.size() / 16;
    std::vector<double> partial(4 * 4096, 0.0); unsigned ticket = 0; float out[4];
    for (int c = 0; c < calls; ++c) {   // the workspace is reused: the kernel must leave the ticket at 0
      launch(grid, cf::kStatsThreads, [&] {
        cf::k_error_stats(reinterpret_cast<const uint4*>(a.data()), reinterpret_cast<const uint4*>(b.data()), n8,
                          partial.data(), &ticket, out);
      });
      printf("%.9g %.9g %.9g %.9g %u\n", out[0], out[1], out[2], out[3], ticket);
    }
  } else {                 // merge <out.bin f32> <block_out.bin f16> <lse.bin> <block_lse.bin> B S H D grid
    auto out = slurp(argv[2]), bo = slurp(argv[3]), lse = slurp(argv[4]), bl = slurp(argv[5]);
    const int B = atoi(argv[6]), S = atoi(argv[7]), H = atoi(argv[8]), D = atoi(argv[9]);
    const unsigned grid = atoi(argv[10]);
    std::vector<float> lse_out(lse.size() / 4, -12345.f);
    const int64_t total4 = (int64_t)B * S * H * (D / 4);
    launch(grid, 256, [&] {
      cf::k_lse_merge(reinterpret_cast<float4*>(out.data()), reinterpret_cast<const uint2*>(bo.data()),
                      reinterpret_cast<const float*>(lse.data()), reinterpret_cast<const float*>(bl.data()),
                      lse_out.data(), total4, S, H, D / 4);
    });
    fwrite(out.data(), 1, out.size(), stdout);
    fwrite(lse_out.data(), 4, lse_out.size(), stdout);
  }
  return 0;
}
'''


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    src = open(os.path.join(ROOT, "compactfusion_b200", "csrc", "cf_consumer.cu")).read()
    m = re.search(r"(namespace cf \{.*?)static int stats_grid", src, flags=re.S)
    assert m, "kernel section of cf_consumer.cu not found"
    kernels = m.group(1) + "}  // namespace cf\n"
    assert "k_lse_merge" in kernels and "k_error_stats" in kernels and "<<<" not in kernels
    d = tmp_path_factory.mktemp("emu")
    cpp = d / "emu.cpp"
    cpp.write_text(SHIM.replace("KERNEL_SOURCE", kernels))
    exe = str(d / "emu")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-pthread", "-I", "/usr/local/cuda/include", str(cpp), "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe, d


@pytest.mark.parametrize("numel,grid", [(8, 1), (8 * 300, 1), (8 * 1000, 3), (8 * 4097, 7)])
def test_k_error_stats_source_on_cpu(emulator, numel, grid):
    exe, d = emulator
    rng = np.random.default_rng(numel)
    b = rng.standard_normal(numel).astype(np.float16)
    a = (b.astype(np.float32) + 0.05 * rng.standard_normal(numel).astype(np.float32)).astype(np.float16)
    (d / "a.bin").write_bytes(a.tobytes())
    (d / "b.bin").write_bytes(b.tobytes())
    r = subprocess.run([exe, "stats", str(d / "a.bin"), str(d / "b.bin"), str(grid), "2"], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    rows = [ln.split() for ln in r.stdout.strip().splitlines()]
    assert len(rows) == 2 and rows[0] == rows[1], "second call on the same workspace must give the same result"
    sse, ssr, me, mr, ticket = (float(v) for v in rows[0])
    diff = a.astype(np.float64) - b.astype(np.float64)
    assert ticket == 0
    assert abs(sse - float((diff * diff).sum())) <= 2e-6 * float((diff * diff).sum())
    assert abs(ssr - float((b.astype(np.float64) ** 2).sum())) <= 2e-6 * float((b.astype(np.float64) ** 2).sum())
    assert me == np.float32(np.abs(diff).max()) and mr == np.float32(np.abs(b.astype(np.float64)).max())


@pytest.mark.parametrize("shape,grid", [((1, 5, 3, 8), 1), ((2, 33, 4, 16), 2), ((1, 70, 3, 64), 5)])
def test_k_lse_merge_source_on_cpu(emulator, shape, grid):
    exe, d = emulator
    bsz, s, h, dd = shape
    rng = np.random.default_rng(s)
    out = rng.standard_normal((bsz, s, h, dd)).astype(np.float32)
    block_out = rng.standard_normal((bsz, s, h, dd)).astype(np.float16)
    lse = (3 * rng.standard_normal((bsz, h, s))).astype(np.float32)
    block_lse = (3 * rng.standard_normal((bsz, h, s))).astype(np.float32)
    block_lse[0, 0, 0], block_lse[0, 0, 1] = lse[0, 0, 0] + 90.0, lse[0, 0, 1] - 90.0  # saturated weights
    for name, arr in (("o", out), ("bo", block_out), ("l", lse), ("bl", block_lse)):
        (d / f"{name}.bin").write_bytes(arr.tobytes())
    r = subprocess.run([exe, "merge", str(d / "o.bin"), str(d / "bo.bin"), str(d / "l.bin"), str(d / "bl.bin"),
                        str(bsz), str(s), str(h), str(dd), str(grid)], capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got_out = np.frombuffer(r.stdout[:out.nbytes], dtype=np.float32).reshape(out.shape)
    got_lse = np.frombuffer(r.stdout[out.nbytes:], dtype=np.float32).reshape(lse.shape)
    # update_out_and_lse (attention.py), in float64: lse (b,h,s) broadcasts over out (b,s,h,d) as (b,s,h,1)
    l64, bl64 = lse.astype(np.float64), block_lse.astype(np.float64)
    w = 1.0 / (1.0 + np.exp(-(bl64 - l64)))
    w_bshd = np.transpose(w, (0, 2, 1))[..., None]
    want_out = out.astype(np.float64) - w_bshd * (out.astype(np.float64) - block_out.astype(np.float64))
    want_lse = np.logaddexp(l64, bl64)
    assert np.abs(got_out - want_out).max() < 5e-6
    assert np.abs(got_lse - want_lse).max() < 2e-5
    assert not np.any(got_lse == -12345.0), "every (b, h, s) must be written exactly by the j == 0 thread"
    assert np.array_equal(got_out[0, 0, 0], block_out[0, 0, 0].astype(np.float32)) or \
        np.abs(got_out[0, 0, 0] - block_out[0, 0, 0].astype(np.float32)).max() < 1e-6
    assert np.abs(got_out[0, 1, 0] - out[0, 1, 0]).max() < 1e-6
