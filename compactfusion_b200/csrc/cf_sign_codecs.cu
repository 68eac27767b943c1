// BINARY (1-bit) and INT2 residual codecs for sm_100a.
//
// Reference semantics: xfuser/compact/fastpath.py (binary_quant_fastpath :124-228,
// _binary_quant_fastpath :13-120, _binary_dequant_fastpath :277-367, int2_quant_fastpath
// :584-669, _int2_quant_fastpath :486-580, _int2_dequant_fastpath :672-741).  The reference
// spends 5-6 eager passes on the scales before its Triton kernel; here one pass over x/base
// produces delta statistics (and, for BINARY, the packed sign bits), a tiny finalize kernel
// turns the partial sums into the fp16 scale vectors, and one streaming pass applies the
// codes to the base (the *same* kernel on sender and receiver, so caches stay bit-identical).
//
// HBM-bound byte work: 128-bit loads, per-thread column ownership (scale fragments and
// column accumulators stay in registers), warp-shuffle row reductions, fixed-order two-level
// fp32 sums (deterministic, no atomics).
#include <stdlib.h>

#include <unordered_map>

#include "cf_common.cuh"
#include "cf_pipe.cuh"

namespace cf {

enum { MODE_BINARY = 0, MODE_INT2 = 1 };

constexpr int kRowChunk = 128;  // rows whose per-warp partial sums are staged in smem at once
// rows in flight per thread: 4 x 128-bit loads per operand when a thread owns one column
// group; fewer when it owns several (keeps the kernels under 128 registers, no spills)
__host__ __device__ constexpr int unroll_for(int G) { return G == 1 ? 4 : (G == 2 ? 2 : 1); }

struct StatsParams {
  const __half* x[CF_MAX_BATCH];
  const __half* base[CF_MAX_BATCH];  // may be null (base = 0)
  uint8_t* packed[CF_MAX_BATCH];     // BINARY only
  __half* rowmean[CF_MAX_BATCH];     // (N) fp16 mean_c |delta|
  float* tokpart[CF_MAX_BATCH];      // (B) partial sums of rowmean
  float* colpart[CF_MAX_BATCH];      // (B, C) partial column sums of |delta|
  int N, C, rows_per_cta;
};

// Fused compress + one-sided exchange (cf_sign_compress_put): instead of a local send buffer, the codec
// kernels store the codes and scales of tensor t straight into `n_dst` receive slots -- local memory or
// peer memory mapped over NVLink -- and the last CTA of the call's last kernel publishes the flags.
struct FanOut {
  unsigned char* dst[CF_MAX_FANOUT];  // [t * n_dst + q]: start of tensor t's payload [codes | U | V] at destination q
  uint32_t* flag[CF_MAX_PEERS];       // per destination: this origin's counter flag
  uint32_t* count;                    // local: puts issued so far on this slot (the value published)
  uint32_t* done;                     // local: CTA ticket counter, reset by the last CTA
  unsigned long long u_off, v_off;    // byte offsets of U (N) and V (C) inside a payload
  int n_dst;
  int publish_mode;                   // 0: every CTA adds 1 to every flag; 1: the last CTA stores the new count
  int vec16;                          // U and V of every destination are 16-byte aligned and N % 8 == 0, C % 32 == 0:
                                      // the finalize kernel stores them 8 scales at a time
};

__device__ __forceinline__ void red_add_relaxed_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Tail of the kernel that completes a fused put.  Every CTA calls it after its last payload store: the
// barrier orders the CTA's stores before thread 0's system-scope fence (cumulative), then thread 0 adds 1
// to this origin's flag at every destination (fence + relaxed RMW = release; an NVLink atomic for the
// peers).  A flag therefore counts CTA arrivals, and *count -- the value the receivers' decompress kernels
// wait for (wait_origins: flag >= count, observed through the RMW chain) -- advances by the grid size per
// put.  Nothing remote sits behind a second fence or a ticket: the critical path of a put is one fence.
// The ticket only keeps the LOCAL count: the CTA that draws the last one adds the grid size to it.
// `ncompute` > 0: only threads [0, ncompute) of the CTA call (named barrier 1, the pipelined kernels'
// compute warps); 0: the whole CTA calls.
__device__ __forceinline__ void fanout_publish_impl(const FanOut& f, unsigned total_ctas, int ncompute) {
  if (ncompute > 0)
    asm volatile("bar.sync 1, %0;" ::"r"(ncompute) : "memory");
  else
    __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (f.publish_mode == 0) {
      for (int q = 0; q < f.n_dst; ++q) red_add_relaxed_sys_u32(f.flag[q], 1u);
      const unsigned ticket = atomicAdd(f.done, 1u);
      if (ticket == total_ctas - 1) {
        *f.count += total_ctas;  // read by kernels launched after this one (stream order / PDL wait)
        *f.done = 0u;
      }
    } else {
      // CF_PUBLISH_MODE=1 (A/B): n_dst flag stores by ONE CTA instead of n_dst remote atomics by every CTA,
      // at the price of a second fence on the critical path (the pattern of k_p2p_put).  Same count units.
      const unsigned ticket = atomicAdd(f.done, 1u);
      if (ticket == total_ctas - 1) {
        __threadfence_system();
        const uint32_t v = *f.count + total_ctas;
        for (int q = 0; q < f.n_dst; ++q) st_relaxed_sys_u32(f.flag[q], v);
        *f.count = v;
        *f.done = 0u;
      }
    }
  }
}
// CF_PUBLISH_MODE=2 (default): the put's kernels do not publish at all.  A one-warp kernel launched right behind
// them (programmatic dependent launch) is ordered after the COMPLETION of those grids -- all their stores, generic
// and bulk-async, local and peer -- so ONE system-scope fence by one thread is cumulative over the whole put, and
// the flags follow as plain relaxed stores.  Measured at W = 2 (profiles/r2_multi_gpu_n2.md): the per-CTA
// fence + remote atomics of modes 0 / 1 cost the finalize kernel 8 us (11.4 us against 3.2 us without them).
// Flags and *count advance by ONE per put, whatever the grids of the ranks look like.
struct PublishParams {
  uint32_t* flag[CF_MAX_PEERS];
  uint32_t* count;
  int n_dst;
  int sc_fence;   // CF_PUBLISH_FENCE=sc: __threadfence_system() (MEMBAR.SC.SYS) instead of fence.acq_rel.sys (MEMBAR.ALL.SYS)
};
__global__ void __launch_bounds__(32) k_publish_flags(const PublishParams p) {
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    // release = fence.acq_rel + relaxed store (PTX memory model); the sequentially consistent fence is not needed
    if (p.sc_fence)
      __threadfence_system();
    else
      asm volatile("fence.acq_rel.sys;" ::: "memory");
    const uint32_t v = *p.count + 1u;
    for (int q = 0; q < p.n_dst; ++q) st_relaxed_sys_u32(p.flag[q], v);
    *p.count = v;  // read by the flag-waiting kernels launched after this one (stream order / PDL wait)
  }
}

__device__ __forceinline__ void fanout_publish(const FanOut& f, unsigned total_ctas) { fanout_publish_impl(f, total_ctas, 0); }
__device__ __forceinline__ void fanout_publish_compute(const FanOut& f, unsigned total_ctas, int ncompute) {
  fanout_publish_impl(f, total_ctas, ncompute);
}

// ---------------------------------------------------------------------------------------
// pass 1: delta statistics (+ sign packing for BINARY)
// grid (B, batch), block (TX, TY)
// ---------------------------------------------------------------------------------------
template <int MODE, int G>
__global__ void __launch_bounds__(512) k_delta_stats(const StatsParams p) {
  extern __shared__ float smem[];
  const int t = blockIdx.y;
  const __half* __restrict__ x = p.x[t];
  const __half* __restrict__ base = p.base[t];
  uint8_t* __restrict__ packed = p.packed[t];
  const int N = p.N, C = p.C;
  const int groups = C >> 3;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int TX = blockDim.x, TY = blockDim.y;
  const int NWX = TX >> 5, warp_x = tx >> 5, lane = tx & 31;
  const int tid = ty * TX + tx, nthreads = TX * TY;
  const int r_begin = blockIdx.x * p.rows_per_cta;
  const int r_end = min(N, r_begin + p.rows_per_cta);

  constexpr int kUnroll = unroll_for(G);
  float colacc[G][8];
#pragma unroll
  for (int j = 0; j < G; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) colacc[j][e] = 0.f;
  float tokacc = 0.f;
  const float inv_c_den = static_cast<float>(C);

  for (int chunk = r_begin; chunk < r_end; chunk += kRowChunk) {
    const int cend = min(r_end, chunk + kRowChunk);
    for (int r = chunk + ty; r < cend; r += TY * kUnroll) {
      uint4 xv[kUnroll][G], bv[kUnroll][G];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int rr = r + u * TY;
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const int g = tx + j * TX;
          xv[u][j] = make_uint4(0, 0, 0, 0);
          bv[u][j] = make_uint4(0, 0, 0, 0);
          if (rr < cend && g < groups) {
            const size_t off = static_cast<size_t>(rr) * C + 8 * g;
            xv[u][j] = ldg_stream(x + off);
            if (base != nullptr) bv[u][j] = ldg_stream(base + off);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int rr = r + u * TY;
        if (rr < cend) {  // warp-uniform: a warp lies inside one ty
          float rs = 0.f;
#pragma unroll
          for (int j = 0; j < G; ++j) {
            const int g = tx + j * TX;
            const H8 d = h8_sub(as_h8(xv[u][j]), as_h8(bv[u][j]));
            if (g < groups) {
              rs += h8_abs_accumulate(d, colacc[j]);
              if (MODE == MODE_BINARY)
                packed[static_cast<size_t>(rr) * groups + g] = static_cast<uint8_t>(h8_ge0_bits(d));
            }
          }
          rs = warp_sum(rs);
          if (lane == 0) smem[(rr - chunk) * NWX + warp_x] = rs;
        }
      }
    }
    __syncthreads();
    // row means of this chunk: fixed-order sum over the row's warps, one rounding to fp16
    for (int i = tid; i < cend - chunk; i += nthreads) {
      float s = 0.f;
      for (int w = 0; w < NWX; ++w) s += smem[i * NWX + w];
      const __half h = __float2half_rn(s / inv_c_den);
      p.rowmean[t][chunk + i] = h;
      tokacc += __half2float(h);
    }
    __syncthreads();
  }

  // ---- CTA partial of sum_n rowmean[n] (fixed order: thread-strided, then warp tree) ----
  {
    float v = warp_sum(tokacc);
    const int wid = tid >> 5, nw = nthreads >> 5;
    if ((tid & 31) == 0) smem[wid] = v;
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int w = 0; w < nw; ++w) s += smem[w];
      p.tokpart[t][blockIdx.x] = s;
    }
    __syncthreads();
  }

  // ---- CTA partial column sums: reduce over ty in order, then one store per column ----
  float* __restrict__ colout = p.colpart[t] + static_cast<size_t>(blockIdx.x) * C;
  if (TY > 1) {
    // smem layout [ty][j][tx][8]
#pragma unroll
    for (int j = 0; j < G; ++j) {
      float4* dst = reinterpret_cast<float4*>(smem + ((static_cast<size_t>(ty) * G + j) * TX + tx) * 8);
      dst[0] = make_float4(colacc[j][0], colacc[j][1], colacc[j][2], colacc[j][3]);
      dst[1] = make_float4(colacc[j][4], colacc[j][5], colacc[j][6], colacc[j][7]);
    }
    __syncthreads();
    if (ty == 0) {
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const int g = tx + j * TX;
        if (g < groups) {
          float acc[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = 0.f;
          for (int yy = 0; yy < TY; ++yy) {
            const float* src = smem + ((static_cast<size_t>(yy) * G + j) * TX + tx) * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += src[e];
          }
          float4* o = reinterpret_cast<float4*>(colout + 8 * g);
          o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
          o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const int g = tx + j * TX;
      if (g < groups) {
        float4* o = reinterpret_cast<float4*>(colout + 8 * g);
        o[0] = make_float4(colacc[j][0], colacc[j][1], colacc[j][2], colacc[j][3]);
        o[1] = make_float4(colacc[j][4], colacc[j][5], colacc[j][6], colacc[j][7]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// finalize: partial sums -> fp16 scale vectors written straight into the payload
//   V[c]  = fp16( sum_b colpart[b][c] / N )                       (fastpath.py:160 / :618)
//   tm    = fp16( sum_n rowmean[n] / N )
//   BINARY: U[n] = fp16( rowmean[n] / tm )                        (fastpath.py:164-165)
//   INT2:   U[n] = fp16( rowmean[n] / fp16(tm + 1e-6) )           (fastpath.py:621-622)
// grid (F, batch), block 256 = 32 columns x 8 partial lanes
// ---------------------------------------------------------------------------------------
struct FinalizeParams {
  const __half* rowmean[CF_MAX_BATCH];
  const float* tokpart[CF_MAX_BATCH];
  const float* colpart[CF_MAX_BATCH];
  __half* scale_u[CF_MAX_BATCH];
  __half* scale_v[CF_MAX_BATCH];
  int N, C, B;
};

template <int MODE, bool PUT>
__global__ void __launch_bounds__(1024) k_finalize_scales(const FinalizeParams p, const FanOut f, const int publish) {
  // block = 32 columns x 32 partial lanes: every thread issues ceil(B/32) independent loads
  __shared__ float red[32][33];
  __shared__ float denom_s;
  const int t = blockIdx.y;
  const int N = p.N, C = p.C, B = p.B;
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
  const float n_f = static_cast<float>(N);
  const float* __restrict__ colpart = p.colpart[t];
  pdl_wait();  // launched as a programmatic dependent of the stats kernel: its partials must be complete
  pdl_launch_dependents();

  // column means (issued first: the long-latency part)
  const int c = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (c < C) {
#pragma unroll 4
    for (int b = py; b < B; b += 32) s += colpart[static_cast<size_t>(b) * C + c];
  }
  red[py][cx] = s;

  // token-mean denominator (every CTA recomputes it in the same order -> identical value)
  if (py == 0) {
    float ts = 0.f;
    for (int b = cx; b < B; b += 32) ts += p.tokpart[t][b];
    ts = warp_sum(ts);
    if (cx == 0) {
      const __half tm = __float2half_rn(ts / n_f);
      float d = __half2float(tm);
      if (MODE == MODE_INT2) d = __half2float(__float2half_rn(d + 1e-6f));
      denom_s = d;
    }
  }
  __syncthreads();
  if (py == 0 && c < C) {
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) tot += red[k][cx];
    const __half v = __float2half_rn(tot / n_f);
    if (PUT && f.vec16) {
      // 8 neighbouring lanes' scales in one 16-byte store per destination (a remote 2-byte store per lane and
      // destination costs an NVLink packet each)
      // (C % 32 == 0 here, so all 32 lanes of this warp are inside the tensor and take part in the shuffles)
      const uint32_t mine = __half_as_ushort(v);
      const uint32_t next = __shfl_xor_sync(0xffffffffu, mine, 1);
      const uint32_t pair = mine | (next << 16);            // valid on even lanes
      uint4 w;
      w.x = pair;
      w.y = __shfl_xor_sync(0xffffffffu, pair, 2);          // lanes 0 / 4 (mod 8): the pair of lanes 2 / 6
      w.z = __shfl_xor_sync(0xffffffffu, pair, 4);          // lane 0 (mod 8): the pair of lane 4
      w.w = __shfl_xor_sync(0xffffffffu, w.y, 4);           // lane 0 (mod 8): the pair of lane 6
      if ((cx & 7) == 0)
        for (int q = 0; q < f.n_dst; ++q)
          *reinterpret_cast<uint4*>(f.dst[t * f.n_dst + q] + f.v_off + 2 * static_cast<size_t>(c)) = w;
    } else if (PUT) {
      for (int q = 0; q < f.n_dst; ++q) reinterpret_cast<__half*>(f.dst[t * f.n_dst + q] + f.v_off)[c] = v;
    } else {
      p.scale_v[t][c] = v;
    }
  }
  const float denom = denom_s;
  if (PUT && f.vec16) {
    for (int n8 = (blockIdx.x * blockDim.x + threadIdx.x) * 8; n8 < N; n8 += gridDim.x * blockDim.x * 8) {
      const uint4 rm = *reinterpret_cast<const uint4*>(p.rowmean[t] + n8);  // workspace: 256-byte aligned
      const __half* rh = reinterpret_cast<const __half*>(&rm);
      __align__(16) __half u8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) u8[e] = __float2half_rn(__half2float(rh[e]) / denom);
      for (int q = 0; q < f.n_dst; ++q)
        *reinterpret_cast<uint4*>(f.dst[t * f.n_dst + q] + f.u_off + 2 * static_cast<size_t>(n8)) =
            *reinterpret_cast<const uint4*>(u8);
    }
  } else {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
      const __half u = __float2half_rn(__half2float(p.rowmean[t][n]) / denom);
      if (PUT) {
        for (int q = 0; q < f.n_dst; ++q) reinterpret_cast<__half*>(f.dst[t * f.n_dst + q] + f.u_off)[n] = u;
      } else {
        p.scale_u[t][n] = u;
      }
    }
  }
  if (PUT && publish) fanout_publish(f, gridDim.x * gridDim.y);
}

// ---------------------------------------------------------------------------------------
// apply: recon = base + dequant(codes, U, V)           grid (B, batch), block (TX, TY)
// also used by the sender for its error-feedback cache update.
// ---------------------------------------------------------------------------------------
struct ApplyParams {
  const uint8_t* packed[CF_MAX_BATCH];
  const __half* scale_u[CF_MAX_BATCH];
  const __half* scale_v[CF_MAX_BATCH];
  const __half* base[CF_MAX_BATCH];  // may be null
  __half* recon[CF_MAX_BATCH];
  // one-sided transport (cf_p2p.cu): payload t is valid once *wait_flag[t] >= *expected
  const uint32_t* wait_flag[CF_MAX_BATCH];  // null entries: no wait
  const uint32_t* expected;                 // null: no waiting at all
  uint32_t* error;                          // set to 1 if a wait timed out
  int wait_mode;                            // 0: acquire loads (default); 1: relaxed polls + system fence
  int N, C, K;
};

__device__ __forceinline__ uint32_t load_v_pair(const __half* v, int c) {
  // scale_v may be only 2-byte aligned (it lives inside the wire payload)
  return static_cast<uint32_t>(__half_as_ushort(v[c])) |
         (static_cast<uint32_t>(__half_as_ushort(v[c + 1])) << 16);
}

// BINARY: 8 elements from one code byte (zero-extended)
__device__ __forceinline__ H8 binary_apply8(const H8& b, uint32_t bits, __half2 u2, const uint32_t* vfrag) {
  H8 r;
  const uint32_t nb = bits ^ 0xFFu;  // 1 where the code bit is 0: (2 bit - 1) * scale flips the sign there (exact)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t scale = h22u(__hmul2_rn(u2, u2h2(vfrag[i])));  // fp16(U[n] V[c]), fastpath.py:109
    // one multiply moves bit 2i to position 15 and bit 2i+1 to position 31: the two partial products
    // occupy disjoint bit ranges (no carries), everything else is masked off
    const uint32_t flip = (nb * ((1u << (15 - 2 * i)) + (1u << (30 - 2 * i)))) & 0x80008000u;
    r.w[i] = h22u(__hadd2_rn(u2h2(b.w[i]), u2h2(scale ^ flip)));  // fastpath.py:116 / :363
  }
  return r;
}

// INT2: 8 elements from two code bytes (element e at bits 2e..2e+1 of the zero-extended 16-bit word).
// level = thr * f with f in {-0.5, -2, +0.5, +2} picked per element by a byte permute: the same fp16
// values as the reference's +-(0.5 thr) / +-(2 thr) (fastpath.py:565-572; scaling by a power of two and
// negation commute with the rounding).
__device__ __forceinline__ H8 int2_apply8(const H8& b, uint32_t codes, __half2 u2, const uint32_t* vfrag) {
  H8 r;
  const uint32_t lut = 0x4038C0B8u;  // high bytes of fp16 -0.5, -2, +0.5, +2 indexed by the 2-bit code
  const uint32_t ca = codes << 4, cb = codes << 10;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 thr = __hmul2_rn(u2h2(vfrag[i]), u2);  // fp16(chan tok), fastpath.py:714
    // selector nibbles: byte 1 <- lut[code of element 2i], byte 3 <- lut[code of element 2i+1], bytes 0/2 <- 0
    const uint32_t sel = ((ca >> (4 * i)) & 0x30u) | ((cb >> (4 * i)) & 0x3000u) | 0x0404u;
    const uint32_t fac = __byte_perm(lut, 0u, sel);
    r.w[i] = h22u(__hadd2_rn(u2h2(b.w[i]), __hmul2_rn(thr, u2h2(fac))));
  }
  return r;
}

template <int MODE, int G>
__global__ void __launch_bounds__(512) k_apply_codes(const ApplyParams p) {
  const int t = blockIdx.y;
  const uint8_t* __restrict__ packed = p.packed[t];
  const __half* __restrict__ su = p.scale_u[t];
  const __half* __restrict__ sv = p.scale_v[t];
  const __half* __restrict__ base = p.base[t];
  __half* __restrict__ recon = p.recon[t];
  const int N = p.N, C = p.C;
  const int groups = C >> 3;
  const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;

  constexpr int kUnroll = unroll_for(G);
  uint32_t vfrag[G][4];
#pragma unroll
  for (int j = 0; j < G; ++j) {
    const int g = tx + j * TX;
#pragma unroll
    for (int i = 0; i < 4; ++i) vfrag[j][i] = (g < groups) ? load_v_pair(sv, 8 * g + 2 * i) : 0u;
  }

  const int row_stride = gridDim.x * TY;
  for (int r = blockIdx.x * TY + ty; r < N; r += row_stride * kUnroll) {
    uint4 bv[kUnroll][G];
    uint32_t code[kUnroll][G];
    __half uu[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int rr = r + u * row_stride;
      uu[u] = __float2half_rn(0.f);
      if (rr < N) uu[u] = su[rr];
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const int g = tx + j * TX;
        bv[u][j] = make_uint4(0, 0, 0, 0);
        code[u][j] = 0;
        if (rr < N && g < groups) {
          if (base != nullptr) bv[u][j] = ldg_stream(base + static_cast<size_t>(rr) * C + 8 * g);
          if (MODE == MODE_BINARY) {
            code[u][j] = packed[static_cast<size_t>(rr) * groups + g];
          } else {
            const uint8_t* q = packed + (static_cast<size_t>(rr) * groups + g) * 2;
            code[u][j] = static_cast<uint32_t>(q[0]) | (static_cast<uint32_t>(q[1]) << 8);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int rr = r + u * row_stride;
      if (rr < N) {
        const __half2 u2 = __half2half2(uu[u]);
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const int g = tx + j * TX;
          if (g < groups) {
            H8 out;
            if (MODE == MODE_BINARY)
              out = binary_apply8(as_h8(bv[u][j]), code[u][j], u2, vfrag[j]);
            else
              out = int2_apply8(as_h8(bv[u][j]), code[u][j], u2, vfrag[j]);
            stg_stream(recon + static_cast<size_t>(rr) * C + 8 * g, as_u4(out));
          }
        }
      }
    }
  }
}

// BINARY with rank-K scales, K > 1 (deprecated in the reference, main.py:188-189): the scale
// sum_k U[n,k] V[c,k] is accumulated in fp32 in k order and rounded once.
__global__ void __launch_bounds__(256) k_binary_apply_rank_k(const uint8_t* __restrict__ packed,
                                                            const __half* __restrict__ su,
                                                            const __half* __restrict__ sv,
                                                            const __half* __restrict__ base,
                                                            __half* __restrict__ recon, int N, int C,
                                                            int K) {
  const int groups = C >> 3;
  const size_t total = static_cast<size_t>(N) * groups;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / groups), g = static_cast<int>(i % groups);
    const uint32_t bits = packed[i];
    uint4 bvec = make_uint4(0, 0, 0, 0);
    if (base != nullptr) bvec = ldg_stream(base + static_cast<size_t>(n) * C + 8 * g);
    const __half* bh = reinterpret_cast<const __half*>(&bvec);
    __align__(16) __half out[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = 8 * g + e;
      float s = 0.f;
      for (int k = 0; k < K; ++k)
        s += __half2float(su[static_cast<size_t>(n) * K + k]) * __half2float(sv[static_cast<size_t>(c) * K + k]);
      __half sc = __float2half_rn(s);
      if (((bits >> e) & 1u) == 0) sc = __hneg(sc);
      out[e] = __hadd_rn(bh[e], sc);
    }
    stg_stream(recon + static_cast<size_t>(n) * C + 8 * g, *reinterpret_cast<uint4*>(out));
  }
}

// ---------------------------------------------------------------------------------------
// INT2 encode: codes (and optional error-feedback base) from x, base and final scales
// ---------------------------------------------------------------------------------------
struct Int2EncodeParams {
  const __half* x[CF_MAX_BATCH];
  const __half* base[CF_MAX_BATCH];
  const __half* scale_u[CF_MAX_BATCH];
  const __half* scale_v[CF_MAX_BATCH];
  uint8_t* packed[CF_MAX_BATCH];
  __half* new_base[CF_MAX_BATCH];  // may be null
  int N, C;
};

template <int G>
__global__ void __launch_bounds__(512) k_int2_encode(const Int2EncodeParams p) {
  const int t = blockIdx.y;
  const __half* __restrict__ x = p.x[t];
  const __half* __restrict__ base = p.base[t];
  const __half* __restrict__ su = p.scale_u[t];
  const __half* __restrict__ sv = p.scale_v[t];
  uint8_t* __restrict__ packed = p.packed[t];
  __half* __restrict__ new_base = p.new_base[t];
  const int N = p.N, C = p.C;
  const int groups = C >> 3;
  const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;

  constexpr int kUnroll = unroll_for(G);
  uint32_t vfrag[G][4];
#pragma unroll
  for (int j = 0; j < G; ++j) {
    const int g = tx + j * TX;
#pragma unroll
    for (int i = 0; i < 4; ++i) vfrag[j][i] = (g < groups) ? load_v_pair(sv, 8 * g + 2 * i) : 0u;
  }
  const __half2 zero2 = __float2half2_rn(0.f);
  const int row_stride = gridDim.x * TY;
  for (int r = blockIdx.x * TY + ty; r < N; r += row_stride * kUnroll) {
    uint4 xv[kUnroll][G], bv[kUnroll][G];
    __half uu[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int rr = r + u * row_stride;
      uu[u] = __float2half_rn(0.f);
      if (rr < N) uu[u] = su[rr];
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const int g = tx + j * TX;
        xv[u][j] = make_uint4(0, 0, 0, 0);
        bv[u][j] = make_uint4(0, 0, 0, 0);
        if (rr < N && g < groups) {
          const size_t off = static_cast<size_t>(rr) * C + 8 * g;
          xv[u][j] = ldg_stream(x + off);
          if (base != nullptr) bv[u][j] = ldg_stream(base + off);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int rr = r + u * row_stride;
      if (rr < N) {
        const __half2 u2 = __half2half2(uu[u]);
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const int g = tx + j * TX;
          if (g < groups) {
            const H8 b = as_h8(bv[u][j]);
            const H8 d = h8_sub(as_h8(xv[u][j]), b);
            uint32_t codes = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const __half2 thr = __hmul2_rn(u2h2(vfrag[j][i]), u2);            // fastpath.py:536
              const uint32_t sgn = __hge2_mask(u2h2(d.w[i]), zero2);             // fastpath.py:539
              const uint32_t mag = __hgt2_mask(__habs2(u2h2(d.w[i])), thr);      // fastpath.py:540
              const uint32_t c0 = ((sgn & 1u) << 1) | (mag & 1u);
              const uint32_t c1 = (((sgn >> 16) & 1u) << 1) | ((mag >> 16) & 1u);
              codes |= (c0 | (c1 << 2)) << (4 * i);
            }
            uint8_t* q = packed + (static_cast<size_t>(rr) * groups + g) * 2;
            q[0] = static_cast<uint8_t>(codes & 0xFFu);
            q[1] = static_cast<uint8_t>(codes >> 8);
            if (new_base != nullptr) {
              const H8 nb = int2_apply8(b, codes, u2, vfrag[j]);
              stg_stream(new_base + static_cast<size_t>(rr) * C + 8 * g, as_u4(nb));
            }
          }
        }
      }
    }
  }
}

}  // namespace cf

#include "cf_sign_tma.cuh"

namespace cf {

// ---------------------------------------------------------------------------------------
// host-side launch logic
// ---------------------------------------------------------------------------------------
static bool env_flag(const char* name, bool dflt) {
  const char* e = getenv(name);
  if (e == nullptr || e[0] == 0) return dflt;
  return e[0] != '0';
}
// CF_LEGACY_KERNELS=1 forces the register-staged kernels (parity tests exercise both paths)
static bool legacy_forced() { return env_flag("CF_LEGACY_KERNELS", false); }
// CF_PDL=0 disables programmatic dependent launch between our own back-to-back kernels
static bool pdl_enabled() { return env_flag("CF_PDL", true); }

template <typename... KArgs, typename... Args>
static cudaError_t launch_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                             Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
  if (smem > 48 * 1024) {
    // once per (kernel, size): the attribute call costs about as much as the launch itself (per-device state:
    // one process drives one GPU here)
    static thread_local std::unordered_map<const void*, size_t> granted;  // kernel -> largest size set so far
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

// L2 residency plan (CF_L2_HINTS=0 disables): pass 1 loads `base` with an evict_last policy when all the
// bases of the launch (batch * N * C * 2 bytes) fit comfortably in the 126 MB L2, so the apply / INT2 encode
// pass that follows re-reads them from L2 instead of HBM; x, the codes and the rewritten base lines are
// evict_first (they are not touched again before the next denoising step).
static int l2_keep_base(int64_t N, int64_t C, int batch) {
  if (pipe_env_int("CF_L2_HINTS", 1) == 0) return 0;
  const int64_t cap = static_cast<int64_t>(pipe_env_int("CF_L2_KEEP_MB", 72)) << 20;
  return N * C * 2 * batch <= cap ? 1 : 0;
}
static int l2_stream_hints() { return pipe_env_int("CF_L2_HINTS", 1) != 0 ? 1 : 0; }

// CF_PUT_BULK=0 keeps the per-thread sub-word stores of the fused put (A/B); the bulk flavour needs whole tiles of
// code bytes to be 16-byte multiples at 16-byte aligned destinations
static bool put_bulk_ok(const FanOut* fan, int batch, int64_t code_row_bytes) {
  if (fan == nullptr || pipe_env_int("CF_PUT_BULK", 1) == 0 || code_row_bytes % 16 != 0) return false;
  for (int i = 0; i < batch * fan->n_dst; ++i)
    if (!aligned16(fan->dst[i])) return false;
  return true;
}

static PipeArgs pipe_args(const PipeGeom& g, int rows_per_cta, int l2_hints, bool early_load) {
  PipeArgs a{};
  a.put_tile = g.put_tile;
  a.l2_hints = l2_hints;
  a.early_load = (early_load && pdl_enabled()) ? 1 : 0;
  a.TX = g.TX; a.TY = g.TY; a.R = g.R; a.stages = g.stages; a.chunk_rows = g.chunk_rows; a.u_cap = g.u_cap;
  a.tile_bytes = g.tile_bytes; a.stage_bytes = g.stage_bytes; a.rows_per_cta = rows_per_cta;
  return a;
}

static TileSched make_tile_sched(const PipeGeom& g, int64_t N, int batch, int* n_cta) {
  TileSched ts{};
  ts.tiles_per_tensor = static_cast<int>((N + g.R - 1) / g.R);
  ts.total_tiles = ts.tiles_per_tensor * batch;
  int ctas = sm_count() * g.ctas_per_sm;
  if (ctas > ts.total_tiles) ctas = ts.total_tiles;
  ts.tiles_per_cta = (ts.total_tiles + ctas - 1) / ctas;
  const int cap = g.u_cap / g.R > 0 ? g.u_cap / g.R : 1;
  if (ts.tiles_per_cta > cap) ts.tiles_per_cta = cap;
  *n_cta = (ts.total_tiles + ts.tiles_per_cta - 1) / ts.tiles_per_cta;
  return ts;
}

struct StatsPlan {
  RowGeom geom;
  bool tma;          // bulk-async pipelined kernel (cf_sign_tma.cuh) instead of the register-staged one
  PipeGeom pipe;
  int B;             // row blocks per tensor
  int rows_per_cta;
  size_t rowmean_bytes, tokpart_bytes, colpart_bytes, per_tensor_bytes;
  size_t smem_bytes;
};

static StatsPlan make_stats_plan(int64_t N, int64_t C, int batch, bool allow_tma = false) {
  StatsPlan pl;
  pl.geom = make_row_geom(C);
  pl.tma = false;
  if (allow_tma && !legacy_forced()) {
    pl.pipe = make_pipe_geom(C, 2, 0, false, static_cast<int>(C / 8));
    if (pl.pipe.ok) {
      int B = sm_count() * pl.pipe.ctas_per_sm / (batch > 0 ? batch : 1);
      if (B < 1) B = 1;
      int64_t rpc = (N + B - 1) / B;
      rpc = (rpc + pl.pipe.R - 1) / pl.pipe.R * pl.pipe.R;  // tile starts stay stage-aligned
      pl.tma = true;
      pl.rows_per_cta = static_cast<int>(rpc);
      pl.B = static_cast<int>((N + rpc - 1) / rpc);
      pl.rowmean_bytes = round_up(static_cast<size_t>(N) * 2, 256);
      pl.tokpart_bytes = round_up(static_cast<size_t>(pl.B) * 4, 256);
      pl.colpart_bytes = round_up(static_cast<size_t>(pl.B) * C * 4, 256);
      pl.per_tensor_bytes = pl.rowmean_bytes + pl.tokpart_bytes + pl.colpart_bytes;
      pl.smem_bytes = pl.pipe.smem_bytes;
      return pl;
    }
  }
  const int threads = pl.geom.TX * pl.geom.TY;
  const int ctas_per_sm = threads >= 512 ? 2 : (1024 / threads);
  const int total = sm_count() * ctas_per_sm;
  int B = total / (batch > 0 ? batch : 1);
  if (B < 1) B = 1;
  const int64_t max_b = (N + pl.geom.TY - 1) / pl.geom.TY;
  if (B > max_b) B = static_cast<int>(max_b);
  pl.rows_per_cta = static_cast<int>((N + B - 1) / B);
  pl.B = static_cast<int>((N + pl.rows_per_cta - 1) / pl.rows_per_cta);
  pl.rowmean_bytes = round_up(static_cast<size_t>(N) * 2, 256);
  pl.tokpart_bytes = round_up(static_cast<size_t>(pl.B) * 4, 256);
  pl.colpart_bytes = round_up(static_cast<size_t>(pl.B) * C * 4, 256);
  pl.per_tensor_bytes = pl.rowmean_bytes + pl.tokpart_bytes + pl.colpart_bytes;
  const int NWX = pl.geom.TX / 32;
  size_t s1 = static_cast<size_t>(kRowChunk) * NWX * 4;
  size_t s2 = pl.geom.TY > 1 ? static_cast<size_t>(pl.geom.TY) * pl.geom.G * pl.geom.TX * 32 : 0;
  size_t s3 = 64 * 4;
  pl.smem_bytes = s1 > s2 ? s1 : s2;
  if (pl.smem_bytes < s3) pl.smem_bytes = s3;
  return pl;
}

size_t sign_codec_workspace_bytes(int64_t N, int64_t C, int batch) {
  // upper bound independent of the batch split: B never exceeds the batch=1 value
  StatsPlan pl = make_stats_plan(N, C, 1);
  return pl.per_tensor_bytes * static_cast<size_t>(batch > 0 ? batch : 1);
}

template <int MODE>
static int launch_stats(const StatsPlan& pl, const StatsParams& sp, int batch, cudaStream_t st, bool stable = false,
                        const FanOut* fan = nullptr) {
  if (pl.tma) {
    PipeArgs a = pipe_args(pl.pipe, pl.rows_per_cta, l2_keep_base(sp.N, sp.C, batch), stable);
    a.put_bulk = (MODE == MODE_BINARY && put_bulk_ok(fan, batch, sp.C / 8)) ? 1 : 0;
    dim3 grid(pl.B, batch), block(pl.pipe.TX * pl.pipe.TY + 32);
    const int variant = (pl.pipe.G == 1 ? 0 : 2) + (pl.pipe.ctas_per_sm == 1 ? 0 : 1);
    // only BINARY emits codes in pass 1: the INT2 fused put stores them from the encode kernel
    const bool put = fan != nullptr && MODE == MODE_BINARY;
    const FanOut f = fan != nullptr ? *fan : FanOut{};
#define CF_LAUNCH_STATS_TMA(GG, OO)                                                                                  \
  CF_CHECK_CUDA(put ? launch_ex(k_delta_stats_tma<MODE, GG, OO, (MODE == MODE_BINARY)>, grid, block,                 \
                                pl.pipe.smem_bytes, st, true, sp, a, f)                                              \
                    : launch_ex(k_delta_stats_tma<MODE, GG, OO, false>, grid, block, pl.pipe.smem_bytes, st, true,   \
                                sp, a, f))
    switch (variant) {
      case 0: CF_LAUNCH_STATS_TMA(1, 1); break;
      case 1: CF_LAUNCH_STATS_TMA(1, 2); break;
      case 2: CF_LAUNCH_STATS_TMA(2, 1); break;
      default: CF_LAUNCH_STATS_TMA(2, 2); break;
    }
#undef CF_LAUNCH_STATS_TMA
    return CF_OK;
  }
  if (fan != nullptr) {
    set_error("the fused put needs the pipelined kernels: base set, C %% 8 == 0, 64 <= C <= 8192");
    return CF_ERR_UNSUPPORTED;
  }
  dim3 grid(pl.B, batch), block(pl.geom.TX, pl.geom.TY);
#define CF_LAUNCH_STATS(GG)                                                                    \
  case GG: {                                                                                   \
    if (pl.smem_bytes > 48 * 1024)                                                             \
      CF_CHECK_CUDA(cudaFuncSetAttribute(k_delta_stats<MODE, GG>,                              \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                         static_cast<int>(pl.smem_bytes)));                    \
    k_delta_stats<MODE, GG><<<grid, block, pl.smem_bytes, st>>>(sp);                           \
  } break;
  switch (pl.geom.G) {
    CF_LAUNCH_STATS(1)
    CF_LAUNCH_STATS(2)
    CF_LAUNCH_STATS(4)
    CF_LAUNCH_STATS(8)
    default:
      set_error("unsupported column geometry G=%d", pl.geom.G);
      return CF_ERR_UNSUPPORTED;
  }
#undef CF_LAUNCH_STATS
  CF_CHECK_LAUNCH();
  return CF_OK;
}

static int apply_grid_x(const RowGeom& g, int64_t N, int batch) {
  const int threads = g.TX * g.TY;
  const int ctas_per_sm = threads >= 512 ? 2 : (1024 / threads);
  int64_t bx = static_cast<int64_t>(sm_count()) * ctas_per_sm / (batch > 0 ? batch : 1);
  const int64_t per = static_cast<int64_t>(g.TY) * unroll_for(g.G);
  const int64_t max_b = (N + per - 1) / per;
  if (bx > max_b) bx = max_b;
  if (bx < 1) bx = 1;
  return static_cast<int>(bx);
}

template <int MODE>
static int launch_apply(const ApplyParams& ap, int batch, cudaStream_t st, bool stable = false) {
  const int code_row = (MODE == MODE_BINARY) ? ap.C / 8 : ap.C / 4;
  const bool need_wait = ap.expected != nullptr;
  bool tma = (need_wait || !legacy_forced()) && code_row % 16 == 0;
  for (int t = 0; t < batch && tma; ++t)
    tma = ap.base[t] != nullptr && aligned16(ap.base[t]) && aligned16(ap.packed[t]);
  if (tma) {
    const PipeGeom pg = make_pipe_geom(ap.C, 1, code_row, true);
    if (pg.ok) {
      int n_cta = 1;
      const TileSched ts = make_tile_sched(pg, ap.N, batch, &n_cta);
      const PipeArgs a = pipe_args(pg, 0, l2_stream_hints(), stable);
      dim3 grid(n_cta), block(pg.TX * pg.TY + 32);
      const int variant = (pg.G == 1 ? 0 : 2) + (pg.ctas_per_sm == 1 ? 0 : 1);
      switch (variant) {
        case 0: CF_CHECK_CUDA(launch_ex(k_apply_codes_tma<MODE, 1, 1>, grid, block, pg.smem_bytes, st, true, ap, a, ts)); break;
        case 1: CF_CHECK_CUDA(launch_ex(k_apply_codes_tma<MODE, 1, 2>, grid, block, pg.smem_bytes, st, true, ap, a, ts)); break;
        case 2: CF_CHECK_CUDA(launch_ex(k_apply_codes_tma<MODE, 2, 1>, grid, block, pg.smem_bytes, st, true, ap, a, ts)); break;
        default: CF_CHECK_CUDA(launch_ex(k_apply_codes_tma<MODE, 2, 2>, grid, block, pg.smem_bytes, st, true, ap, a, ts)); break;
      }
      return CF_OK;
    }
  }
  if (need_wait) {
    set_error("flag-waiting decompress needs the pipelined kernel: 16-byte aligned base / codes, C %% 128 == 0 "
              "(C %% 64 for INT2), C <= 8192");
    return CF_ERR_UNSUPPORTED;
  }
  const RowGeom g = make_row_geom(ap.C);
  dim3 grid(apply_grid_x(g, ap.N, batch), batch), block(g.TX, g.TY);
  switch (g.G) {
    case 1: k_apply_codes<MODE, 1><<<grid, block, 0, st>>>(ap); break;
    case 2: k_apply_codes<MODE, 2><<<grid, block, 0, st>>>(ap); break;
    case 4: k_apply_codes<MODE, 4><<<grid, block, 0, st>>>(ap); break;
    case 8: k_apply_codes<MODE, 8><<<grid, block, 0, st>>>(ap); break;
    default:
      set_error("unsupported column geometry G=%d", g.G);
      return CF_ERR_UNSUPPORTED;
  }
  CF_CHECK_LAUNCH();
  return CF_OK;
}

static int launch_int2_encode(const Int2EncodeParams& ep, int batch, cudaStream_t st, bool stable = false,
                              const FanOut* fan = nullptr) {
  bool tma = fan != nullptr || !legacy_forced();
  for (int t = 0; t < batch && tma; ++t)
    tma = ep.base[t] != nullptr && aligned16(ep.base[t]) && aligned16(ep.x[t]) && aligned2(ep.packed[t]);
  if (tma) {
    const PipeGeom pg = make_pipe_geom(ep.C, 2, 0, true, ep.C / 4);
    if (pg.ok) {
      int n_cta = 1;
      const TileSched ts = make_tile_sched(pg, ep.N, batch, &n_cta);
      PipeArgs a = pipe_args(pg, 0, l2_keep_base(ep.N, ep.C, batch), stable);
      a.put_bulk = put_bulk_ok(fan, batch, ep.C / 4) ? 1 : 0;
      dim3 grid(n_cta), block(pg.TX * pg.TY + 32);
      const int variant = (pg.G == 1 ? 0 : 2) + (pg.ctas_per_sm == 1 ? 0 : 1);
      const bool put = fan != nullptr;
      const FanOut f = put ? *fan : FanOut{};
#define CF_LAUNCH_ENC_TMA(GG, OO)                                                                                      \
  CF_CHECK_CUDA(put ? launch_ex(k_int2_encode_tma<GG, OO, true>, grid, block, pg.smem_bytes, st, true, ep, a, ts, f)   \
                    : launch_ex(k_int2_encode_tma<GG, OO, false>, grid, block, pg.smem_bytes, st, true, ep, a, ts, f))
      switch (variant) {
        case 0: CF_LAUNCH_ENC_TMA(1, 1); break;
        case 1: CF_LAUNCH_ENC_TMA(1, 2); break;
        case 2: CF_LAUNCH_ENC_TMA(2, 1); break;
        default: CF_LAUNCH_ENC_TMA(2, 2); break;
      }
#undef CF_LAUNCH_ENC_TMA
      return CF_OK;
    }
  }
  if (fan != nullptr) {
    set_error("the fused put needs the pipelined kernels: 16-byte aligned x / base, C %% 8 == 0, 64 <= C <= 8192");
    return CF_ERR_UNSUPPORTED;
  }
  const RowGeom g = make_row_geom(ep.C);
  dim3 grid(apply_grid_x(g, ep.N, batch), batch), block(g.TX, g.TY);
  switch (g.G) {
    case 1: k_int2_encode<1><<<grid, block, 0, st>>>(ep); break;
    case 2: k_int2_encode<2><<<grid, block, 0, st>>>(ep); break;
    case 4: k_int2_encode<4><<<grid, block, 0, st>>>(ep); break;
    case 8: k_int2_encode<8><<<grid, block, 0, st>>>(ep); break;
    default:
      set_error("unsupported column geometry G=%d", g.G);
      return CF_ERR_UNSUPPORTED;
  }
  CF_CHECK_LAUNCH();
  return CF_OK;
}

static int check_shape(int64_t N, int64_t C, int batch) {
  CF_CHECK_ARG(batch >= 1 && batch <= CF_MAX_BATCH, "batch %d out of range [1,%d]", batch, CF_MAX_BATCH);
  CF_CHECK_ARG(N >= 1 && N < (int64_t(1) << 31), "N=%lld out of range", (long long)N);
  CF_CHECK_ARG(C >= 8 && C % 8 == 0 && C <= 32768, "C=%lld must be a multiple of 8 in [8, 32768]", (long long)C);
  CF_CHECK_ARG(N * C < (int64_t(1) << 40), "tensor too large");
  return CF_OK;
}

template <int MODE>
static int sign_compress(int batch, const void* const* x, const void* const* base, void* const* new_base,
                         void* const* packed, void* const* scale_u, void* const* scale_v, int64_t N,
                         int64_t C, void* workspace, size_t workspace_bytes, cf_stream_t stream,
                         int passes = CF_PASS_ALL, bool stable = false) {
  if (int rc = check_shape(N, C, batch)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bool all_base = base != nullptr;
  for (int t = 0; t < batch && all_base; ++t) all_base = base[t] != nullptr;
  StatsPlan pl = make_stats_plan(N, C, batch, all_base);
  CF_CHECK_ARG(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0,
               "workspace must be non-null and 256-byte aligned");
  if (pl.per_tensor_bytes * batch > workspace_bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", pl.per_tensor_bytes * batch, workspace_bytes);
    return CF_ERR_WORKSPACE;
  }
  StatsParams sp{};
  FinalizeParams fp{};
  sp.N = fp.N = static_cast<int>(N);
  sp.C = fp.C = static_cast<int>(C);
  sp.rows_per_cta = pl.rows_per_cta;
  fp.B = pl.B;
  bool any_update = false;
  for (int t = 0; t < batch; ++t) {
    CF_CHECK_ARG(x[t] && packed[t] && scale_u[t] && scale_v[t], "null pointer in tensor %d", t);
    CF_CHECK_ARG(aligned16(x[t]) && (!base || !base[t] || aligned16(base[t])) &&
                     (!new_base || !new_base[t] || aligned16(new_base[t])),
                 "x/base/new_base must be 16-byte aligned (tensor %d)", t);
    CF_CHECK_ARG(aligned2(scale_u[t]) && aligned2(scale_v[t]), "scales must be 2-byte aligned");
    char* ws = static_cast<char*>(workspace) + pl.per_tensor_bytes * t;
    sp.x[t] = static_cast<const __half*>(x[t]);
    sp.base[t] = base ? static_cast<const __half*>(base[t]) : nullptr;
    sp.packed[t] = static_cast<uint8_t*>(packed[t]);
    sp.rowmean[t] = reinterpret_cast<__half*>(ws);
    sp.tokpart[t] = reinterpret_cast<float*>(ws + pl.rowmean_bytes);
    sp.colpart[t] = reinterpret_cast<float*>(ws + pl.rowmean_bytes + pl.tokpart_bytes);
    fp.rowmean[t] = sp.rowmean[t];
    fp.tokpart[t] = sp.tokpart[t];
    fp.colpart[t] = sp.colpart[t];
    fp.scale_u[t] = static_cast<__half*>(scale_u[t]);
    fp.scale_v[t] = static_cast<__half*>(scale_v[t]);
    if (new_base && new_base[t]) any_update = true;
  }
  if (passes & CF_PASS_STATS)
    if (int rc = launch_stats<MODE>(pl, sp, batch, st, stable)) return rc;
  if (passes & CF_PASS_FINALIZE) {
    dim3 grid(static_cast<unsigned>((C + 31) / 32), batch);  // one CTA per 32 columns
    CF_CHECK_CUDA(launch_ex(k_finalize_scales<MODE, false>, grid, dim3(1024), 0, st, true, fp, FanOut{}, 0));
  }
  if (!(passes & CF_PASS_ENCODE)) return CF_OK;
  if (MODE == MODE_BINARY) {
    if (any_update) {
      ApplyParams ap{};
      ap.N = sp.N; ap.C = sp.C; ap.K = 1;
      for (int t = 0; t < batch; ++t) {
        CF_CHECK_ARG(new_base[t] != nullptr, "batched compress: new_base must be set for all tensors or none");
        ap.packed[t] = sp.packed[t];
        ap.scale_u[t] = fp.scale_u[t];
        ap.scale_v[t] = fp.scale_v[t];
        ap.base[t] = sp.base[t];
        ap.recon[t] = static_cast<__half*>(new_base[t]);
      }
      if (int rc = launch_apply<MODE_BINARY>(ap, batch, st, stable)) return rc;
    }
  } else {
    Int2EncodeParams ep{};
    ep.N = sp.N; ep.C = sp.C;
    for (int t = 0; t < batch; ++t) {
      CF_CHECK_ARG(!any_update || new_base[t] != nullptr,
                   "batched compress: new_base must be set for all tensors or none");
      ep.x[t] = sp.x[t];
      ep.base[t] = sp.base[t];
      ep.scale_u[t] = fp.scale_u[t];
      ep.scale_v[t] = fp.scale_v[t];
      ep.packed[t] = sp.packed[t];
      ep.new_base[t] = any_update ? static_cast<__half*>(new_base[t]) : nullptr;
    }
    if (int rc = launch_int2_encode(ep, batch, st, stable)) return rc;
  }
  return CF_OK;
}

template <int MODE>
static int sign_decompress(int batch, const void* const* packed, const void* const* scale_u,
                           const void* const* scale_v, int K, const void* const* base, void* const* recon,
                           int64_t N, int64_t C, cf_stream_t stream, const void* const* wait_flag = nullptr,
                           const void* expected = nullptr, void* error = nullptr, bool stable = false) {
  if (int rc = check_shape(N, C, batch)) return rc;
  CF_CHECK_ARG(K >= 1, "K must be >= 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ApplyParams ap{};
  ap.N = static_cast<int>(N); ap.C = static_cast<int>(C); ap.K = K;
  for (int t = 0; t < batch; ++t) {
    CF_CHECK_ARG(packed[t] && scale_u[t] && scale_v[t] && recon[t], "null pointer in tensor %d", t);
    CF_CHECK_ARG(aligned16(recon[t]) && (!base || !base[t] || aligned16(base[t])),
                 "base/recon must be 16-byte aligned (tensor %d)", t);
    CF_CHECK_ARG(aligned2(scale_u[t]) && aligned2(scale_v[t]), "scales must be 2-byte aligned");
    ap.packed[t] = static_cast<const uint8_t*>(packed[t]);
    ap.scale_u[t] = static_cast<const __half*>(scale_u[t]);
    ap.scale_v[t] = static_cast<const __half*>(scale_v[t]);
    ap.base[t] = base ? static_cast<const __half*>(base[t]) : nullptr;
    ap.recon[t] = static_cast<__half*>(recon[t]);
    ap.wait_flag[t] = (wait_flag && expected) ? static_cast<const uint32_t*>(wait_flag[t]) : nullptr;
  }
  ap.expected = (wait_flag && expected) ? static_cast<const uint32_t*>(expected) : nullptr;
  ap.error = static_cast<uint32_t*>(error);
  ap.wait_mode = pipe_env_int("CF_WAIT_MODE", 0);
  CF_CHECK_ARG(ap.expected == nullptr || (K == 1 && ap.error != nullptr), "flag waiting needs K == 1 and an error word");
  if (K > 1) {
    CF_CHECK_ARG(MODE == MODE_BINARY, "rank-K scales are only defined for BINARY");
    for (int t = 0; t < batch; ++t) {
      const size_t total = static_cast<size_t>(N) * (C / 8);
      int blocks = static_cast<int>((total + 255) / 256);
      const int cap = sm_count() * 8;
      if (blocks > cap) blocks = cap;
      k_binary_apply_rank_k<<<blocks, 256, 0, st>>>(ap.packed[t], ap.scale_u[t], ap.scale_v[t], ap.base[t],
                                                    ap.recon[t], ap.N, ap.C, K);
      CF_CHECK_LAUNCH();
    }
    return CF_OK;
  }
  return launch_apply<MODE>(ap, batch, st, stable);
}

// Fused compress + put: no local send buffer, no put kernel.
//   BINARY: stats<PUT> stores the sign bytes into every destination slot, finalize<PUT> stores the scale
//           vectors and publishes the flags.
//   INT2:   finalize<PUT> stores the scale vectors, encode<PUT> (which reads them back from the local
//           destination `self_dst`) stores the code words and publishes the flags.
template <int MODE>
static int sign_compress_put(int passes, int batch, const void* const* x, const void* const* base, int n_dst,
                             int self_dst, void* const* dst_payload, void* const* dst_flag, void* local_count,
                             void* local_ticket, int64_t N, int64_t C, void* workspace, size_t workspace_bytes,
                             cf_stream_t stream, bool stable) {
  if (int rc = check_shape(N, C, batch)) return rc;
  CF_CHECK_ARG(x && base && dst_payload && dst_flag && local_count && local_ticket, "null pointer");
  CF_CHECK_ARG(MODE == MODE_BINARY || (self_dst >= 0 && self_dst < n_dst),
               "INT2 needs self_dst (the local destination the encode pass reads the scales from)");
  CF_CHECK_ARG(n_dst >= 1 && n_dst <= CF_MAX_PEERS && batch * n_dst <= CF_MAX_FANOUT,
               "n_dst %d / batch %d out of range (n_dst <= %d, batch * n_dst <= %d)", n_dst, batch, CF_MAX_PEERS,
               CF_MAX_FANOUT);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int t = 0; t < batch; ++t)
    CF_CHECK_ARG(x[t] && base[t] && aligned16(x[t]) && aligned16(base[t]), "x/base must be set and 16-byte aligned");
  StatsPlan pl = make_stats_plan(N, C, batch, true);
  CF_CHECK_ARG(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0,
               "workspace must be non-null and 256-byte aligned");
  if (pl.per_tensor_bytes * batch > workspace_bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", pl.per_tensor_bytes * batch, workspace_bytes);
    return CF_ERR_WORKSPACE;
  }
  StatsParams sp{};
  FinalizeParams fp{};
  FanOut f{};
  sp.N = fp.N = static_cast<int>(N);
  sp.C = fp.C = static_cast<int>(C);
  sp.rows_per_cta = pl.rows_per_cta;
  fp.B = pl.B;
  f.n_dst = n_dst;
  f.publish_mode = pipe_env_int("CF_PUBLISH_MODE", 2);
  const bool publish_kernel = f.publish_mode == 2;
  f.u_off = static_cast<unsigned long long>(N) * (MODE == MODE_BINARY ? C / 8 : C / 4);
  f.v_off = f.u_off + static_cast<unsigned long long>(N) * 2;
  f.count = static_cast<uint32_t*>(local_count);
  f.done = static_cast<uint32_t*>(local_ticket);
  for (int q = 0; q < n_dst; ++q) {
    CF_CHECK_ARG(dst_flag[q] != nullptr, "destination %d: null flag", q);
    f.flag[q] = static_cast<uint32_t*>(dst_flag[q]);
  }
  bool vec16 = pipe_env_int("CF_PUT_BULK", 1) != 0 && N % 8 == 0 && C % 32 == 0 && f.u_off % 16 == 0;
  for (int i = 0; i < batch * n_dst && vec16; ++i) vec16 = aligned16(dst_payload[i]);
  f.vec16 = vec16 ? 1 : 0;
  for (int t = 0; t < batch; ++t) {
    char* ws = static_cast<char*>(workspace) + pl.per_tensor_bytes * t;
    sp.x[t] = static_cast<const __half*>(x[t]);
    sp.base[t] = static_cast<const __half*>(base[t]);
    sp.rowmean[t] = reinterpret_cast<__half*>(ws);
    sp.tokpart[t] = reinterpret_cast<float*>(ws + pl.rowmean_bytes);
    sp.colpart[t] = reinterpret_cast<float*>(ws + pl.rowmean_bytes + pl.tokpart_bytes);
    fp.rowmean[t] = sp.rowmean[t];
    fp.tokpart[t] = sp.tokpart[t];
    fp.colpart[t] = sp.colpart[t];
    for (int q = 0; q < n_dst; ++q) {
      void* d = dst_payload[t * n_dst + q];
      CF_CHECK_ARG(d != nullptr && aligned2(d), "tensor %d destination %d: bad payload pointer", t, q);
      f.dst[t * n_dst + q] = static_cast<unsigned char*>(d);
    }
  }
  if (passes & CF_PASS_STATS)
    if (int rc = launch_stats<MODE>(pl, sp, batch, st, stable, &f)) return rc;
  if (passes & CF_PASS_FINALIZE) {
    dim3 grid(static_cast<unsigned>((C + 31) / 32), batch);
    CF_CHECK_CUDA(launch_ex(k_finalize_scales<MODE, true>, grid, dim3(1024), 0, st, true, fp, f,
                            (MODE == MODE_BINARY && !publish_kernel) ? 1 : 0));
  }
  if (MODE == MODE_INT2 && (passes & CF_PASS_ENCODE)) {
    Int2EncodeParams ep{};
    ep.N = sp.N; ep.C = sp.C;
    for (int t = 0; t < batch; ++t) {
      unsigned char* own = f.dst[t * n_dst + self_dst];
      ep.x[t] = sp.x[t];
      ep.base[t] = sp.base[t];
      ep.scale_u[t] = reinterpret_cast<const __half*>(own + f.u_off);
      ep.scale_v[t] = reinterpret_cast<const __half*>(own + f.v_off);
      ep.packed[t] = own;  // unused by the PUT kernel (it stores to every f.dst)
      ep.new_base[t] = nullptr;
    }
    if (int rc = launch_int2_encode(ep, batch, st, stable, &f)) return rc;
  }
  // the call's LAST kernel completes the put: publish behind it
  const int last_pass = MODE == MODE_BINARY ? CF_PASS_FINALIZE : CF_PASS_ENCODE;
  if (publish_kernel && (passes & last_pass)) {
    PublishParams pp{};
    pp.n_dst = n_dst;
    pp.count = f.count;
    {
      const char* e = getenv("CF_PUBLISH_FENCE");
      pp.sc_fence = (e && e[0] == 's') ? 1 : 0;
    }
    for (int q = 0; q < n_dst; ++q) pp.flag[q] = f.flag[q];
    CF_CHECK_CUDA(launch_ex(k_publish_flags, dim3(1), dim3(32), 0, st, true, pp));
  }
  return CF_OK;
}

}  // namespace cf

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" {

int cf_binary_compress_batched(int batch, const void* const* x, const void* const* base, void* const* new_base,
                               void* const* packed, void* const* scale_u, void* const* scale_v, int64_t N,
                               int64_t C, void* workspace, size_t workspace_bytes, cf_stream_t stream) {
  return cf::sign_compress<cf::MODE_BINARY>(batch, x, base, new_base, packed, scale_u, scale_v, N, C, workspace,
                                            workspace_bytes, stream);
}
int cf_binary_compress(const void* x, const void* base, void* new_base, void* packed, void* scale_u,
                       void* scale_v, int64_t N, int64_t C, void* workspace, size_t workspace_bytes,
                       cf_stream_t stream) {
  return cf_binary_compress_batched(1, &x, &base, &new_base, &packed, &scale_u, &scale_v, N, C, workspace,
                                    workspace_bytes, stream);
}
int cf_binary_decompress_batched(int batch, const void* const* packed, const void* const* scale_u,
                                 const void* const* scale_v, const void* const* base, void* const* recon,
                                 int64_t N, int64_t C, cf_stream_t stream) {
  return cf::sign_decompress<cf::MODE_BINARY>(batch, packed, scale_u, scale_v, 1, base, recon, N, C, stream);
}
int cf_binary_decompress(const void* packed, const void* scale_u, const void* scale_v, int K, const void* base,
                         void* recon, int64_t N, int64_t C, cf_stream_t stream) {
  return cf::sign_decompress<cf::MODE_BINARY>(1, &packed, &scale_u, &scale_v, K, &base, &recon, N, C, stream);
}

int cf_int2_compress_batched(int batch, const void* const* x, const void* const* base, void* const* new_base,
                             void* const* packed, void* const* scale_u, void* const* scale_v, int64_t N,
                             int64_t C, void* workspace, size_t workspace_bytes, cf_stream_t stream) {
  return cf::sign_compress<cf::MODE_INT2>(batch, x, base, new_base, packed, scale_u, scale_v, N, C, workspace,
                                          workspace_bytes, stream);
}
int cf_int2_compress(const void* x, const void* base, void* new_base, void* packed, void* scale_u, void* scale_v,
                     int64_t N, int64_t C, void* workspace, size_t workspace_bytes, cf_stream_t stream) {
  return cf_int2_compress_batched(1, &x, &base, &new_base, &packed, &scale_u, &scale_v, N, C, workspace,
                                  workspace_bytes, stream);
}
int cf_int2_decompress_batched(int batch, const void* const* packed, const void* const* scale_u,
                               const void* const* scale_v, const void* const* base, void* const* recon, int64_t N,
                               int64_t C, cf_stream_t stream) {
  return cf::sign_decompress<cf::MODE_INT2>(batch, packed, scale_u, scale_v, 1, base, recon, N, C, stream);
}
int cf_int2_decompress(const void* packed, const void* scale_u, const void* scale_v, const void* base,
                       void* recon, int64_t N, int64_t C, cf_stream_t stream) {
  return cf::sign_decompress<cf::MODE_INT2>(1, &packed, &scale_u, &scale_v, 1, &base, &recon, N, C, stream);
}
int cf_sign_compress_passes(int codec, int passes, int batch, const void* const* x, const void* const* base,
                            void* const* new_base, void* const* packed, void* const* scale_u, void* const* scale_v,
                            int64_t N, int64_t C, void* workspace, size_t workspace_bytes, cf_stream_t stream) {
  const bool stable = (codec & CF_FLAG_INPUTS_STABLE) != 0;
  codec &= ~CF_FLAG_INPUTS_STABLE;
  CF_CHECK_ARG(codec == CF_CODEC_BINARY || codec == CF_CODEC_INT2, "codec must be CF_CODEC_BINARY or CF_CODEC_INT2");
  CF_CHECK_ARG(passes > 0 && (passes & ~CF_PASS_ALL) == 0, "bad pass mask %d", passes);
  if (codec == CF_CODEC_BINARY)
    return cf::sign_compress<cf::MODE_BINARY>(batch, x, base, new_base, packed, scale_u, scale_v, N, C, workspace,
                                              workspace_bytes, stream, passes, stable);
  return cf::sign_compress<cf::MODE_INT2>(batch, x, base, new_base, packed, scale_u, scale_v, N, C, workspace,
                                          workspace_bytes, stream, passes, stable);
}
int cf_sign_compress_put(int codec, int passes, int batch, const void* const* x, const void* const* base, int n_dst,
                         int self_dst, void* const* dst_payload, void* const* dst_flag, void* local_count,
                         void* local_ticket, int64_t N, int64_t C, void* workspace, size_t workspace_bytes,
                         cf_stream_t stream) {
  const bool stable = (codec & CF_FLAG_INPUTS_STABLE) != 0;
  codec &= ~CF_FLAG_INPUTS_STABLE;
  CF_CHECK_ARG(codec == CF_CODEC_BINARY || codec == CF_CODEC_INT2, "codec must be CF_CODEC_BINARY or CF_CODEC_INT2");
  CF_CHECK_ARG(passes > 0 && (passes & ~CF_PASS_ALL) == 0, "bad pass mask %d", passes);
  if (codec == CF_CODEC_BINARY)
    return cf::sign_compress_put<cf::MODE_BINARY>(passes, batch, x, base, n_dst, self_dst, dst_payload, dst_flag,
                                                  local_count, local_ticket, N, C, workspace, workspace_bytes, stream,
                                                  stable);
  return cf::sign_compress_put<cf::MODE_INT2>(passes, batch, x, base, n_dst, self_dst, dst_payload, dst_flag,
                                              local_count, local_ticket, N, C, workspace, workspace_bytes, stream,
                                              stable);
}
int cf_sign_decompress_batched_wait(int codec, int batch, const void* const* packed, const void* const* scale_u,
                                    const void* const* scale_v, const void* const* base, void* const* recon,
                                    const void* const* wait_flag, const void* expected, void* error_word, int64_t N,
                                    int64_t C, cf_stream_t stream) {
  const bool stable = (codec & CF_FLAG_INPUTS_STABLE) != 0;
  codec &= ~CF_FLAG_INPUTS_STABLE;
  CF_CHECK_ARG(codec == CF_CODEC_BINARY || codec == CF_CODEC_INT2, "codec must be CF_CODEC_BINARY or CF_CODEC_INT2");
  if (codec == CF_CODEC_BINARY)
    return cf::sign_decompress<cf::MODE_BINARY>(batch, packed, scale_u, scale_v, 1, base, recon, N, C, stream,
                                                wait_flag, expected, error_word, stable);
  return cf::sign_decompress<cf::MODE_INT2>(batch, packed, scale_u, scale_v, 1, base, recon, N, C, stream, wait_flag,
                                            expected, error_word, stable);
}
int cf_int2_encode_with_scales(const void* x, const void* base, const void* scale_u, const void* scale_v,
                               void* new_base, void* packed, int64_t N, int64_t C, cf_stream_t stream) {
  if (int rc = cf::check_shape(N, C, 1)) return rc;
  CF_CHECK_ARG(x && scale_u && scale_v && packed, "null pointer");
  CF_CHECK_ARG(cf::aligned16(x) && (!base || cf::aligned16(base)) && (!new_base || cf::aligned16(new_base)),
               "x/base/new_base must be 16-byte aligned");
  cf::Int2EncodeParams ep{};
  ep.N = static_cast<int>(N); ep.C = static_cast<int>(C);
  ep.x[0] = static_cast<const __half*>(x);
  ep.base[0] = static_cast<const __half*>(base);
  ep.scale_u[0] = static_cast<const __half*>(scale_u);
  ep.scale_v[0] = static_cast<const __half*>(scale_v);
  ep.packed[0] = static_cast<uint8_t*>(packed);
  ep.new_base[0] = static_cast<__half*>(new_base);
  return cf::launch_int2_encode(ep, 1, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
