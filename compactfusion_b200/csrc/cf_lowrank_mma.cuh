// Tensor-core kernels of the LOW_RANK projector (included by cf_lowrank.cu).
//
// The two skinny products of the subspace iteration, Y = A Q (N x r) and Z = A^T Y (C x r) with
// A = x - base, stream A once per pass: 2r flop per 2 bytes, below the B200 ridge, so the
// roofline is HBM.  The SIMT fp32 version was FMA-issue bound (ncu: 220-440 us per pass at
// 4608 x 3072); here the contraction runs on the tensor cores with `mma.sync.m16n8k8` TF32:
//   * A is fp16: its fp32 image is exactly representable in TF32, so A needs no splitting;
//   * the skinny operand (Q / Y, fp32) is pre-split into hi + lo TF32 terms (2 MMAs), which
//     keeps ~21 bits of it: the result is fp32-grade, not TF32-grade;
//   * A tiles land in shared memory with cp.async (x and base separately; the fp16 subtraction
//     happens on the ldmatrix fragments, one rounding like the reference's `x - base`), and
//     the fp16 m16n8k16 fragment is re-read as two TF32 k8 fragments (even / odd k), with the
//     skinny operand indexed to match.
// (tcgen05 would need the N x r result in TMEM with a 64-wide minimum N tile and buys nothing
// for a memory-bound contraction; mma.sync keeps the kernel small.)
#pragma once

#include "cf_common.cuh"

namespace cf {

constexpr int kLrBM = 128;       // M tile (rows of the output handled by a CTA): 8 warps x 16 rows
constexpr int kLrBK = 64;        // K chunk per pipeline stage
constexpr int kLrStages = 3;     // cp.async ring depth (2 chunks in flight while one is consumed)
constexpr int kLrThreads = kLrBM * 2;

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem));
  const int sz = valid ? 16 : 0;  // src-size 0: zero-fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float2 split_tf32(float v) {
  uint32_t hi, lo;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
  const float rest = v - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(rest));
  return make_float2(__uint_as_float(hi), __uint_as_float(lo));
}

// fp16 image of a skinny operand with entries of modest size (an orthonormal basis, the N(0, 1) start, or a Y scaled
// by a power of two, below): v ~= hi + lo * 2^-11 with hi = fp16(v), lo = fp16((v - hi) * 2^11): 22 bits, two fp16
// MMAs per term instead of four TF32 ones and no conversion of the streamed operand (k_lr_gemm<.., HB = true>).
// Stored as two planes, hi at [i] and lo at [count + i], in the buffer that otherwise holds the TF32 pairs.
__device__ __forceinline__ void split_h16_store(float v, __half* planes, size_t i, size_t count) {
  const __half hi = __float2half_rn(v);
  planes[i] = hi;
  planes[count + i] = __float2half_rn((v - __half2float(hi)) * 2048.f);
}

// out2[i] = split(sum_s part[s][i]);  optionally the fp32 sum as well.
// absmax != nullptr: out2 receives fp16 planes of the sum times 2^-e, e chosen from the bound
// S * max |partial| <= 2^14 * 2^e that the product kernel left in *absmax (fp32 bits of a non-negative number).
// A power of two common to the whole matrix is exact and changes neither span(A^T Y) nor what the
// orthonormalisation (scale-invariant: its shift is relative to the trace) returns.
__global__ void __launch_bounds__(256) k_lr_sum_split(const float* __restrict__ part, int S, size_t stride,
                                                     float2* __restrict__ out2, float* __restrict__ out32, size_t count,
                                                     const unsigned* __restrict__ absmax) {
  float scale = 1.f;
  if (absmax != nullptr) {
    const float bound = static_cast<float>(S) * __uint_as_float(*absmax);
    if (bound > 0.f && bound < 3.0e38f) scale = ldexpf(1.f, 13 - ilogbf(bound));   // bound * scale in [2^13, 2^14)
  }
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < S; ++p) s += part[static_cast<size_t>(p) * stride + i];
    if (out2 && absmax != nullptr)
      split_h16_store(s * scale, reinterpret_cast<__half*>(out2), i, count);
    else if (out2)
      out2[i] = split_tf32(s);
    if (out32) out32[i] = s;
  }
}

// one row of RP values held in registers -> both planes, 16 bytes per store (2-byte stores cost the orthonormalisation
// 6 us per call: 2 x RP store instructions per thread on 16 SMs)
template <int RP>
__device__ __forceinline__ void split_h16_store_row(const float (&xr)[RP], __half* planes, size_t row_off, size_t count) {
#pragma unroll
  for (int q = 0; q < RP / 8; ++q) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float v0 = xr[8 * q + 2 * w], v1 = xr[8 * q + 2 * w + 1];
      const __half2 hi = __floats2half2_rn(v0, v1);
      const float2 hf = __half22float2(hi);
      h[w] = h22u(hi);
      l[w] = h22u(__floats2half2_rn((v0 - hf.x) * 2048.f, (v1 - hf.y) * 2048.f));
    }
    *reinterpret_cast<uint4*>(planes + row_off + 8 * q) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(planes + count + row_off + 8 * q) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// (rows, r) compact fp32 -> (rows, RP) {hi, lo} pairs (or fp16 planes), zero padded
__global__ void __launch_bounds__(256) k_lr_pad_split(const float* __restrict__ src, float2* __restrict__ dst, int rows,
                                                     int r, int RP, int half_planes) {
  const size_t total = static_cast<size_t>(rows) * RP;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / RP), j = static_cast<int>(i % RP);
    const float v = (j < r) ? src[static_cast<size_t>(row) * r + j] : 0.f;
    if (half_planes)
      split_h16_store(v, reinterpret_cast<__half*>(dst), i, total);
    else
      dst[i] = split_tf32(v);
  }
}

// ---------------------------------------------------------------------------------------
// out[split][m][0..RP) = sum_{k in split} A'[m][k] * B[k][0..RP)
//   !TRANS: A' = A      (M = N, K = C)      Y = A Q
//    TRANS: A' = A^T    (M = C, K = N)      Z = A^T Y
// B2 = B pre-split into {hi, lo} TF32 pairs, (K, RP) row-major.   grid (ceil(M/128), splits), block 256
// ---------------------------------------------------------------------------------------
template <int RP, bool TRANS, int BK = kLrBK, int STAGES = kLrStages, bool HB = false>
__global__ void __launch_bounds__(2 * kLrThreads) k_lr_gemm(const __half* __restrict__ x, const __half* __restrict__ base,
                                                const float2* __restrict__ B2, float* __restrict__ out, int N, int C,
                                                int k_per_split, unsigned* __restrict__ absmax) {
  extern __shared__ __align__(128) unsigned char lr_smem_raw[];
  constexpr int kLdB = RP + 2;                               // float2 pitch: conflict-free 64-bit fragment loads
  // fp16 tile as stored in global memory: !TRANS 128 (m) x 64 (k), pitch 72;  TRANS 64 (k) x 128 (m), pitch 136
  constexpr int kRowsA = TRANS ? BK : kLrBM, kColsA = TRANS ? kLrBM : BK;
  constexpr int kLdA = kColsA + 8;
  constexpr int kTileA = kRowsA * kLdA * 2;                  // bytes
  constexpr int kLdH = RP + 8;                               // HB: fp16 pitch of a B plane (16-byte rows, conflict-free ldmatrix)
  constexpr int kTileB = HB ? 2 * BK * kLdH * 2 : BK * kLdB * 8;
  constexpr int kStage = 2 * kTileA + kTileB;
  const int M = TRANS ? C : N, K = TRANS ? N : C;
  const int m0 = blockIdx.x * kLrBM;
  const int k_begin = blockIdx.y * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  // 8 warps cover the 128 rows of the tile; a block of 16 warps (RP >= 16) splits the RP / 8 column tiles between
  // two warp groups: twice the warps to hide the ldmatrix -> convert -> MMA latency chain, half the B fragment
  // loads per warp (ncu, round 2: 12 % warps active, stalls on fixed-latency dependencies)
  const int tid = threadIdx.x, lane = tid & 31;
  const int nthreads = blockDim.x;
  const int warp = (tid >> 5) & 7, warp_n = tid >> 8;
  const int j_begin = (nthreads > kLrThreads) ? warp_n * (RP / 16) : 0;
  const int j_end = (nthreads > kLrThreads) ? j_begin + RP / 16 : RP / 8;
  const int g = lane >> 2, t = lane & 3;
  const bool has_base = base != nullptr;

  auto stage_ptr = [&](int s) { return lr_smem_raw + s * kStage; };
  auto issue = [&](int s, int k0) {
    unsigned char* sp = stage_ptr(s);
    __half* xs = reinterpret_cast<__half*>(sp);
    __half* bs = reinterpret_cast<__half*>(sp + kTileA);
    float2* Bs = reinterpret_cast<float2*>(sp + 2 * kTileA);
    constexpr int kChunkCols = kColsA / 8;
    for (int ch = tid; ch < kRowsA * kChunkCols; ch += nthreads) {  // 16-byte chunks of the fp16 tile
      const int r = ch / kChunkCols, cc = ch % kChunkCols;
      const int grow = TRANS ? (k0 + r) : (m0 + r);
      const int gcol = TRANS ? (m0 + 8 * cc) : (k0 + 8 * cc);
      const bool ok = grow < N && gcol < C && (TRANS ? (grow < k_end) : (gcol < k_end));
      const size_t off = ok ? (static_cast<size_t>(grow) * C + gcol) : 0;
      cp_async16(xs + r * kLdA + 8 * cc, x + off, ok);
      if (has_base) cp_async16(bs + r * kLdA + 8 * cc, base + off, ok);
    }
    if (HB) {   // two fp16 planes (hi | lo) of the (K, RP) operand, `K * RP` halves apart
      constexpr int kChunksH = BK * (RP / 8);
      const __half* Bg = reinterpret_cast<const __half*>(B2);
      __half* Bh = reinterpret_cast<__half*>(sp + 2 * kTileA);
      for (int ch = tid; ch < 2 * kChunksH; ch += nthreads) {
        const int plane = ch / kChunksH, rem = ch % kChunksH;
        const int r = rem / (RP / 8), cc = rem % (RP / 8);
        const bool ok = k0 + r < k_end;
        const size_t off = ok ? (static_cast<size_t>(plane) * K * RP + static_cast<size_t>(k0 + r) * RP + 8 * cc) : 0;
        cp_async16(Bh + (plane * BK + r) * kLdH + 8 * cc, Bg + off, ok);
      }
      return;
    }
    constexpr int kChunksB = BK * RP / 2;  // 16-byte chunks = 2 float2
    for (int ch = tid; ch < kChunksB; ch += nthreads) {
      const int r = ch / (RP / 2), cc = ch % (RP / 2);
      const bool ok = k0 + r < k_end;
      const size_t off = ok ? (static_cast<size_t>(k0 + r) * RP + 2 * cc) : 0;
      cp_async16(Bs + r * kLdB + 2 * cc, B2 + off, ok);
    }
  };

  float acc[RP / 8][4];
  float accl[HB ? RP / 8 : 1][4];   // HB: the lo plane's products, scaled by 2^11
#pragma unroll
  for (int j = 0; j < RP / 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      acc[j][e] = 0.f;
      if (HB) accl[HB ? j : 0][e] = 0.f;
    }

  const int nchunks = (k_end - k_begin + BK - 1) / BK;
#pragma unroll
  for (int s0 = 0; s0 < STAGES - 1; ++s0) {  // prologue: always commit, so group accounting stays uniform
    if (s0 < nchunks) issue(s0, k_begin + s0 * BK);
    cp_async_commit();
  }
  for (int it = 0; it < nchunks; ++it) {
    cp_async_wait<STAGES - 2>();  // chunk `it` has landed (for this thread's copies)
    __syncthreads();                 // ... for everyone's; and everyone is done with the stage refilled below
    {
      const int nx = it + STAGES - 1;
      if (nx < nchunks) issue(nx % STAGES, k_begin + nx * BK);
      cp_async_commit();
    }
    unsigned char* sp = stage_ptr(it % STAGES);
    const __half* xs = reinterpret_cast<const __half*>(sp);
    const __half* bs = reinterpret_cast<const __half*>(sp + kTileA);
    const float2* Bs = reinterpret_cast<const float2*>(sp + 2 * kTileA);
#pragma unroll
    for (int kk = 0; kk < BK / 16; ++kk) {
      uint32_t fx[4], fb[4];
      const int mi = lane >> 3, l8 = lane & 7;
      int srow, scol;
      if (!TRANS) {
        srow = 16 * warp + l8 + (mi & 1) * 8;
        scol = 16 * kk + (mi >> 1) * 8;
        ldmatrix_x4(fx, xs + srow * kLdA + scol);
        if (has_base) ldmatrix_x4(fb, bs + srow * kLdA + scol);
      } else {
        srow = 16 * kk + l8 + (mi >> 1) * 8;
        scol = 16 * warp + (mi & 1) * 8;
        ldmatrix_x4_trans(fx, xs + srow * kLdA + scol);
        if (has_base) ldmatrix_x4_trans(fb, bs + srow * kLdA + scol);
      }
      if (HB) {
        // fp16 m16n8k16: the delta fragment is the A operand as it is; B fragments of two adjacent column tiles
        // per ldmatrix.trans from the [k][n] planes
        uint32_t a[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q] = has_base ? h22u(__hsub2_rn(u2h2(fx[q]), u2h2(fb[q]))) : fx[q];
        const __half* Bh = reinterpret_cast<const __half*>(sp + 2 * kTileA);
        const __half* brow_h = Bh + (16 * kk + l8 + (mi & 1) * 8) * kLdH + (mi >> 1) * 8;
#pragma unroll
        for (int j2 = 0; j2 < RP / 16; ++j2) {
          uint32_t bh[4], bl[4];
          ldmatrix_x4_trans(bh, brow_h + 16 * j2);
          ldmatrix_x4_trans(bl, brow_h + BK * kLdH + 16 * j2);
          mma_f16(acc[2 * j2], a, bh[0], bh[1]);
          mma_f16(acc[2 * j2 + 1], a, bh[2], bh[3]);
          mma_f16(accl[HB ? 2 * j2 : 0], a, bl[0], bl[1]);
          mma_f16(accl[HB ? 2 * j2 + 1 : 0], a, bl[2], bl[3]);
        }
        continue;
      }
      uint32_t a_even[4], a_odd[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        __half2 d = u2h2(fx[q]);
        if (has_base) d = __hsub2_rn(d, u2h2(fb[q]));  // one fp16 rounding, like `x - base` (main.py:229)
        const float2 f = __half22float2(d);            // exact; a valid TF32 bit pattern
        a_even[q] = __float_as_uint(f.x);
        a_odd[q] = __float_as_uint(f.y);
      }
      const float2* brow = Bs + (16 * kk + 2 * t) * kLdB + g;
      // the 4 MMAs into one accumulator are a dependent chain (the wrappers keep program order): issue them
      // across the RP / 8 independent accumulators, term by term, so consecutive MMAs never wait on each other
      float2 be0[RP / 8], be1[RP / 8], bo0[RP / 8], bo1[RP / 8];
#pragma unroll
      for (int j = 0; j < RP / 8; ++j) {
        if (j < j_begin || j >= j_end) continue;   // warp-uniform: the other warp group's column tiles
        be0[j] = brow[8 * j];
        be1[j] = brow[8 * kLdB + 8 * j];
        bo0[j] = brow[kLdB + 8 * j];
        bo1[j] = brow[9 * kLdB + 8 * j];
      }
#pragma unroll
      for (int j = 0; j < RP / 8; ++j)
        if (j >= j_begin && j < j_end) mma_tf32(acc[j], a_even, __float_as_uint(be0[j].x), __float_as_uint(be1[j].x));
#pragma unroll
      for (int j = 0; j < RP / 8; ++j)
        if (j >= j_begin && j < j_end) mma_tf32(acc[j], a_even, __float_as_uint(be0[j].y), __float_as_uint(be1[j].y));
#pragma unroll
      for (int j = 0; j < RP / 8; ++j)
        if (j >= j_begin && j < j_end) mma_tf32(acc[j], a_odd, __float_as_uint(bo0[j].x), __float_as_uint(bo1[j].x));
#pragma unroll
      for (int j = 0; j < RP / 8; ++j)
        if (j >= j_begin && j < j_end) mma_tf32(acc[j], a_odd, __float_as_uint(bo0[j].y), __float_as_uint(bo1[j].y));
    }
  }
  if (HB) {
#pragma unroll
    for (int j = 0; j < RP / 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[j][e] = fmaf(accl[HB ? j : 0][e], 1.0f / 2048.f, acc[j][e]);
  }
  if (absmax != nullptr) {
    // largest |partial| of the launch (fp32 bits of non-negative numbers order like unsigned integers): the bound
    // k_lr_sum_split scales the sum by before it splits it into fp16 planes
    __shared__ unsigned cta_max;
    if (tid == 0) cta_max = 0u;
    __syncthreads();
    float mx = 0.f;
#pragma unroll
    for (int j = 0; j < RP / 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) mx = fmaxf(mx, fabsf(acc[j][e]));
#pragma unroll
    for (int o_ = 16; o_ > 0; o_ >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o_));
    if (lane == 0) atomicMax(&cta_max, __float_as_uint(mx));
    __syncthreads();
    if (tid == 0) atomicMax(absmax, cta_max);
  }
  float* o = out + static_cast<size_t>(blockIdx.y) * M * RP;
  const int r0 = m0 + 16 * warp + g;
#pragma unroll
  for (int j = 0; j < RP / 8; ++j) {
    if (j < j_begin || j >= j_end) continue;
    const int col = 8 * j + 2 * t;
    if (r0 < M) *reinterpret_cast<float2*>(o + static_cast<size_t>(r0) * RP + col) = make_float2(acc[j][0], acc[j][1]);
    if (r0 + 8 < M)
      *reinterpret_cast<float2*>(o + static_cast<size_t>(r0 + 8) * RP + col) = make_float2(acc[j][2], acc[j][3]);
  }
}

template <int RP, bool TRANS, int BK = kLrBK, int STAGES = kLrStages, bool HB = false>
constexpr size_t lr_gemm_smem() {
  constexpr size_t rows = TRANS ? BK : kLrBM, cols = TRANS ? kLrBM : BK;
  constexpr size_t tile_b = HB ? 2 * static_cast<size_t>(BK) * (RP + 8) * 2 : static_cast<size_t>(BK) * (RP + 2) * 8;
  return STAGES * (2 * rows * (cols + 8) * 2 + tile_b);
}

// ---------------------------------------------------------------------------------------
// CholeskyQR building blocks.
//   k_lr_gram_chol: X <- sum_s Xpart[s];  G = X^T X in fp64 (per-CTA partials, <= 16 CTAs); the CTA that
//                   draws the last ticket adds the partials in a fixed order and factors G = R^T R (fp64);
//                   R (upper, fp32) and 1 / diag(R) go to global memory.
//   k_lr_solve_out: X <- X R^{-1} by forward substitution, one row per thread (no explicit inverse: its
//                   back substitution was a serial fp64 chain of r^2 / 2 steps on one CTA); optional outputs
//                   as {hi,lo} TF32 pairs (the next product's skinny operand), fp16 and compact fp32.
// ---------------------------------------------------------------------------------------
constexpr int kLrMaxRank = 64;

// Upper-triangular R with G = R^T R by all 256 threads of the CTA, one barrier per step: the trailing
// update uses the unscaled pivot row, G[i][j] -= G[k][i] G[k][j] / G[k][k], so row k never has to be scaled
// in place first; thread (ti, tj) of a 16 x 16 grid owns the entries (ti + 16a, tj + 16b).
// On return G[k][j] (j >= k) holds the unscaled pivot rows: R[k][j] = G[k][j] * piv[k], piv[k] = G[k][k]^-1/2.
// 1/d and d^-1/2 in fp64 from fp32 seeds + two Newton steps each (the library double division / rsqrt are
// ~40-instruction dependent chains that sat on the critical path of every elimination step)
__device__ __forceinline__ double fast_rcp(double d) {
  if (!(d > 1e-30 && d < 1e30)) return 1.0 / d;  // outside the fp32 seed's comfortable range
  double y = static_cast<double>(__frcp_rn(static_cast<float>(d)));
  y = y * (2.0 - d * y);
  return y * (2.0 - d * y);
}
__device__ __forceinline__ double fast_rsqrt(double d) {
  if (!(d > 1e-30 && d < 1e30)) return rsqrt(d);
  double y = static_cast<double>(rsqrtf(static_cast<float>(d)));
  y = y * (1.5 - 0.5 * d * y * y);
  return y * (1.5 - 0.5 * d * y * y);
}

__device__ void cholesky_upper_256(double (*G)[kLrMaxRank + 1], double* piv, int r, int t) {
  const int ti = t >> 4, tj = t & 15;
  double maxdiag = 0.0;
  for (int i = 0; i < r; ++i) maxdiag = fmax(maxdiag, G[i][i]);
  const double floor_piv = fmax(maxdiag, 1e-300) * 1e-14;
  __syncthreads();
  for (int k = 0; k < r; ++k) {
    double d = G[k][k];
    if (!(d > floor_piv)) d = floor_piv;  // rank-deficient input: keep things finite
    // (d is within fp32 range: a squared column norm of fp16-derived data)
    const double inv_d = fast_rcp(d);
    if (t == 0) piv[k] = fast_rsqrt(d);
#pragma unroll
    for (int a = 0; a < kLrMaxRank / 16; ++a) {
      const int i = ti + 16 * a;
      if (i > k && i < r) {
        const double gki = G[k][i] * inv_d;
#pragma unroll
        for (int b = 0; b < kLrMaxRank / 16; ++b) {
          const int j = tj + 16 * b;
          if (j >= i && j < r) G[i][j] -= gki * G[k][j];
        }
      }
    }
    __syncthreads();
  }
}

struct GramParams {
  const float* xpart;   // S partial copies of X, `part_stride` floats apart (S = 1: X itself)
  int S;
  size_t part_stride;
  float* X;             // (M, RP) sum of the partials (written when S > 1 or X != xpart)
  int M, r, rows_per_cta;
  double* gpart;        // (gridDim.x, r*r)
  unsigned* ticket;
  float* r_out;         // (RP, RP) upper-triangular factor, row-major, zero padded
  float* rdinv_out;     // (RP) 1 / diag
};

template <int RP>
__global__ void __launch_bounds__(256) k_lr_gram_chol(const GramParams p) {
  // one raw buffer: the fp64 row tile while streaming, then G for the factorisation (static smem <= 48 KB)
  constexpr int kLdX = RP + 2;  // doubles; even: 16-byte aligned pairs
  constexpr size_t kStream = sizeof(double) * 32 * kLdX;
  constexpr size_t kFactor = sizeof(double) * (kLrMaxRank * (kLrMaxRank + 1) + kLrMaxRank);
  __shared__ __align__(16) unsigned char raw[kStream > kFactor ? kStream : kFactor];
  double* Xd = reinterpret_cast<double*>(raw);  // [32][kLdX]
  __shared__ bool is_last;
  const int t = threadIdx.x, r = p.r;
  const int m_begin = blockIdx.x * p.rows_per_cta;
  const int m_end = min(p.M, m_begin + p.rows_per_cta);
  const bool write_x = p.S > 1 || p.X != p.xpart;
  // thread (ti, tj) of a 16 x 16 grid owns the BxB block of Gram entries (B*ti + a, B*tj + b), B = RP / 16
  constexpr int B = RP >= 16 ? RP / 16 : 1;
  constexpr int TG = RP >= 16 ? 16 : RP;  // RP = 8: an 8 x 8 grid of threads, one entry each
  const int ti = t / TG, tj = t % TG;
  const bool active = ti < TG;
  double acc[B][B];
#pragma unroll
  for (int a = 0; a < B; ++a)
#pragma unroll
    for (int b = 0; b < B; ++b) acc[a][b] = 0.0;
  for (int mc = m_begin; mc < m_end; mc += 32) {
    __syncthreads();
    for (int i = t; i < 32 * RP; i += 256) {
      const int rr = i / RP, j = i % RP;
      float s = 0.f;
      if (mc + rr < m_end) {
        const size_t o = static_cast<size_t>(mc + rr) * RP + j;
        for (int sidx = 0; sidx < p.S; ++sidx) s += p.xpart[static_cast<size_t>(sidx) * p.part_stride + o];
        if (write_x) p.X[o] = s;
      }
      Xd[rr * kLdX + j] = static_cast<double>(s);
    }
    __syncthreads();
    if (active) {
#pragma unroll 4
      for (int rr = 0; rr < 32; ++rr) {
        double xi[B], xj[B];
#pragma unroll
        for (int a = 0; a < B; ++a) xi[a] = Xd[rr * kLdX + B * ti + a];
#pragma unroll
        for (int b = 0; b < B; ++b) xj[b] = Xd[rr * kLdX + B * tj + b];
#pragma unroll
        for (int a = 0; a < B; ++a)
#pragma unroll
          for (int b = 0; b < B; ++b) acc[a][b] = fma(xi[a], xj[b], acc[a][b]);
      }
    }
  }
  if (active) {
#pragma unroll
    for (int a = 0; a < B; ++a)
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const int i = B * ti + a, j = B * tj + b;
        if (i < r && j < r) p.gpart[static_cast<size_t>(blockIdx.x) * r * r + i * r + j] = acc[a][b];
      }
  }
  __syncthreads();
  if (t == 0) {
    __threadfence();
    const unsigned ticket = atomicAdd(p.ticket, 1u);
    is_last = ticket == gridDim.x - 1;
    if (is_last) *p.ticket = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // last CTA: G = sum of the partials (fixed order), Cholesky
  double (*G)[kLrMaxRank + 1] = reinterpret_cast<double (*)[kLrMaxRank + 1]>(raw);
  double* piv = reinterpret_cast<double*>(raw) + kLrMaxRank * (kLrMaxRank + 1);
  for (int e = t; e < r * r; e += 256) {
    double s = 0.0;
#pragma unroll 4
    for (unsigned b = 0; b < gridDim.x; ++b) s += p.gpart[static_cast<size_t>(b) * r * r + e];
    G[e / r][e % r] = s;
  }
  __syncthreads();
  cholesky_upper_256(G, piv, r, t);
  for (int e = t; e < RP * RP; e += 256) {
    const int i = e / RP, j = e % RP;
    p.r_out[e] = (i < r && j < r && j >= i) ? static_cast<float>(G[i][j] * piv[i]) : 0.f;
  }
  for (int j = t; j < RP; j += 256) p.rdinv_out[j] = (j < r) ? static_cast<float>(piv[j]) : 0.f;
}

// X <- X R^{-1} (right-looking forward substitution, one row per thread); block 128
template <int RP>
__global__ void __launch_bounds__(128) k_lr_solve_out(float* __restrict__ X, const float* __restrict__ R,
                                                     const float* __restrict__ rdinv, int M, int r,
                                                     float2* __restrict__ out2, __half* __restrict__ out16,
                                                     float* __restrict__ out32c, int half_planes) {
  __shared__ __align__(16) float Rs[RP * RP];
  __shared__ float Ds[RP];
  for (int e = threadIdx.x; e < RP * RP; e += 128) Rs[e] = R[e];
  for (int e = threadIdx.x; e < RP; e += 128) Ds[e] = rdinv[e];
  __syncthreads();
  const int m = blockIdx.x * 128 + threadIdx.x;
  if (m >= M) return;
  float xr[RP];
  float4* xrow = reinterpret_cast<float4*>(X + static_cast<size_t>(m) * RP);
#pragma unroll
  for (int q = 0; q < RP / 4; ++q) {
    const float4 v = xrow[q];
    xr[4 * q] = v.x; xr[4 * q + 1] = v.y; xr[4 * q + 2] = v.z; xr[4 * q + 3] = v.w;
  }
#pragma unroll
  for (int i = 0; i < RP; ++i) {
    const float xi = xr[i] * Ds[i];  // columns >= r: Ds = 0 -> exact zeros in the padding
    xr[i] = xi;
#pragma unroll
    for (int j = i + 1; j < RP; ++j) xr[j] = fmaf(-xi, Rs[i * RP + j], xr[j]);
  }
#pragma unroll
  for (int q = 0; q < RP / 4; ++q) xrow[q] = make_float4(xr[4 * q], xr[4 * q + 1], xr[4 * q + 2], xr[4 * q + 3]);
  if (out2 && half_planes) {
    split_h16_store_row<RP>(xr, reinterpret_cast<__half*>(out2), static_cast<size_t>(m) * RP, static_cast<size_t>(M) * RP);
  } else if (out2) {
#pragma unroll
    for (int j = 0; j < RP; ++j) out2[static_cast<size_t>(m) * RP + j] = split_tf32(xr[j]);
  }
#pragma unroll
  for (int j = 0; j < RP; ++j)
    if (j < r) {
      if (out16) out16[static_cast<size_t>(m) * r + j] = __float2half_rn(xr[j]);
      if (out32c) out32c[static_cast<size_t>(m) * r + j] = xr[j];
    }
}

// ---------------------------------------------------------------------------------------
// recon = base + fp16(U V) on the tensor cores (fp16 m16n8k16, fp32 accumulate): replaces
// torch.matmul(u, v) + add (slowpath.py:152-154, main.py:376).  CTA tile 64 rows x 256 columns,
// warp w owns rows [16w, 16w+16); K = rank padded to KS*16 with zeros.  The fp16 product tile is
// staged through shared memory so that base is read and recon written in 16-byte row segments.
// grid (ceil(C/256), ceil(N/64)), block 128
// ---------------------------------------------------------------------------------------

template <int KS>
__global__ void __launch_bounds__(128) k_lr_reconstruct_mma(const __half* __restrict__ U, const __half* __restrict__ V,
                                                           const __half* __restrict__ base, __half* __restrict__ recon,
                                                           int N, int C, int r) {
  constexpr int KP = KS * 16, BN = 256;
  constexpr int kLdU = KP + 8, kLdV = BN + 8;
  __shared__ __align__(16) __half Us[64 * kLdU];
  __shared__ __align__(16) __half Vs[(KP > 64 ? KP : 64) * kLdV];   // reused as the fp16 product tile [64][kLdV]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int c0 = blockIdx.x * BN, n0 = blockIdx.y * 64;
  const __half zero = __float2half_rn(0.f);
  for (int i = tid; i < 64 * KP; i += 128) {
    const int rr = i / KP, k = i % KP;
    Us[rr * kLdU + k] = (n0 + rr < N && k < r) ? U[static_cast<size_t>(n0 + rr) * r + k] : zero;
  }
  for (int i = tid; i < KP * (BN / 8); i += 128) {
    const int k = i / (BN / 8), cc = i % (BN / 8);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (k < r && c0 + 8 * cc < C) v = *reinterpret_cast<const uint4*>(V + static_cast<size_t>(k) * C + c0 + 8 * cc);
    *reinterpret_cast<uint4*>(Vs + k * kLdV + 8 * cc) = v;
  }
  __syncthreads();
  float acc[BN / 8][4];
#pragma unroll
  for (int j = 0; j < BN / 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
  const int mi = lane >> 3, l8 = lane & 7;
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) {
    uint32_t a[4];
    ldmatrix_x4(a, Us + (16 * warp + l8 + (mi & 1) * 8) * kLdU + 16 * kk + (mi >> 1) * 8);
#pragma unroll
    for (int j2 = 0; j2 < BN / 16; ++j2) {
      // stored [k][n]: .trans yields the col-major B fragments of two adjacent n-tiles
      uint32_t b[4];
      ldmatrix_x4_trans(b, Vs + (16 * kk + l8 + (mi & 1) * 8) * kLdV + 16 * j2 + (mi >> 1) * 8);
      mma_f16(acc[2 * j2], a, b[0], b[1]);
      mma_f16(acc[2 * j2 + 1], a, b[2], b[3]);
    }
  }
  __syncthreads();  // all warps are done reading Vs: reuse it for the product tile
  __half* Ps = Vs;  // [64][kLdV]
#pragma unroll
  for (int j = 0; j < BN / 8; ++j) {
    const int col = 8 * j + 2 * t;
    *reinterpret_cast<__half2*>(Ps + (16 * warp + g) * kLdV + col) = __floats2half2_rn(acc[j][0], acc[j][1]);
    *reinterpret_cast<__half2*>(Ps + (16 * warp + g + 8) * kLdV + col) = __floats2half2_rn(acc[j][2], acc[j][3]);
  }
  __syncwarp();  // a warp only re-reads its own 16 rows
#pragma unroll 4
  for (int i = lane; i < 16 * (BN / 8); i += 32) {
    const int rr = 16 * warp + i / (BN / 8), cc = i % (BN / 8);
    const int n = n0 + rr, c = c0 + 8 * cc;
    if (n >= N || c >= C) continue;
    const uint4 pv = *reinterpret_cast<const uint4*>(Ps + rr * kLdV + 8 * cc);
    const size_t off = static_cast<size_t>(n) * C + c;
    uint4 out = pv;
    if (base != nullptr) {
      const H8 b = as_h8(ldg_stream(base + off)), pr = as_h8(pv);
      H8 o;
#pragma unroll
      for (int q = 0; q < 4; ++q) o.w[q] = h22u(__hadd2_rn(u2h2(b.w[q]), u2h2(pr.w[q])));  // base + recv_delta
      out = as_u4(o);
    }
    stg_stream(recon + off, out);
  }
}

// V (r, C) fp16 = (sum_s Vt_part[s] (C, RP))^T
__global__ void __launch_bounds__(256) k_lr_store_v_sum(const float* __restrict__ part, int S, size_t stride,
                                                       __half* __restrict__ V, int C, int RP, int r) {
  const size_t total = static_cast<size_t>(r) * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i / C), c = static_cast<int>(i % C);
    float s = 0.f;
    for (int p = 0; p < S; ++p) s += part[static_cast<size_t>(p) * stride + static_cast<size_t>(c) * RP + k];
    V[i] = __float2half_rn(s);
  }
}

}  // namespace cf
