// Cluster orthonormalisation (shifted CholeskyQR3) and the prefetching reconstruct of the LOW_RANK codec (included
// by cf_lowrank.cu after cf_lowrank_mma.cuh, whose split_tf32 / ldmatrix / mma wrappers they use).
#pragma once

#include "cf_lowrank_mma.cuh"

namespace cf {

// ---------------------------------------------------------------------------------------
// Orthonormalisation of X = sum of S split-K partials (M x RP, fp32) as ONE kernel on ONE thread-block cluster:
// shifted CholeskyQR3, entirely in fp32.
//
// Round 1 measured 93 us per orthonormalisation (sum kernel + 2 x (fp64 Gram / Cholesky on a 16-CTA grid with a
// global ticket + a forward-substitution kernel)): five dependent launches, three times per projector call, 2/3
// of the call.  Two cluster versions with an fp64 Gram / fp64 factorisation followed (109 and 79 us,
// profiles/r2_kernel_times_lowrank_orth_v{1,2}.md): this part issues about 2.4 fp64 operations per clock per
// SM, so ANY fp64 in the loop dominates.  CholeskyQR2 needs the fp64 Gram because it fails once
// cond(X)^2 u32 ~ 1; shifted CholeskyQR3 (Fukaya, Kannan, Nakatsukasa, Yamamoto, Yanagisawa 2020) does not:
//   pass 1:  R1 = chol(X^T X + s I),  X1 = X R1^-1     the shift makes the factorisation succeed for any X and
//                                                       leaves cond(X1) <= sqrt(1 + sigma_1^2 / s)
//   pass 2, 3: plain CholeskyQR on X1 (cond <= 100 with s = 1e-4 trace(X^T X)): orthogonal to fp32 rounding
// Every pass spans the same column space as X, so U V -- the only quantity the codec ships -- is unchanged.
//
// The kClusterCtas CTAs each own a contiguous row range of X (the matrix is < 1 MB: L2-resident) and stream it
// once per pass; partial Grams (fp32, Kahan-compensated over blocks of 32 rows) are reduced over the cluster
// through distributed shared memory in a FIXED rank order (identical bits in every CTA, no atomics) and every CTA
// factors the r x r matrix redundantly, so nothing goes back to global memory or to the host between passes.
// ---------------------------------------------------------------------------------------
constexpr int kClusterCtas = 8;    // portable cluster size; 16 (non-portable, opt-in attribute) halves the rows per CTA
// rows staged per step: one row per thread in the substitution passes (RP = 64: half, to fit shared memory)
// 512 threads per CTA: the Gram blocks are dealt to the first 256 thread slots and the rows of a chunk to two row
// groups (slot, group), so the substitution has one row per thread for up to 512 rows and twice the loads in flight
constexpr int kOrthThreads = 512;
template <int RP>
constexpr int orth_chunk() { return RP <= 32 ? 512 : 256; }

struct OrthParams {
  const float* part;     // S partial copies of X, `part_stride` floats apart
  int S;
  size_t part_stride;
  float* X;              // (M, RP) scratch: the sum, then the intermediate X R^-1
  int M, r;
  float2* out2;          // optional (M, RP) {hi, lo} TF32 pairs (the next product's skinny operand)
  __half* out16;         // optional (M, r) fp16
  float* out32c;         // optional (M, r) compact fp32
  int half_planes;       // out2 as two fp16 planes (hi | lo * 2^11) instead of TF32 pairs (split_h16_store)
  int light;             // 1: X only seeds the next product (an intermediate basis of the subspace iteration): any
                         // well-conditioned basis of span(X) gives the same next subspace, so ONE factorisation is
                         // enough when X is well conditioned (orthogonality error ~cond^2 u ~ 2e-4), two otherwise
};

template <int RP>
constexpr size_t lr_orth_smem() {
  return sizeof(float) * orth_chunk<RP>() * (RP + 2)      // row chunk
         + sizeof(float) * 2 * RP * RP                    // this CTA's partial Gram (read by the cluster) + row group 1's
         + sizeof(float) * (2 * RP * (RP + 1) + RP)       // G, its copy (+ pivots) for the factorisation
         + sizeof(float) * (RP * RP + RP) + 64;           // R (upper), 1 / diag, scalars
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// read a float from the same shared-memory offset of CTA `rank` of the cluster (DSMEM)
__device__ __forceinline__ float ld_dsmem_f32(const float* local, uint32_t rank) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(local));
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(a), "r"(rank));
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
  return v;
}

// Upper-triangular R with G = R^T R by the 256 threads of the CTA, fp32, one barrier per step: the trailing update
// uses the unscaled pivot row, G[i][j] -= G[k][i] G[k][j] / G[k][k]; thread (ti, tj) of a 16 x 16 grid owns the
// entries (ti + 16a, tj + 16b).  On return G[k][j] (j >= k) holds the unscaled pivot rows:
// R[k][j] = G[k][j] * piv[k], piv[k] = G[k][k]^-1/2.  A pivot below `floor_piv` (rank-deficient input) is floored.
template <int RP>
__device__ void cholesky_upper_f32(float (*G)[RP + 1], float* piv, int r, int t, float floor_piv) {
  const int ti = t >> 4, tj = t & 15;
  const bool active = t < 256;   // (threads beyond the 16 x 16 grid only keep the barriers company)
  for (int k = 0; k < r; ++k) {
    float d = G[k][k];
    if (!(d > floor_piv)) d = floor_piv;
    const float inv_d = 1.0f / d;
    if (t == 0) piv[k] = rsqrtf(d);
#pragma unroll
    for (int a = 0; a < (RP + 15) / 16; ++a) {
      const int i = ti + 16 * a;
      if (active && i > k && i < r) {
        const float gki = G[k][i] * inv_d;
#pragma unroll
        for (int b = 0; b < (RP + 15) / 16; ++b) {
          const int j = tj + 16 * b;
          if (j >= i && j < r) G[i][j] = fmaf(-gki, G[k][j], G[i][j]);
        }
      }
    }
    __syncthreads();
  }
}

template <int RP, int NC>
__global__ void __launch_bounds__(kOrthThreads, 1) k_lr_orth(const OrthParams p) {
  extern __shared__ __align__(16) unsigned char orth_raw[];
  constexpr int kLdX = RP + 2, kOrthChunk = orth_chunk<RP>();
  float* Xd = reinterpret_cast<float*>(orth_raw);                         // [kOrthChunk][kLdX]
  float* Gp = Xd + kOrthChunk * kLdX;                                     // [RP][RP] partial Gram of this CTA
  float* Gp2 = Gp + RP * RP;                                              // [RP][RP] row group 1's share of it
  float (*G)[RP + 1] = reinterpret_cast<float (*)[RP + 1]>(Gp2 + RP * RP);
  float (*Gc)[RP + 1] = reinterpret_cast<float (*)[RP + 1]>(reinterpret_cast<float*>(G) + RP * (RP + 1));
  float* piv = reinterpret_cast<float*>(Gc) + RP * (RP + 1);
  float* Rs = piv + RP;                                                   // [RP][RP]
  float* Ds = Rs + RP * RP;                                               // [RP]
  float* scal = Ds + RP;                                                  // [0]: trace(G), [1]: well-conditioned?
  const int t = threadIdx.x, r = p.r, M = p.M;
  const uint32_t rank = cluster_ctarank();
  int rows_per = (M + NC - 1) / NC;
  rows_per = (rows_per + 3) / 4 * 4;
  const int m_begin = min(M, static_cast<int>(rank) * rows_per);
  const int m_end = min(M, m_begin + rows_per);

  // upper-triangular 2x2 blocks of the Gram matrix, dealt to the threads round-robin
  constexpr int NB = RP / 2, kBlocks = NB * (NB + 1) / 2, KB = (kBlocks + 255) / 256;
  int bi[KB], bj[KB];
#pragma unroll
  for (int q = 0; q < KB; ++q) {
    int id = (t & 255) + 256 * q, i = 0;
    if (id >= kBlocks) id = -1;
    if (id >= 0)
      while (id >= NB - i) { id -= NB - i; ++i; }   // row i of the block triangle holds NB - i blocks
    bi[q] = id >= 0 ? i : -1;
    bj[q] = id >= 0 ? i + id : -1;
  }
  const int rg = t >> 8;             // row group: 32-row blocks rg, rg + 2, ... of a chunk
  float acc[KB][4], comp[KB][4];   // Kahan sum of the 32-row block sums

  auto zero_acc = [&]() {
#pragma unroll
    for (int q = 0; q < KB; ++q)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[q][e] = comp[q][e] = 0.f;
  };
  auto kahan = [&](float& sum, float& c, float v) {
    const float y = v - c;
    const float tsum = sum + y;
    c = (tsum - sum) - y;
    sum = tsum;
  };
  auto gram_chunk = [&](int rows) {   // acc += Xd[0..rows)^T Xd[0..rows) on this thread's blocks
    for (int rb = 32 * rg; rb < rows; rb += 32 * (kOrthThreads / 256)) {
      const int re = min(rows, rb + 32);
#pragma unroll
      for (int q = 0; q < KB; ++q) {
        if (bi[q] >= 0) {
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
          for (int rr = rb; rr < re; ++rr) {
            const float2 xi = *reinterpret_cast<const float2*>(Xd + rr * kLdX + 2 * bi[q]);
            const float2 xj = *reinterpret_cast<const float2*>(Xd + rr * kLdX + 2 * bj[q]);
            a0 = fmaf(xi.x, xj.x, a0);
            a1 = fmaf(xi.x, xj.y, a1);
            a2 = fmaf(xi.y, xj.x, a2);
            a3 = fmaf(xi.y, xj.y, a3);
          }
          kahan(acc[q][0], comp[q][0], a0);
          kahan(acc[q][1], comp[q][1], a1);
          kahan(acc[q][2], comp[q][2], a2);
          kahan(acc[q][3], comp[q][3], a3);
        }
      }
    }
  };
  // partial Gram -> Gp; cluster-wide sum in rank order -> G (upper); (+ shift) Cholesky -> Rs, Ds
  bool well = false;   // set by the adaptive first factorisation
  auto reduce_and_factor = [&](float shift_rel, bool adaptive) {
#pragma unroll
    for (int q = 0; q < KB; ++q)
      if (bi[q] >= 0 && rg == 1) {
        const int i = 2 * bi[q], j = 2 * bj[q];
        Gp2[i * RP + j] = acc[q][0];
        Gp2[i * RP + j + 1] = acc[q][1];
        Gp2[(i + 1) * RP + j] = acc[q][2];
        Gp2[(i + 1) * RP + j + 1] = acc[q][3];
      }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < KB; ++q)
      if (bi[q] >= 0 && rg == 0) {   // group 0 + group 1, always in this order
        const int i = 2 * bi[q], j = 2 * bj[q];
        Gp[i * RP + j] = acc[q][0] + Gp2[i * RP + j];
        Gp[i * RP + j + 1] = acc[q][1] + Gp2[i * RP + j + 1];
        Gp[(i + 1) * RP + j] = acc[q][2] + Gp2[(i + 1) * RP + j];
        Gp[(i + 1) * RP + j + 1] = acc[q][3] + Gp2[(i + 1) * RP + j + 1];
      }
    cluster_sync_all();   // every CTA's Gp is complete and visible cluster-wide
#pragma unroll
    for (int q = 0; q < KB; ++q)
      if (bi[q] >= 0 && rg == 0) {
        const int i = 2 * bi[q], j = 2 * bj[q];
        float v[NC][4];
#pragma unroll
        for (uint32_t rk = 0; rk < NC; ++rk) {   // all remote loads in flight before the first add
          v[rk][0] = ld_dsmem_f32(Gp + i * RP + j, rk);
          v[rk][1] = ld_dsmem_f32(Gp + i * RP + j + 1, rk);
          v[rk][2] = ld_dsmem_f32(Gp + (i + 1) * RP + j, rk);
          v[rk][3] = ld_dsmem_f32(Gp + (i + 1) * RP + j + 1, rk);
        }
        float sm[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (uint32_t rk = 0; rk < NC; ++rk) {   // fixed order: identical bits in every CTA
          sm[0] += v[rk][0];
          sm[1] += v[rk][1];
          sm[2] += v[rk][2];
          sm[3] += v[rk][3];
        }
        G[i][j] = sm[0];
        G[i][j + 1] = sm[1];
        if (i != j) G[i + 1][j] = sm[2];   // (below the diagonal inside a diagonal block: never read)
        G[i + 1][j + 1] = sm[3];
      }
    cluster_sync_all();   // all remote reads of Gp are done: it may be overwritten by the next pass
    __syncthreads();
    if (t == 0) {
      float tr = 0.f, mx = 0.f;
      for (int i = 0; i < r; ++i) {
        tr += G[i][i];
        mx = fmaxf(mx, G[i][i]);
      }
      scal[0] = tr;
      scal[2] = mx;
    }
    __syncthreads();
    const float tr = scal[0];
    const float floor_piv = fmaxf(tr, 1e-30f) * 1e-7f;
    bool shifted = false;
    if (adaptive) {
      // first try the plain factorisation on a copy: if every pivot stays above 1 % of the largest diagonal entry,
      // X is well conditioned (cond^2 <~ 100 r) and CholeskyQR2 is enough -- the caller then runs one pass less.
      // Every CTA holds the same bits of G, so every CTA takes the same decision.
      for (int e = t; e < RP * (RP + 1); e += kOrthThreads) (&Gc[0][0])[e] = (&G[0][0])[e];
      __syncthreads();
      cholesky_upper_f32<RP>(G, piv, r, t, floor_piv);
      if (t == 0) {
        float mn = 3.4e38f;
        for (int k = 0; k < r; ++k) mn = fminf(mn, 1.0f / (piv[k] * piv[k]));   // the pivots d_k
        scal[1] = (mn >= 1e-2f * scal[2]) ? 1.f : 0.f;
      }
      __syncthreads();
      well = scal[1] != 0.f;
      if (!well) {   // ill conditioned: restore G, shift, factor again
        for (int e = t; e < RP * (RP + 1); e += kOrthThreads) (&G[0][0])[e] = (&Gc[0][0])[e];
        __syncthreads();
        shifted = true;
      }
    } else if (shift_rel > 0.f) {
      shifted = true;
    }
    if (shifted) {
      if (t < r) G[t][t] += shift_rel * tr;
      __syncthreads();
    }
    if (!adaptive || shifted) cholesky_upper_f32<RP>(G, piv, r, t, floor_piv);
    for (int e = t; e < RP * RP; e += kOrthThreads) {
      const int i = e / RP, j = e % RP;
      Rs[e] = (i < r && j < r && j >= i) ? G[i][j] * piv[i] : 0.f;
    }
    for (int j = t; j < RP; j += kOrthThreads) Ds[j] = (j < r) ? piv[j] : 0.f;
    __syncthreads();
  };
  // X[m] <- X[m] R^-1 for one row held in registers (right-looking forward substitution)
  auto solve_row = [&](float (&xr)[RP]) {
#pragma unroll
    for (int i = 0; i < RP; ++i) {
      const float xi = xr[i] * Ds[i];   // columns >= r: Ds = 0 -> exact zeros in the padding
      xr[i] = xi;
      // row i of R in 16-byte pieces (one LDS.128 per 4 FMAs instead of one LDS per FMA: the kernel is
      // instruction-issue bound, ncu profiles/r2_ncu_lowrank_orth_v3.md)
#pragma unroll
      for (int q = i / 4; q < RP / 4; ++q) {
        const float4 rv = *reinterpret_cast<const float4*>(Rs + i * RP + 4 * q);
        if (4 * q + 0 > i) xr[4 * q + 0] = fmaf(-xi, rv.x, xr[4 * q + 0]);
        if (4 * q + 1 > i) xr[4 * q + 1] = fmaf(-xi, rv.y, xr[4 * q + 1]);
        if (4 * q + 2 > i) xr[4 * q + 2] = fmaf(-xi, rv.z, xr[4 * q + 2]);
        if (4 * q + 3 > i) xr[4 * q + 3] = fmaf(-xi, rv.w, xr[4 * q + 3]);
      }
      // keep row i's loads of R inside step i (hoisting all RP^2 / 2 of them spills the row out of registers)
      asm volatile("" ::: "memory");
    }
  };
  auto load_row = [&](float (&xr)[RP], int m) {
    const float4* xrow = reinterpret_cast<const float4*>(p.X + static_cast<size_t>(m) * RP);
#pragma unroll
    for (int q = 0; q < RP / 4; ++q) {
      const float4 v = xrow[q];
      xr[4 * q] = v.x; xr[4 * q + 1] = v.y; xr[4 * q + 2] = v.z; xr[4 * q + 3] = v.w;
    }
  };

  // ---- pass 1 (first half): sum of the partials, Gram ----
  zero_acc();
  for (int mc = m_begin; mc < m_end; mc += kOrthChunk) {
    const int rows = min(kOrthChunk, m_end - mc);
    __syncthreads();
    for (int i = t; i < rows * (RP / 4); i += kOrthThreads) {
      const int rr = i / (RP / 4), c4 = i % (RP / 4);
      const size_t o = static_cast<size_t>(mc + rr) * RP + 4 * c4;
      float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s0 = 0; s0 < p.S; s0 += 8) {   // up to 8 partial copies' loads in flight, added in order
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          v[u] = (s0 + u < p.S)
                     ? *reinterpret_cast<const float4*>(p.part + static_cast<size_t>(s0 + u) * p.part_stride + o)
                     : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 8; ++u) { sum.x += v[u].x; sum.y += v[u].y; sum.z += v[u].z; sum.w += v[u].w; }
      }
      *reinterpret_cast<float4*>(p.X + o) = sum;
      float* d = Xd + rr * kLdX + 4 * c4;
      d[0] = sum.x; d[1] = sum.y; d[2] = sum.z; d[3] = sum.w;
    }
    __syncthreads();
    gram_chunk(rows);
  }
  // well-conditioned X (the common case): plain factorisation, CholeskyQR2; otherwise the shifted factorisation,
  // which always succeeds and leaves cond(X R1^-1) <= ~100, and one more pass (CholeskyQR3)
  reduce_and_factor(1e-4f, true);
  const int more_passes = (well ? 1 : 2) - (p.light ? 1 : 0);

  // ---- passes 1 (second half) and 2: X <- X R^-1, Gram of the new X, plain Cholesky ----
  for (int pass = 0; pass < more_passes; ++pass) {
    zero_acc();
    for (int mc = m_begin; mc < m_end; mc += kOrthChunk) {
      const int rows = min(kOrthChunk, m_end - mc);
      __syncthreads();
      if (t < rows) {
        float xr[RP];
        load_row(xr, mc + t);
        solve_row(xr);
        float4* xrow = reinterpret_cast<float4*>(p.X + static_cast<size_t>(mc + t) * RP);
#pragma unroll
        for (int q = 0; q < RP / 4; ++q) xrow[q] = make_float4(xr[4 * q], xr[4 * q + 1], xr[4 * q + 2], xr[4 * q + 3]);
#pragma unroll
        for (int j = 0; j < RP; ++j) Xd[t * kLdX + j] = xr[j];
      }
      __syncthreads();
      gram_chunk(rows);
    }
    reduce_and_factor(0.f, false);
  }

  // ---- last pass (second half): X <- X R^-1, outputs ----
  for (int m = m_begin + t; m < m_end; m += kOrthThreads) {
    float xr[RP];
    load_row(xr, m);
    solve_row(xr);
    if (p.out2 && p.half_planes) {
      split_h16_store_row<RP>(xr, reinterpret_cast<__half*>(p.out2), static_cast<size_t>(m) * RP, static_cast<size_t>(M) * RP);
    } else if (p.out2) {
#pragma unroll
      for (int j = 0; j < RP; ++j) p.out2[static_cast<size_t>(m) * RP + j] = split_tf32(xr[j]);
    }
#pragma unroll
    for (int j = 0; j < RP; ++j)
      if (j < r) {
        if (p.out16) p.out16[static_cast<size_t>(m) * r + j] = __float2half_rn(xr[j]);
        if (p.out32c) p.out32c[static_cast<size_t>(m) * r + j] = xr[j];
      }
  }
}

// ---------------------------------------------------------------------------------------
// recon = base + fp16(U V), second version.  k_lr_reconstruct_mma loads U (2-byte loads), V, multiplies, and
// only then issues its reads of base: three dependent phases per CTA and 134 registers (ncu, round 2: 30 us
// for 56.6 MB at r = 32, long-scoreboard stalls).  Here a CTA tile is 64 rows x 128 columns: U and the V slab
// arrive by cp.async, the base fragments of the tile are fetched into registers BEFORE the MMAs (so HBM latency
// overlaps the tensor work), and the product tile goes through shared memory for 16-byte row segments.
// grid (ceil(C / 128), ceil(N / 64)), block 128; r % 8 == 0 and 16-byte aligned U for the cp.async path.
// ---------------------------------------------------------------------------------------
template <int KS>
__global__ void __launch_bounds__(128, 3) k_lr_reconstruct_v2(const __half* __restrict__ U, const __half* __restrict__ V,
                                                             const __half* __restrict__ base, __half* __restrict__ recon,
                                                             int N, int C, int r) {
  constexpr int KP = KS * 16, BN = 128;
  constexpr int kLdU = KP + 8, kLdV = BN + 8;
  __shared__ __align__(16) __half Us[64 * kLdU];
  __shared__ __align__(16) __half Vs[(KP > 64 ? KP : 64) * kLdV];   // reused as the fp16 product tile [64][kLdV]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int c0 = blockIdx.x * BN, n0 = blockIdx.y * 64;
  // U tile: rows of r halves (r % 8 == 0: 16-byte chunks), zero-filled beyond r / N
  const int uch = KP / 8;
  for (int i = tid; i < 64 * uch; i += 128) {
    const int rr = i / uch, kc = i % uch;
    const bool ok = n0 + rr < N && 8 * kc < r;
    cp_async16(Us + rr * kLdU + 8 * kc, U + (ok ? static_cast<size_t>(n0 + rr) * r + 8 * kc : 0), ok);
  }
  for (int i = tid; i < KP * (BN / 8); i += 128) {
    const int k = i / (BN / 8), cc = i % (BN / 8);
    const bool ok = k < r && c0 + 8 * cc < C;
    cp_async16(Vs + k * kLdV + 8 * cc, V + (ok ? static_cast<size_t>(k) * C + c0 + 8 * cc : 0), ok);
  }
  cp_async_commit();
  // this warp's 16 rows x 128 columns of base: 256 16-byte segments, 8 per lane, in flight during the MMAs
  uint4 bq[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int i = lane + 32 * q;
    const int n = n0 + 16 * warp + i / (BN / 8), c = c0 + 8 * (i % (BN / 8));
    bq[q] = make_uint4(0, 0, 0, 0);
    if (base != nullptr && n < N && c < C) bq[q] = ldg_stream(base + static_cast<size_t>(n) * C + c);
  }
  cp_async_wait<0>();
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
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int i = lane + 32 * q;
    const int rr = 16 * warp + i / (BN / 8), cc = i % (BN / 8);
    const int n = n0 + rr, c = c0 + 8 * cc;
    if (n >= N || c >= C) continue;
    const H8 pr = as_h8(*reinterpret_cast<const uint4*>(Ps + rr * kLdV + 8 * cc));
    H8 o = pr;
    if (base != nullptr) {
      const H8 b = as_h8(bq[q]);
#pragma unroll
      for (int w = 0; w < 4; ++w) o.w[w] = h22u(__hadd2_rn(u2h2(b.w[w]), u2h2(pr.w[w])));  // base + recv_delta
    }
    stg_stream(recon + static_cast<size_t>(n) * C + c, as_u4(o));
  }
}


// ---------------------------------------------------------------------------------------
// LOW_RANK_Q decode fused into the reconstruct: recon = base + fp16(deq(qU) deq(qV^T)^T) straight from the wire
// payload [qU (N/2, r) u8 | sU (r) | mU (r) | qV^T (C/2, r) u8 | sV (r) | mV (r)] (slowpath.py:69-75, :156-164).
// The per-call path decodes U and V^T with two int4 kernels, transposes V^T, copies it, and then reconstructs:
// ~10 launches per tensor, 117 us each in the 2-GPU FLUX run.  Here the tile loaders of k_lr_reconstruct_v2 read
// the packed nibbles themselves: a 64-row tile of U is 32 x r contiguous bytes, a 128-column slab of V is
// 64 x r contiguous bytes of qV^T.  value = fp16(fp16(code * scale) + min) (compress_quantize.py:636), per column
// of U / of V^T, i.e. per k.      grid (ceil(C / 128), ceil(N / 64)), block 128; N and C even.
// ---------------------------------------------------------------------------------------
template <int KS>
__global__ void __launch_bounds__(128, 3) k_lrq_reconstruct(const uint8_t* __restrict__ qU, const __half* __restrict__ sU,
                                                           const __half* __restrict__ mU, const uint8_t* __restrict__ qVt,
                                                           const __half* __restrict__ sV, const __half* __restrict__ mV,
                                                           const __half* __restrict__ base, __half* __restrict__ recon,
                                                           int N, int C, int r, int vec) {
  constexpr int KP = KS * 16, BN = 128;
  constexpr int kLdU = KP + 8, kLdV = BN + 8;
  __shared__ __align__(16) __half Us[64 * kLdU];
  __shared__ __align__(16) __half Vs[(KP > 64 ? KP : 64) * kLdV];   // reused as the fp16 product tile [64][kLdV]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int c0 = blockIdx.x * BN, n0 = blockIdx.y * 64;
  const __half zero = __float2half_rn(0.f);
  // this warp's 16 rows x 128 columns of base first: in flight while the nibbles are decoded
  uint4 bq[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int i = lane + 32 * q;
    const int n = n0 + 16 * warp + i / (BN / 8), c = c0 + 8 * (i % (BN / 8));
    bq[q] = make_uint4(0, 0, 0, 0);
    if (base != nullptr && n < N && c < C) bq[q] = ldg_stream(base + static_cast<size_t>(n) * C + c);
  }
  // the four scale vectors once per CTA: [sU | mU | sV | mV], zero beyond r
  __shared__ __align__(16) __half sc_s[4][KP];
  for (int i = tid; i < 4 * KP; i += 128) {
    const int which = i / KP, k = i % KP;
    const __half* src = which == 0 ? sU : (which == 1 ? mU : (which == 2 ? sV : mV));
    sc_s[which][k] = k < r ? src[k] : zero;
  }
  __syncthreads();
  // two codes -> fp16(fp16(code * scale) + min), both roundings as in compress_quantize.py:636
  auto deq2 = [](uint32_t c0_, uint32_t c1_, __half2 sc, __half2 mn) {
    const __half2 q = __halves2half2(__ushort2half_rn(static_cast<unsigned short>(c0_)),
                                     __ushort2half_rn(static_cast<unsigned short>(c1_)));
    return __hadd2_rn(__hmul2_rn(q, sc), mn);
  };
  if (vec) {
    // r % 16 == 0 and both code planes 16-byte aligned: one 16-byte load = 16 consecutive k of one row pair
    const int per = r >> 4;
    for (int i = tid; i < 32 * per; i += 128) {
      const int rp = i / per, k0 = (i % per) << 4;
      const int n = n0 + 2 * rp;
      uint4 q = make_uint4(0, 0, 0, 0);
      if (n < N) q = __ldg(reinterpret_cast<const uint4*>(qU + (static_cast<size_t>(n) >> 1) * r + k0));
      const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
      H8 lo[2], hi[2];
#pragma unroll
      for (int w = 0; w < 4; ++w) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int k = k0 + 4 * w + 2 * j;
          const uint32_t b0 = (w4[w] >> (16 * j)) & 0xFFu, b1 = (w4[w] >> (16 * j + 8)) & 0xFFu;
          const __half2 sc = *reinterpret_cast<const __half2*>(&sc_s[0][k]), mn = *reinterpret_cast<const __half2*>(&sc_s[1][k]);
          lo[w >> 1].w[2 * (w & 1) + j] = h22u(deq2(b0 & 0xFu, b1 & 0xFu, sc, mn));
          hi[w >> 1].w[2 * (w & 1) + j] = h22u(deq2(b0 >> 4, b1 >> 4, sc, mn));
        }
      }
      *reinterpret_cast<uint4*>(Us + (2 * rp) * kLdU + k0) = as_u4(lo[0]);
      *reinterpret_cast<uint4*>(Us + (2 * rp) * kLdU + k0 + 8) = as_u4(lo[1]);
      *reinterpret_cast<uint4*>(Us + (2 * rp + 1) * kLdU + k0) = as_u4(hi[0]);
      *reinterpret_cast<uint4*>(Us + (2 * rp + 1) * kLdU + k0 + 8) = as_u4(hi[1]);
    }
    for (int i = tid; i < (BN / 2) * per; i += 128) {
      const int cp = i / per, k0 = (i % per) << 4;
      const int c = c0 + 2 * cp;
      uint4 q = make_uint4(0, 0, 0, 0);
      if (c < C) q = __ldg(reinterpret_cast<const uint4*>(qVt + (static_cast<size_t>(c) >> 1) * r + k0));
      const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int w = 0; w < 4; ++w) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = k0 + 4 * w + j;
          const uint32_t b = (w4[w] >> (8 * j)) & 0xFFu;
          const __half2 sc = __half2half2(sc_s[2][k]), mn = __half2half2(sc_s[3][k]);
          *reinterpret_cast<__half2*>(Vs + k * kLdV + 2 * cp) = deq2(b & 0xFu, b >> 4, sc, mn);
        }
      }
    }
  } else {
    for (int i = tid; i < 32 * KP; i += 128) {   // U: row pair rp, column k
      const int rp = i / KP, k = i % KP;
      const int n = n0 + 2 * rp;
      __half2 val = __halves2half2(zero, zero);
      if (k < r && n < N) {
        const uint32_t b = qU[(static_cast<size_t>(n) >> 1) * r + k];
        val = deq2(b & 0xFu, b >> 4, __half2half2(sc_s[0][k]), __half2half2(sc_s[1][k]));
      }
      Us[(2 * rp) * kLdU + k] = __low2half(val);
      Us[(2 * rp + 1) * kLdU + k] = __high2half(val);
    }
    for (int i = tid; i < (BN / 2) * KP; i += 128) {   // V: column pair cp, row k
      const int cp = i / KP, k = i % KP;
      const int c = c0 + 2 * cp;
      __half2 val = __halves2half2(zero, zero);
      if (k < r && c < C) {
        const uint32_t b = qVt[(static_cast<size_t>(c) >> 1) * r + k];
        val = deq2(b & 0xFu, b >> 4, __half2half2(sc_s[2][k]), __half2half2(sc_s[3][k]));
      }
      *reinterpret_cast<__half2*>(Vs + k * kLdV + 2 * cp) = val;
    }
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
      uint32_t b[4];
      ldmatrix_x4_trans(b, Vs + (16 * kk + l8 + (mi & 1) * 8) * kLdV + 16 * j2 + (mi >> 1) * 8);
      mma_f16(acc[2 * j2], a, b[0], b[1]);
      mma_f16(acc[2 * j2 + 1], a, b[2], b[3]);
    }
  }
  __syncthreads();
  __half* Ps = Vs;
#pragma unroll
  for (int j = 0; j < BN / 8; ++j) {
    const int col = 8 * j + 2 * t;
    *reinterpret_cast<__half2*>(Ps + (16 * warp + g) * kLdV + col) = __floats2half2_rn(acc[j][0], acc[j][1]);
    *reinterpret_cast<__half2*>(Ps + (16 * warp + g + 8) * kLdV + col) = __floats2half2_rn(acc[j][2], acc[j][3]);
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int i = lane + 32 * q;
    const int rr = 16 * warp + i / (BN / 8), cc = i % (BN / 8);
    const int n = n0 + rr, c = c0 + 8 * cc;
    if (n >= N || c >= C) continue;
    const H8 pr = as_h8(*reinterpret_cast<const uint4*>(Ps + rr * kLdV + 8 * cc));
    H8 o = pr;
    if (base != nullptr) {
      const H8 b = as_h8(bq[q]);
#pragma unroll
      for (int w = 0; w < 4; ++w) o.w[w] = h22u(__hadd2_rn(u2h2(b.w[w]), u2h2(pr.w[w])));
    }
    stg_stream(recon + static_cast<size_t>(n) * C + c, as_u4(o));
  }
}

}  // namespace cf
