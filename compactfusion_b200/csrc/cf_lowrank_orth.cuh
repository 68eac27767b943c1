// Cluster CholeskyQR2 of the LOW_RANK projector (included by cf_lowrank.cu after cf_lowrank_mma.cuh, whose
// cholesky_upper_256 / split_tf32 it uses).
#pragma once

#include "cf_lowrank_mma.cuh"

namespace cf {

// ---------------------------------------------------------------------------------------
// CholeskyQR2 of X = sum of S split-K partials (M x RP, fp32) as ONE kernel on ONE thread-block cluster.
//
// Round 1 measured 93 us per orthonormalisation (sum kernel + 2 x (Gram / Cholesky on a 16-CTA grid with a
// global ticket + a forward-substitution kernel)): five dependent launches of latency-bound work, three times
// per projector call -- 2/3 of the call.  Here the kClusterCtas CTAs of a cluster each own a contiguous row
// range of X and make three streaming passes over it (the matrix is < 1 MB: L2-resident):
//   A: X = sum_s part[s]            -> X (global), partial Gram G_cta = X_cta^T X_cta   (upper 2x2 blocks)
//   B: X = X R1^-1                  -> X (global), partial Gram of the new X
//   C: X = X R2^-1                  -> outputs ({hi,lo} TF32 pairs / fp16 / compact fp32)
// Between the passes the partial Grams are reduced over the cluster through distributed shared memory in a
// FIXED rank order (every CTA computes the same bits, no atomics), and every CTA factors the r x r Gram
// redundantly (cholesky_upper_256), so nothing ever goes back to global memory or to the host.
//
// Precision of the Gram matrix.  fp64 FMAs are the wrong tool on this part: the first version of this kernel
// accumulated every product in fp64 and spent ~100 us per call in DFMA (profiles/r2_kernel_times_lowrank_orth_v1.md:
// ~2.4 DFMA per clock per SM).  Now products and sums are fp32 inside blocks of 32 rows and the block sums are
// added in fp64 (one DADD per 32 FFMAs): the error of an entry is ~0.5 u32 |x_i| |x_j| (product rounding
// ~u32 / sqrt(M), summation ~u32 * 32 / sqrt(M)), where plain fp32 accumulation over M = 4608 rows would give
// ~sqrt(M) u32.  The factorisation itself stays fp64.  CholeskyQR2 then holds for cond(X) up to ~1e3; beyond that
// the pivot floor in cholesky_upper_256 keeps the result finite (rank-deficient directions carry no energy).
// ---------------------------------------------------------------------------------------
constexpr int kClusterCtas = 8;    // portable cluster size
// rows staged per step: one row per thread in the substitution passes (RP = 64: half, to fit shared memory)
template <int RP>
constexpr int orth_chunk() { return RP <= 32 ? 256 : 128; }

struct OrthParams {
  const float* part;     // S partial copies of X, `part_stride` floats apart
  int S;
  size_t part_stride;
  float* X;              // (M, RP) scratch: the sum, then the intermediate X R1^-1
  int M, r;
  float2* out2;          // optional (M, RP) {hi, lo} TF32 pairs (the next product's skinny operand)
  __half* out16;         // optional (M, r) fp16
  float* out32c;         // optional (M, r) compact fp32
};

template <int RP>
constexpr size_t lr_orth_smem() {
  return sizeof(float) * orth_chunk<RP>() * (RP + 2)                  // row chunk (fp32)
         + sizeof(double) * RP * RP                                   // this CTA's partial Gram (read by the cluster)
         + sizeof(double) * (kLrMaxRank * (kLrMaxRank + 1) + kLrMaxRank)  // G + pivots for the factorisation
         + sizeof(float) * (RP * RP + RP);                            // R (upper) and 1 / diag
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
// read a double from the same shared-memory offset of CTA `rank` of the cluster (DSMEM)
__device__ __forceinline__ double ld_dsmem_f64(const double* local, uint32_t rank) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(local));
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(a), "r"(rank));
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote) : "memory");
  return v;
}

template <int RP>
__global__ void __launch_bounds__(256, 1) k_lr_orth(const OrthParams p) {
  extern __shared__ __align__(16) unsigned char orth_raw[];
  constexpr int kLdX = RP + 2, kOrthChunk = orth_chunk<RP>();
  float* Xd = reinterpret_cast<float*>(orth_raw);                         // [kOrthChunk][kLdX]
  double* Gp = reinterpret_cast<double*>(Xd + kOrthChunk * kLdX);         // [RP][RP] partial Gram of this CTA
  double (*G)[kLrMaxRank + 1] = reinterpret_cast<double (*)[kLrMaxRank + 1]>(Gp + RP * RP);
  double* piv = reinterpret_cast<double*>(G) + kLrMaxRank * (kLrMaxRank + 1);
  float* Rs = reinterpret_cast<float*>(piv + kLrMaxRank);                 // [RP][RP]
  float* Ds = Rs + RP * RP;                                               // [RP]
  const int t = threadIdx.x, r = p.r, M = p.M;
  const uint32_t rank = cluster_ctarank();
  int rows_per = (M + kClusterCtas - 1) / kClusterCtas;
  rows_per = (rows_per + 3) / 4 * 4;
  const int m_begin = min(M, static_cast<int>(rank) * rows_per);
  const int m_end = min(M, m_begin + rows_per);

  // upper-triangular 2x2 blocks of the Gram matrix, dealt to the threads round-robin
  constexpr int NB = RP / 2, kBlocks = NB * (NB + 1) / 2, KB = (kBlocks + 255) / 256;
  int bi[KB], bj[KB];
#pragma unroll
  for (int q = 0; q < KB; ++q) {
    int id = t + 256 * q, i = 0;
    if (id >= kBlocks) id = -1;
    if (id >= 0)
      while (id >= NB - i) { id -= NB - i; ++i; }   // row i of the block triangle holds NB - i blocks
    bi[q] = id >= 0 ? i : -1;
    bj[q] = id >= 0 ? i + id : -1;
  }
  double acc[KB][4];

  auto zero_acc = [&]() {
#pragma unroll
    for (int q = 0; q < KB; ++q)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[q][e] = 0.0;
  };
  auto gram_chunk = [&](int rows) {   // acc += Xd[0..rows)^T Xd[0..rows) on this thread's blocks
    for (int rb = 0; rb < rows; rb += 32) {   // fp32 inside a block of 32 rows, fp64 across blocks
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
          acc[q][0] += static_cast<double>(a0);
          acc[q][1] += static_cast<double>(a1);
          acc[q][2] += static_cast<double>(a2);
          acc[q][3] += static_cast<double>(a3);
        }
      }
    }
  };
  // partial Gram -> Gp; cluster-wide sum in rank order -> G (upper); Cholesky -> Rs, Ds
  auto reduce_and_factor = [&]() {
#pragma unroll
    for (int q = 0; q < KB; ++q)
      if (bi[q] >= 0) {
        const int i = 2 * bi[q], j = 2 * bj[q];
        Gp[i * RP + j] = acc[q][0];
        Gp[i * RP + j + 1] = acc[q][1];
        Gp[(i + 1) * RP + j] = acc[q][2];
        Gp[(i + 1) * RP + j + 1] = acc[q][3];
      }
    cluster_sync_all();   // every CTA's Gp is complete and visible cluster-wide
#pragma unroll
    for (int q = 0; q < KB; ++q)
      if (bi[q] >= 0) {
        const int i = 2 * bi[q], j = 2 * bj[q];
        double v[kClusterCtas][4];
#pragma unroll
        for (uint32_t rk = 0; rk < kClusterCtas; ++rk) {   // all 32 remote loads in flight before the first add
          v[rk][0] = ld_dsmem_f64(Gp + i * RP + j, rk);
          v[rk][1] = ld_dsmem_f64(Gp + i * RP + j + 1, rk);
          v[rk][2] = ld_dsmem_f64(Gp + (i + 1) * RP + j, rk);
          v[rk][3] = ld_dsmem_f64(Gp + (i + 1) * RP + j + 1, rk);
        }
        double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (uint32_t rk = 0; rk < kClusterCtas; ++rk) {   // fixed order: identical bits in every CTA
          s[0] += v[rk][0];
          s[1] += v[rk][1];
          s[2] += v[rk][2];
          s[3] += v[rk][3];
        }
        G[i][j] = s[0];
        G[i][j + 1] = s[1];
        if (i != j) G[i + 1][j] = s[2];   // (below the diagonal inside a diagonal block: never read)
        G[i + 1][j + 1] = s[3];
      }
    cluster_sync_all();   // all remote reads of Gp are done: it may be overwritten by the next round
    __syncthreads();
    cholesky_upper_256(G, piv, r, t);
    for (int e = t; e < RP * RP; e += 256) {
      const int i = e / RP, j = e % RP;
      Rs[e] = (i < r && j < r && j >= i) ? static_cast<float>(G[i][j] * piv[i]) : 0.f;
    }
    for (int j = t; j < RP; j += 256) Ds[j] = (j < r) ? static_cast<float>(piv[j]) : 0.f;
    __syncthreads();
  };
  // X[m] <- X[m] R^-1 for one row held in registers (right-looking forward substitution)
  auto solve_row = [&](float (&xr)[RP]) {
#pragma unroll
    for (int i = 0; i < RP; ++i) {
      const float xi = xr[i] * Ds[i];   // columns >= r: Ds = 0 -> exact zeros in the padding
      xr[i] = xi;
#pragma unroll
      for (int j = i + 1; j < RP; ++j) xr[j] = fmaf(-xi, Rs[i * RP + j], xr[j]);
      // keep row i's loads of R inside step i (hoisting all RP^2 / 2 of them spills the row out of registers)
      asm volatile("" ::: "memory");
    }
  };

  // ---- pass A: sum of the partials, first Gram ----
  zero_acc();
  for (int mc = m_begin; mc < m_end; mc += kOrthChunk) {
    const int rows = min(kOrthChunk, m_end - mc);
    __syncthreads();
    for (int i = t; i < rows * (RP / 4); i += 256) {
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
  reduce_and_factor();

  // ---- pass B: X <- X R1^-1, second Gram ----
  zero_acc();
  for (int mc = m_begin; mc < m_end; mc += kOrthChunk) {
    const int rows = min(kOrthChunk, m_end - mc);
    __syncthreads();
    if (t < rows) {
      float xr[RP];
      float4* xrow = reinterpret_cast<float4*>(p.X + static_cast<size_t>(mc + t) * RP);
#pragma unroll
      for (int q = 0; q < RP / 4; ++q) {
        const float4 v = xrow[q];
        xr[4 * q] = v.x; xr[4 * q + 1] = v.y; xr[4 * q + 2] = v.z; xr[4 * q + 3] = v.w;
      }
      solve_row(xr);
#pragma unroll
      for (int q = 0; q < RP / 4; ++q) xrow[q] = make_float4(xr[4 * q], xr[4 * q + 1], xr[4 * q + 2], xr[4 * q + 3]);
#pragma unroll
      for (int j = 0; j < RP; ++j) Xd[t * kLdX + j] = xr[j];
    }
    __syncthreads();
    gram_chunk(rows);
  }
  reduce_and_factor();

  // ---- pass C: X <- X R2^-1, outputs ----
  for (int m = m_begin + t; m < m_end; m += 256) {
    float xr[RP];
    const float4* xrow = reinterpret_cast<const float4*>(p.X + static_cast<size_t>(m) * RP);
#pragma unroll
    for (int q = 0; q < RP / 4; ++q) {
      const float4 v = xrow[q];
      xr[4 * q] = v.x; xr[4 * q + 1] = v.y; xr[4 * q + 2] = v.z; xr[4 * q + 3] = v.w;
    }
    solve_row(xr);
    if (p.out2) {
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

}  // namespace cf
