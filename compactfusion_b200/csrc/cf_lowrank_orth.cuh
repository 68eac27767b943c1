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
//   A: X = sum_s part[s]            -> X (global), partial Gram G_cta = X_cta^T X_cta   (fp64, upper 2x2 blocks)
//   B: X = X R1^-1                  -> X (global), partial Gram of the new X
//   C: X = X R2^-1                  -> outputs ({hi,lo} TF32 pairs / fp16 / compact fp32)
// Between the passes the partial Grams are reduced over the cluster through distributed shared memory in a
// FIXED rank order (every CTA computes the same bits, no atomics), and every CTA factors the r x r Gram
// redundantly (cholesky_upper_256), so nothing ever goes back to global memory or to the host.
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
  return sizeof(double) * orth_chunk<RP>() * (RP + 2)                 // row chunk as fp64
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
  double* Xd = reinterpret_cast<double*>(orth_raw);                       // [kOrthChunk][kLdX]
  double* Gp = Xd + kOrthChunk * kLdX;                                    // [RP][RP] partial Gram of this CTA
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
#pragma unroll 4
    for (int rr = 0; rr < rows; ++rr) {
#pragma unroll
      for (int q = 0; q < KB; ++q) {
        if (bi[q] >= 0) {
          const double2 xi = *reinterpret_cast<const double2*>(Xd + rr * kLdX + 2 * bi[q]);
          const double2 xj = *reinterpret_cast<const double2*>(Xd + rr * kLdX + 2 * bj[q]);
          acc[q][0] = fma(xi.x, xj.x, acc[q][0]);
          acc[q][1] = fma(xi.x, xj.y, acc[q][1]);
          acc[q][2] = fma(xi.y, xj.x, acc[q][2]);
          acc[q][3] = fma(xi.y, xj.y, acc[q][3]);
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
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        for (uint32_t rk = 0; rk < kClusterCtas; ++rk) {   // fixed order: identical bits in every CTA
          s[0] += ld_dsmem_f64(Gp + i * RP + j, rk);
          s[1] += ld_dsmem_f64(Gp + i * RP + j + 1, rk);
          s[2] += ld_dsmem_f64(Gp + (i + 1) * RP + j, rk);
          s[3] += ld_dsmem_f64(Gp + (i + 1) * RP + j + 1, rk);
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
      for (int sidx = 0; sidx < p.S; ++sidx) {
        const float4 v = *reinterpret_cast<const float4*>(p.part + static_cast<size_t>(sidx) * p.part_stride + o);
        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
      }
      *reinterpret_cast<float4*>(p.X + o) = sum;
      double* d = Xd + rr * kLdX + 4 * c4;
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

}  // namespace cf
