// LOW_RANK projector (randomised subspace iteration) and fused reconstruct for sm_100a.
//
// Reference: xfuser/compact/compress_lowrank.py:16-62 (subspace_iter: A.float(); Q = q0;
// iters x { Z = A^T (A Q); Q = qr(Z) }; U = qr(A Q); V = U^T A) and slowpath.py:152-154
// (decompress = torch.matmul(u, v)).  The reference runs 6 cuBLAS skinny GEMMs over an fp32
// copy of A plus 4 cuSOLVER QR calls; here A = x - base is formed on the fly from the fp16
// operands (never materialised), the two skinny products are tiled SIMT fp32 kernels (r <= 64:
// 2r flop per 2 bytes of A, below the B200 ridge), and QR is replaced by CholeskyQR2 with an
// fp64 Gram matrix, which spans the same subspace -- U V, the only quantity the wire format
// carries, is invariant to that choice.
//
// All intermediates are fp32 row-major with leading dimension RP = round_up(r, 8); padding
// columns are kept at zero.
#include <stdlib.h>

#include "cf_common.cuh"
#include "cf_lowrank_mma.cuh"
#include "cf_lowrank_orth.cuh"

namespace cf {

constexpr int kMaxRank = 64;
constexpr int kGramParts = 64;

__device__ __forceinline__ float delta_at(const __half* __restrict__ x, const __half* __restrict__ base, size_t i) {
  // one fp16 rounding, like `x - base` in the reference (main.py:229), then exact widening
  return base ? __half2float(__hsub_rn(x[i], base[i])) : __half2float(x[i]);
}

// ---------------------------------------------------------------------------------------
// Y (N, RP) = A (N, C) * Q (C, RP)             grid ceil(N/32), block 128
// thread (tr = t>>3, tc = t&7) owns rows {2tr, 2tr+1} x cols {tc + 8j}
// ---------------------------------------------------------------------------------------
template <int RP>
__global__ void __launch_bounds__(128) k_lr_AQ(const __half* __restrict__ x, const __half* __restrict__ base,
                                              const float* __restrict__ Q, float* __restrict__ Y, int N, int C) {
  constexpr int BM = 32, BK = 32, NJ = RP / 8;
  __shared__ float As[BM][BK + 1];
  __shared__ float Qs[BK][RP];
  const int t = threadIdx.x, tr = t >> 3, tc = t & 7;
  const int row0 = blockIdx.x * BM;
  float acc0[NJ], acc1[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) acc0[j] = acc1[j] = 0.f;
  for (int c0 = 0; c0 < C; c0 += BK) {
    // A tile: 32 rows x 32 cols, coalesced along c
    for (int i = t; i < BM * BK; i += 128) {
      const int r = i / BK, k = i % BK;
      const int n = row0 + r, c = c0 + k;
      As[r][k] = (n < N && c < C) ? delta_at(x, base, static_cast<size_t>(n) * C + c) : 0.f;
    }
    for (int i = t; i < BK * RP; i += 128) {
      const int k = i / RP, j = i % RP;
      Qs[k][j] = (c0 + k < C) ? Q[static_cast<size_t>(c0 + k) * RP + j] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < BK; ++k) {
      const float a0 = As[2 * tr][k], a1 = As[2 * tr + 1][k];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const float q = Qs[k][tc + 8 * j];
        acc0[j] = fmaf(a0, q, acc0[j]);
        acc1[j] = fmaf(a1, q, acc1[j]);
      }
    }
    __syncthreads();
  }
  const int n0 = row0 + 2 * tr;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    if (n0 < N) Y[static_cast<size_t>(n0) * RP + tc + 8 * j] = acc0[j];
    if (n0 + 1 < N) Y[static_cast<size_t>(n0 + 1) * RP + tc + 8 * j] = acc1[j];
  }
}

// ---------------------------------------------------------------------------------------
// Zpart[s] (C, RP) = A[rows of split s]^T (C, n_s) * Y (n_s, RP)     grid (ceil(C/32), S), block 128
// thread (tr, tc) owns cols-of-A {2tr, 2tr+1} x cols-of-Y {tc + 8j}
// ---------------------------------------------------------------------------------------
template <int RP>
__global__ void __launch_bounds__(128) k_lr_AtY(const __half* __restrict__ x, const __half* __restrict__ base,
                                               const float* __restrict__ Y, float* __restrict__ Zpart, int N, int C,
                                               int rows_per_split) {
  constexpr int BC = 32, BK = 32, NJ = RP / 8;
  __shared__ float As[BK][BC + 1];  // [n][c]
  __shared__ float Ys[BK][RP];
  const int t = threadIdx.x, tr = t >> 3, tc = t & 7;
  const int col0 = blockIdx.x * BC;
  const int n_begin = blockIdx.y * rows_per_split;
  const int n_end = min(N, n_begin + rows_per_split);
  float acc0[NJ], acc1[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) acc0[j] = acc1[j] = 0.f;
  for (int n0 = n_begin; n0 < n_end; n0 += BK) {
    for (int i = t; i < BK * BC; i += 128) {
      const int k = i / BC, cc = i % BC;
      const int n = n0 + k, c = col0 + cc;
      As[k][cc] = (n < n_end && c < C) ? delta_at(x, base, static_cast<size_t>(n) * C + c) : 0.f;
    }
    for (int i = t; i < BK * RP; i += 128) {
      const int k = i / RP, j = i % RP;
      Ys[k][j] = (n0 + k < n_end) ? Y[static_cast<size_t>(n0 + k) * RP + j] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < BK; ++k) {
      const float a0 = As[k][2 * tr], a1 = As[k][2 * tr + 1];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const float y = Ys[k][tc + 8 * j];
        acc0[j] = fmaf(a0, y, acc0[j]);
        acc1[j] = fmaf(a1, y, acc1[j]);
      }
    }
    __syncthreads();
  }
  float* Z = Zpart + static_cast<size_t>(blockIdx.y) * C * RP;
  const int c = col0 + 2 * tr;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    if (c < C) Z[static_cast<size_t>(c) * RP + tc + 8 * j] = acc0[j];
    if (c + 1 < C) Z[static_cast<size_t>(c + 1) * RP + tc + 8 * j] = acc1[j];
  }
}

__global__ void __launch_bounds__(256) k_lr_sum_parts(const float* __restrict__ part, float* __restrict__ out,
                                                     size_t count, int S) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < S; ++p) s += part[static_cast<size_t>(p) * count + i];
    out[i] = s;
  }
}

// ---------------------------------------------------------------------------------------
// CholeskyQR: G = X^T X (fp64, r x r) -> R = chol(G) upper -> X <- X R^{-1}
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_lr_gram(const float* __restrict__ X, double* __restrict__ Gpart, int M,
                                                int RP, int r, int rows_per_part) {
  __shared__ float Xs[32][kMaxRank + 1];
  const int m_begin = blockIdx.x * rows_per_part;
  const int m_end = min(M, m_begin + rows_per_part);
  const int t = threadIdx.x;
  // thread owns entries e = t, t+256, ... of the r x r matrix (<= 16 each)
  double acc[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) acc[q] = 0.0;
  for (int m0 = m_begin; m0 < m_end; m0 += 32) {
    for (int i = t; i < 32 * r; i += 256) {
      const int rr = i / r, j = i % r;
      Xs[rr][j] = (m0 + rr < m_end) ? X[static_cast<size_t>(m0 + rr) * RP + j] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int e = t + 256 * q;
      if (e < r * r) {
        const int i = e / r, j = e % r;
        double s = 0.0;
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) s += static_cast<double>(Xs[rr][i]) * static_cast<double>(Xs[rr][j]);
        acc[q] += s;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int e = t + 256 * q;
    if (e < r * r) Gpart[static_cast<size_t>(blockIdx.x) * r * r + e] = acc[q];
  }
}

// one CTA: sum partial Grams, Cholesky (upper R, G = R^T R), invert R; writes Rinv (r x r fp32, row-major)
__global__ void __launch_bounds__(256) k_lr_chol_inv(const double* __restrict__ Gpart, int parts, int r,
                                                    float* __restrict__ Rinv) {
  __shared__ double G[kMaxRank][kMaxRank + 1];  // upper: R;  strictly lower: Rinv^T (written last)
  __shared__ double dinv[kMaxRank];             // diagonal of Rinv
  const int t = threadIdx.x;
  for (int e = t; e < r * r; e += 256) {
    double s = 0.0;
    for (int p = 0; p < parts; ++p) s += Gpart[static_cast<size_t>(p) * r * r + e];
    G[e / r][e % r] = s;
  }
  __syncthreads();
  // right-looking Cholesky on the upper triangle: after step k, row k of G holds R[k][k..]
  double maxdiag = 0.0;
  for (int i = 0; i < r; ++i) maxdiag = fmax(maxdiag, G[i][i]);
  const double floor_piv = fmax(maxdiag, 1e-300) * 1e-14;
  __syncthreads();
  for (int k = 0; k < r; ++k) {
    if (t == 0) {
      double d = G[k][k];
      if (!(d > floor_piv)) d = floor_piv;  // rank-deficient input: keep things finite
      G[k][k] = sqrt(d);
    }
    __syncthreads();
    const double dk = G[k][k];
    for (int j = k + 1 + t; j < r; j += 256) G[k][j] /= dk;
    __syncthreads();
    // trailing update: G[i][j] -= R[k][i] R[k][j] for k < i <= j
    const int rem = r - k - 1;
    for (int e = t; e < rem * rem; e += 256) {
      const int i = k + 1 + e / rem, j = k + 1 + e % rem;
      if (j >= i) G[i][j] -= G[k][i] * G[k][j];
    }
    __syncthreads();
  }
  // invert upper-triangular R by back substitution, one column j per thread; Rinv[i][j] (i < j)
  // is stored at G[j][i] (row j of the lower triangle belongs to thread j alone)
  for (int j = t; j < r; j += 256) {
    dinv[j] = 1.0 / G[j][j];
    for (int i = j - 1; i >= 0; --i) {
      double s = -G[i][j] * dinv[j];
      for (int k = i + 1; k < j; ++k) s -= G[i][k] * G[j][k];
      G[j][i] = s / G[i][i];
    }
  }
  __syncthreads();
  for (int e = t; e < r * r; e += 256) {
    const int i = e / r, j = e % r;
    Rinv[e] = static_cast<float>(i == j ? dinv[i] : (i < j ? G[j][i] : 0.0));
  }
}

// X (M, RP) <- X * Rinv (r x r upper), optional fp16 copy of the result (leading dim r)
__global__ void __launch_bounds__(128) k_lr_apply_rinv(float* __restrict__ X, const float* __restrict__ Rinv, int M,
                                                      int RP, int r, __half* __restrict__ out16) {
  __shared__ float Rs[kMaxRank * kMaxRank];
  for (int e = threadIdx.x; e < r * r; e += blockDim.x) Rs[e] = Rinv[e];
  __syncthreads();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float xr[kMaxRank];
  for (int i = 0; i < r; ++i) xr[i] = X[static_cast<size_t>(m) * RP + i];
  for (int j = 0; j < r; ++j) {
    float s = 0.f;
    for (int i = 0; i <= j; ++i) s = fmaf(xr[i], Rs[i * r + j], s);
    X[static_cast<size_t>(m) * RP + j] = s;
    if (out16) out16[static_cast<size_t>(m) * r + j] = __float2half_rn(s);
  }
}

// V (r, C) fp16 = Vt (C, RP)^T
__global__ void __launch_bounds__(256) k_lr_store_v(const float* __restrict__ Vt, __half* __restrict__ V, int C, int RP,
                                                   int r) {
  const size_t total = static_cast<size_t>(r) * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i / C), c = static_cast<int>(i % C);
    V[i] = __float2half_rn(Vt[static_cast<size_t>(c) * RP + k]);
  }
}

__global__ void __launch_bounds__(256) k_lr_pad_copy(const float* __restrict__ src, float* __restrict__ dst, int rows,
                                                    int r, int RP, bool src_padded) {
  const size_t total = static_cast<size_t>(rows) * RP;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / RP), j = static_cast<int>(i % RP);
    if (src_padded) {  // (rows, RP) -> (rows, r) compact
      if (j < r) dst[static_cast<size_t>(row) * r + j] = src[i];
    } else {           // (rows, r) compact -> (rows, RP) zero padded
      dst[i] = (j < r) ? src[static_cast<size_t>(row) * r + j] : 0.f;
    }
  }
}

// ---------------------------------------------------------------------------------------
// recon = base + fp16(U V)        grid (ceil(C/256), ceil(N/32)), block 256
// thread (tr = t>>5, tc = t&31): rows {tr + 8i, i<4} x 8 consecutive columns
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_lr_reconstruct(const __half* __restrict__ U, const __half* __restrict__ V,
                                                       const __half* __restrict__ base, __half* __restrict__ recon,
                                                       int N, int C, int r) {
  extern __shared__ __half lr_smem[];
  __half* Vs = lr_smem;             // [r][256]
  __half* Us = lr_smem + r * 256;   // [32][r]
  const int t = threadIdx.x, tr = t >> 5, tc = t & 31;
  const int c0 = blockIdx.x * 256, n0 = blockIdx.y * 32;
  for (int i = t; i < r * 32; i += 256) {  // 32 groups of 8 columns per k
    const int k = i >> 5, g = i & 31;
    const int c = c0 + 8 * g;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (c < C) v = *reinterpret_cast<const uint4*>(V + static_cast<size_t>(k) * C + c);
    *reinterpret_cast<uint4*>(Vs + k * 256 + 8 * g) = v;
  }
  for (int i = t; i < 32 * r; i += 256) {
    const int rr = i / r, k = i % r;
    Us[i] = (n0 + rr < N) ? U[static_cast<size_t>(n0 + rr) * r + k] : __float2half_rn(0.f);
  }
  __syncthreads();
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[i][e] = 0.f;
  for (int k = 0; k < r; ++k) {
    const uint4 vv = *reinterpret_cast<const uint4*>(Vs + k * 256 + 8 * tc);
    const __half2* vh = reinterpret_cast<const __half2*>(&vv);
    float vf[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __half22float2(vh[e]);
      vf[2 * e] = f.x;
      vf[2 * e + 1] = f.y;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float u = __half2float(Us[(tr + 8 * i) * r + k]);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[i][e] = fmaf(u, vf[e], acc[i][e]);
    }
  }
  const int c = c0 + 8 * tc;
  if (c >= C) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + tr + 8 * i;
    if (n >= N) continue;
    const size_t off = static_cast<size_t>(n) * C + c;
    uint4 bv = make_uint4(0, 0, 0, 0);
    if (base != nullptr) bv = ldg_stream(base + off);
    const __half* bh = reinterpret_cast<const __half*>(&bv);
    __align__(16) __half o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const __half p = __float2half_rn(acc[i][e]);          // matmul output rounded to fp16 (slowpath.py:154)
      o[e] = (base != nullptr) ? __hadd_rn(bh[e], p) : p;   // base + recv_delta (main.py:376)
    }
    stg_stream(recon + off, *reinterpret_cast<uint4*>(o));
  }
}

// ---------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------
static int rp_of(int r) { return r <= 8 ? 8 : (r <= 16 ? 16 : (r <= 32 ? 32 : 64)); }

struct LrPlan {
  int RP, S, rows_per_split, gram_parts;
  size_t q_off, y_off, zpart_off, gpart_off, rinv_off, total;
};
static LrPlan make_lr_plan(int64_t N, int64_t C, int r) {
  LrPlan p;
  p.RP = rp_of(r);
  const int col_tiles = static_cast<int>((C + 31) / 32);
  int S = (2 * sm_count() + col_tiles - 1) / col_tiles;
  if (S < 1) S = 1;
  if (S > 16) S = 16;
  p.rows_per_split = static_cast<int>(((N + S - 1) / S + 31) / 32 * 32);
  p.S = static_cast<int>((N + p.rows_per_split - 1) / p.rows_per_split);
  p.gram_parts = kGramParts;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += round_up(bytes, 256); return o; };
  p.q_off = take(static_cast<size_t>(C) * p.RP * 4);
  p.y_off = take(static_cast<size_t>(N) * p.RP * 4);
  p.zpart_off = take(static_cast<size_t>(p.S) * C * p.RP * 4);
  p.gpart_off = take(static_cast<size_t>(p.gram_parts) * r * r * 8);
  p.rinv_off = take(static_cast<size_t>(r) * r * 4);
  p.total = off;
  return p;
}
struct LrMmaPlan;
static size_t lr_mma_total(int64_t N, int64_t C, int r);
size_t lowrank_workspace_bytes(int64_t N, int64_t C, int rank) {
  if (rank < 1 || rank > kMaxRank) return 256;
  const size_t a = make_lr_plan(N, C, rank).total, b = (C % 8 == 0) ? lr_mma_total(N, C, rank) : 0;
  return a > b ? a : b;
}

template <int RP>
static void launch_AQ(const __half* x, const __half* b, const float* Q, float* Y, int N, int C, cudaStream_t st) {
  k_lr_AQ<RP><<<(N + 31) / 32, 128, 0, st>>>(x, b, Q, Y, N, C);
}
template <int RP>
static void launch_AtY(const __half* x, const __half* b, const float* Y, float* Zp, int N, int C, const LrPlan& p,
                       cudaStream_t st) {
  dim3 grid((C + 31) / 32, p.S);
  k_lr_AtY<RP><<<grid, 128, 0, st>>>(x, b, Y, Zp, N, C, p.rows_per_split);
}

static int orthonormalise(float* X, int M, int RP, int r, double* gpart, float* rinv, __half* out16, cudaStream_t st) {
  for (int pass = 0; pass < 2; ++pass) {
    int parts = kGramParts;
    int rows_per_part = ((M + parts - 1) / parts + 31) / 32 * 32;
    parts = (M + rows_per_part - 1) / rows_per_part;
    k_lr_gram<<<parts, 256, 0, st>>>(X, gpart, M, RP, r, rows_per_part);
    CF_CHECK_LAUNCH();
    k_lr_chol_inv<<<1, 256, 0, st>>>(gpart, parts, r, rinv);
    CF_CHECK_LAUNCH();
    k_lr_apply_rinv<<<(M + 127) / 128, 128, 0, st>>>(X, rinv, M, RP, r, pass == 1 ? out16 : nullptr);
    CF_CHECK_LAUNCH();
  }
  return CF_OK;
}


// ---------------------------------------------------------------------------------------
// tensor-core path (cf_lowrank_mma.cuh): C % 8 == 0 and 16-byte aligned operands
// ---------------------------------------------------------------------------------------
struct LrMmaPlan {
  int RP;
  int aq_splits, aq_kper;    // Y = A Q:   M = N, K = C
  int aty_splits, aty_kper;  // Z = A^T Y: M = C, K = N
  int gram_ctas_c, gram_rows_c, gram_ctas_n, gram_rows_n;
  size_t q2_off, y2_off, xsum_off, part_off, gpart_off, rinv_off, ticket_off, total;
};
// 32-wide K chunks, 2 stages (58 KB of shared memory: two CTAs per SM, twice the split-K factor): 22.3 / 24.6 us per
// pass against 24.8 / 28.1 us for the 64-wide, 3-stage, one-CTA-per-SM tiles (CF_LR_GEMM=big) at 4608 x 3072, r = 32
static bool lr_gemm_small() {
  static const bool v = [] { const char* e = getenv("CF_LR_GEMM"); return !(e && e[0] == 'b'); }();
  return v;
}
static void lr_split_cfg(int64_t M, int64_t K, int* splits, int* kper) {
  const int64_t m_tiles = (M + kLrBM - 1) / kLrBM;
  // CF_LR_CTAS_PER_SM (A/B): CTAs per SM the split-K factor aims at (default 2 with the small tiles, else 1)
  static const int per_sm_env = [] { const char* e = getenv("CF_LR_CTAS_PER_SM"); return (e && e[0] >= '1' && e[0] <= '4') ? e[0] - '0' : 0; }();
  const int64_t ctas = static_cast<int64_t>(sm_count()) * (per_sm_env ? per_sm_env : (lr_gemm_small() ? 2 : 1));
  int64_t s = (ctas + m_tiles / 2) / m_tiles;  // ~one CTA per SM (two with the small tiles), one wave
  if (s < 1) s = 1;
  if (s > 16) s = 16;
  int64_t kp = ((K + s - 1) / s + kLrBK - 1) / kLrBK * kLrBK;   // (a multiple of 64: fine for both chunk widths)
  *kper = static_cast<int>(kp);
  *splits = static_cast<int>((K + kp - 1) / kp);
}
static void lr_gram_cfg(int64_t M, int r, int* ctas, int* rows) {
  (void)r;
  int64_t parts = 16;  // few partials: the CTA that draws the last ticket adds them up alone
  int64_t rp = ((M + parts - 1) / parts + 31) / 32 * 32;
  *rows = static_cast<int>(rp);
  *ctas = static_cast<int>((M + rp - 1) / rp);
}
static LrMmaPlan make_lr_mma_plan(int64_t N, int64_t C, int r) {
  LrMmaPlan p;
  p.RP = rp_of(r);
  lr_split_cfg(N, C, &p.aq_splits, &p.aq_kper);
  lr_split_cfg(C, N, &p.aty_splits, &p.aty_kper);
  lr_gram_cfg(C, r, &p.gram_ctas_c, &p.gram_rows_c);
  lr_gram_cfg(N, r, &p.gram_ctas_n, &p.gram_rows_n);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += round_up(bytes, 256); return o; };
  const size_t maxm = static_cast<size_t>(N > C ? N : C);
  p.q2_off = take(static_cast<size_t>(C) * p.RP * 8);
  p.y2_off = take(static_cast<size_t>(N) * p.RP * 8);
  p.xsum_off = take(maxm * p.RP * 4);
  const size_t part_aq = static_cast<size_t>(p.aq_splits) * N * p.RP * 4;
  const size_t part_aty = static_cast<size_t>(p.aty_splits) * C * p.RP * 4;
  p.part_off = take(part_aq > part_aty ? part_aq : part_aty);
  p.gpart_off = take(static_cast<size_t>(16) * r * r * 8);
  p.rinv_off = take(static_cast<size_t>(p.RP) * p.RP * 4 + static_cast<size_t>(p.RP) * 4);
  p.ticket_off = take(4096);   // [0]: the legacy Gram kernels' ticket; [1 + it]: max |partial| of iteration it's A Q
  p.total = off;
  return p;
}

// one instantiation of the product kernel: opt-in shared memory size set once, then the launch
template <int RP, bool TRANS, int BK, int ST, bool HB>
static cudaError_t launch_lr_gemm(dim3 grid, cudaStream_t st, const __half* xh, const __half* bh, const float2* B2, float* part,
                                  int n, int c, int kper, unsigned* absmax) {
  constexpr size_t smem = lr_gemm_smem<RP, TRANS, BK, ST, HB>();
  static bool ready = false;   // (per instantiation; one device per process)
  if (!ready) {
    cudaError_t e = cudaFuncSetAttribute(k_lr_gemm<RP, TRANS, BK, ST, HB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    ready = true;
  }
  k_lr_gemm<RP, TRANS, BK, ST, HB><<<grid, kLrThreads, smem, st>>>(xh, bh, B2, part, n, c, kper, absmax);
  return cudaGetLastError();
}
// CF_LR_HB_TILE (A/B of the fp16-plane products): 0 = 32-wide K chunks x 2 stages, 1 = 32 x 3, 2 = 64 x 2 (default),
// 3 = 64 x 3.  Measured at 4608 x 3072, r = 32 (A Q / A^T Y, us): 16.3 / 14.6, 15.0 / 14.6, 14.8 / 14.2, 16.4 / 16.4 --
// a plateau at ~0.6 of the HBM roofline whatever the tile.
static int lr_hb_tile() {
  static const int v = [] { const char* e = getenv("CF_LR_HB_TILE"); return (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 2; }();
  return v;
}

template <int RP>
static int lr_mma_project(const __half* xh, const __half* bh, const float* q0, __half* U, __half* V, float* q_out,
                          int n, int c, int r, int iters, char* ws, const LrMmaPlan& p, cudaStream_t st) {
  float2* Q2 = reinterpret_cast<float2*>(ws + p.q2_off);
  float2* Y2 = reinterpret_cast<float2*>(ws + p.y2_off);
  float* Xsum = reinterpret_cast<float*>(ws + p.xsum_off);
  float* part = reinterpret_cast<float*>(ws + p.part_off);
  double* gpart = reinterpret_cast<double*>(ws + p.gpart_off);
  float* rfac = reinterpret_cast<float*>(ws + p.rinv_off);
  float* rdinv = rfac + RP * RP;
  unsigned* ticket = reinterpret_cast<unsigned*>(ws + p.ticket_off);
  const int small_grid = 2 * sm_count();
  const size_t smem_n = lr_gemm_smem<RP, false>(), smem_t = lr_gemm_smem<RP, true>();
  CF_CHECK_CUDA(cudaFuncSetAttribute(k_lr_gemm<RP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_n)));
  CF_CHECK_CUDA(cudaFuncSetAttribute(k_lr_gemm<RP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_t)));
  CF_CHECK_CUDA(cudaMemsetAsync(ticket, 0, 4096, st));
  // fp16 planes (hi | lo * 2^11) instead of TF32 pairs for the skinny operands of modest size -- the bases Q (the
  // N(0, 1) start, then orthonormal) and the orthonormal U of the last product: two fp16 MMAs per term instead of
  // four TF32 ones and no conversion of the streamed delta fragments.  The raw Y of the intermediate A^T Y products
  // has no such bound and keeps the TF32 pairs.  CF_LR_HALF=0: TF32 everywhere (A/B).
  static const bool half_env = [] { const char* e = getenv("CF_LR_HALF"); return !(e && e[0] == '0'); }();
  const bool hb = half_env && RP >= 16 && lr_gemm_small();
  k_lr_pad_split<<<small_grid, 256, 0, st>>>(q0, Q2, c, r, RP, hb ? 1 : 0);
  CF_CHECK_LAUNCH();

  // 16 warps per CTA split the column tiles between two warp groups
  // (measured: 31-35 us per pass against 22 us with 8 warps -- both groups repeat the ldmatrix / convert work on
  //  the streamed operand, which is the larger half of the loop; kept as an opt-in for A/B: CF_LR_WARPS=16)
  static const bool wide = [] { const char* e = getenv("CF_LR_WARPS"); return e && e[0] == '1' && e[1] == '6'; }();
  const int gemm_threads = (RP >= 16 && wide) ? 2 * kLrThreads : kLrThreads;
  const bool small = lr_gemm_small();
  const size_t smem_ns = lr_gemm_smem<RP, false, 32, 2>(), smem_ts = lr_gemm_smem<RP, true, 32, 2>();
  constexpr bool kHalfOk = RP >= 16;   // (one ldmatrix.trans covers two 8-column tiles)
  if (small) {
    CF_CHECK_CUDA(cudaFuncSetAttribute(k_lr_gemm<RP, false, 32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_ns)));
    CF_CHECK_CUDA(cudaFuncSetAttribute(k_lr_gemm<RP, true, 32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_ts)));
  }
  auto gemm_AQ = [&](unsigned* absmax = nullptr) -> cudaError_t {  // part[s] (N, RP) = A Q
    dim3 grid((n + kLrBM - 1) / kLrBM, p.aq_splits);
    if (hb) {
      switch (lr_hb_tile()) {
        case 1: return launch_lr_gemm<RP, false, 32, 3, kHalfOk>(grid, st, xh, bh, Q2, part, n, c, p.aq_kper, absmax);
        case 2: return launch_lr_gemm<RP, false, 64, 2, kHalfOk>(grid, st, xh, bh, Q2, part, n, c, p.aq_kper, absmax);
        case 3: return launch_lr_gemm<RP, false, 64, 3, kHalfOk>(grid, st, xh, bh, Q2, part, n, c, p.aq_kper, absmax);
        default: return launch_lr_gemm<RP, false, 32, 2, kHalfOk>(grid, st, xh, bh, Q2, part, n, c, p.aq_kper, absmax);
      }
    }
    if (small)
      k_lr_gemm<RP, false, 32, 2><<<grid, kLrThreads, smem_ns, st>>>(xh, bh, Q2, part, n, c, p.aq_kper, nullptr);
    else
      k_lr_gemm<RP, false><<<grid, gemm_threads, smem_n, st>>>(xh, bh, Q2, part, n, c, p.aq_kper, nullptr);
    return cudaGetLastError();
  };
  auto gemm_AtY = [&](bool planes = false) -> cudaError_t {  // part[s] (C, RP) = A^T Y; planes: Y2 holds fp16 planes
    dim3 grid((c + kLrBM - 1) / kLrBM, p.aty_splits);
    if (planes) {
      switch (lr_hb_tile()) {
        case 1: return launch_lr_gemm<RP, true, 32, 3, kHalfOk>(grid, st, xh, bh, Y2, part, n, c, p.aty_kper, nullptr);
        case 2: return launch_lr_gemm<RP, true, 64, 2, kHalfOk>(grid, st, xh, bh, Y2, part, n, c, p.aty_kper, nullptr);
        case 3: return launch_lr_gemm<RP, true, 64, 3, kHalfOk>(grid, st, xh, bh, Y2, part, n, c, p.aty_kper, nullptr);
        default: return launch_lr_gemm<RP, true, 32, 2, kHalfOk>(grid, st, xh, bh, Y2, part, n, c, p.aty_kper, nullptr);
      }
    }
    if (small)
      k_lr_gemm<RP, true, 32, 2><<<grid, kLrThreads, smem_ts, st>>>(xh, bh, Y2, part, n, c, p.aty_kper, nullptr);
    else
      k_lr_gemm<RP, true><<<grid, gemm_threads, smem_t, st>>>(xh, bh, Y2, part, n, c, p.aty_kper, nullptr);
    return cudaGetLastError();
  };
  // CholeskyQR2 of (sum of `S` partials, M x RP): result as TF32 pairs in out2 (+ fp16 / compact fp32 copies).
  // One launch on one 8-CTA cluster (k_lr_orth); CF_LR_ORTH=legacy keeps round 1's five-launch chain (A/B).
  static const bool legacy_orth = [] { const char* e = getenv("CF_LR_ORTH"); return e && e[0] == 'l'; }();
  const size_t smem_o = lr_orth_smem<RP>();
  // cluster size: 16 CTAs (non-portable, needs the opt-in attribute) when the device can co-schedule them, else 8
  static int orth_nc = 0;
  if (!legacy_orth && orth_nc == 0) {
    orth_nc = 8;
    CF_CHECK_CUDA(cudaFuncSetAttribute(k_lr_orth<RP, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_o)));
    const char* e = getenv("CF_LR_CLUSTER");
    if (!(e && e[0] == '8') &&
        cudaFuncSetAttribute(k_lr_orth<RP, 16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
        cudaFuncSetAttribute(k_lr_orth<RP, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_o)) == cudaSuccess) {
      cudaLaunchConfig_t q{};
      q.gridDim = dim3(16); q.blockDim = dim3(kOrthThreads); q.dynamicSmemBytes = smem_o;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 16; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, k_lr_orth<RP, 16>, &q) == cudaSuccess && nclusters > 0) orth_nc = 16;
    }
    (void)cudaGetLastError();
  }
  if (!legacy_orth) {  // (the attributes are per kernel instantiation: set for every RP this process uses)
    CF_CHECK_CUDA(cudaFuncSetAttribute(k_lr_orth<RP, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_o)));
    if (orth_nc == 16) {
      CF_CHECK_CUDA(cudaFuncSetAttribute(k_lr_orth<RP, 16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      CF_CHECK_CUDA(cudaFuncSetAttribute(k_lr_orth<RP, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_o)));
    }
  }
  static const bool light_ok = [] { const char* e = getenv("CF_LR_LIGHT"); return !(e && e[0] == '0'); }();
  auto orth = [&](int S, int M, int ctas, int rows, float2* out2, __half* out16, float* out32c, bool light = false) -> int {
    if (!legacy_orth) {
      OrthParams o{};
      o.light = (light && light_ok) ? 1 : 0;
      o.half_planes = (hb && out2 != nullptr) ? 1 : 0;
      o.part = part; o.S = S; o.part_stride = static_cast<size_t>(M) * RP; o.X = Xsum; o.M = M; o.r = r;
      o.out2 = out2; o.out16 = out16; o.out32c = out32c;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(orth_nc);
      cfg.blockDim = dim3(kOrthThreads);
      cfg.dynamicSmemBytes = smem_o;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = orth_nc;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (orth_nc == 16)
        CF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k_lr_orth<RP, 16>, o));
      else
        CF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k_lr_orth<RP, 8>, o));
      return CF_OK;
    }
    // the S split-K partials are added by a wide kernel (all SMs); the 16 Gram CTAs then read one copy
    k_lr_sum_split<<<small_grid, 256, 0, st>>>(part, S, static_cast<size_t>(M) * RP, nullptr, Xsum,
                                               static_cast<size_t>(M) * RP, nullptr);
    CF_CHECK_LAUNCH();
    GramParams g{};
    g.xpart = Xsum; g.S = 1; g.part_stride = static_cast<size_t>(M) * RP; g.X = Xsum;
    g.M = M; g.r = r; g.rows_per_cta = rows; g.gpart = gpart; g.ticket = ticket; g.r_out = rfac; g.rdinv_out = rdinv;
    k_lr_gram_chol<RP><<<ctas, 256, 0, st>>>(g);
    CF_CHECK_LAUNCH();
    k_lr_solve_out<RP><<<(M + 127) / 128, 128, 0, st>>>(Xsum, rfac, rdinv, M, r, nullptr, nullptr, nullptr, 0);
    CF_CHECK_LAUNCH();
    g.xpart = Xsum; g.S = 1;
    k_lr_gram_chol<RP><<<ctas, 256, 0, st>>>(g);
    CF_CHECK_LAUNCH();
    k_lr_solve_out<RP><<<(M + 127) / 128, 128, 0, st>>>(Xsum, rfac, rdinv, M, r, out2, out16, out32c, hb ? 1 : 0);
    CF_CHECK_LAUNCH();
    return CF_OK;
  };

  bool q_written = false;
  for (int it = 0; it < iters; ++it) {
    // the raw Y = A Q of an iteration has no a-priori bound: the product kernel records max |partial| and the sum
    // kernel scales Y by a power of two before splitting it into fp16 planes (span(A^T Y) is unchanged)
    unsigned* amax = (hb && it < 1000) ? ticket + 1 + it : nullptr;
    CF_CHECK_CUDA(gemm_AQ(amax));
    k_lr_sum_split<<<small_grid, 256, 0, st>>>(part, p.aq_splits, static_cast<size_t>(n) * RP, Y2, nullptr,
                                               static_cast<size_t>(n) * RP, amax);
    CF_CHECK_LAUNCH();
    CF_CHECK_CUDA(gemm_AtY(amax != nullptr));
    const bool last = it == iters - 1;
    // the bases between iterations only seed the next product; the one handed back to the caller (q_out) and U are
    // orthonormalised in full
    if (int rc = orth(p.aty_splits, c, p.gram_ctas_c, p.gram_rows_c, Q2, nullptr, (last && q_out) ? q_out : nullptr,
                      !(last && q_out)))
      return rc;
    q_written = q_written || (last && q_out);
  }
  if (q_out != nullptr && !q_written)
    CF_CHECK_CUDA(cudaMemcpyAsync(q_out, q0, static_cast<size_t>(c) * r * 4, cudaMemcpyDeviceToDevice, st));
  CF_CHECK_CUDA(gemm_AQ());  // U_temp = A Q
  if (int rc = orth(p.aq_splits, n, p.gram_ctas_n, p.gram_rows_n, Y2, U, nullptr)) return rc;  // U = orth(A Q)
  CF_CHECK_CUDA(gemm_AtY(hb));  // V^T = A^T U
  k_lr_store_v_sum<<<small_grid, 256, 0, st>>>(part, p.aty_splits, static_cast<size_t>(c) * RP, V, c, RP, r);
  CF_CHECK_LAUNCH();
  return CF_OK;
}

static size_t lr_mma_total(int64_t N, int64_t C, int r) { return make_lr_mma_plan(N, C, r).total; }

static bool lr_mma_eligible(const void* x, const void* base, int64_t N, int64_t C) {
  const char* e = getenv("CF_LEGACY_KERNELS");
  if (e && e[0] == '1') return false;
  return C % 8 == 0 && C >= 8 && N >= 1 && aligned16(x) && (!base || aligned16(base));
}

}  // namespace cf

extern "C" {

int cf_lowrank_project(const void* x, const void* base, const float* q0, void* U, void* V, float* q_out, int64_t N,
                       int64_t C, int rank, int iters, void* workspace, size_t workspace_bytes,
                       cf_stream_t stream) {
  using namespace cf;
  CF_CHECK_ARG(x && q0 && U && V, "null pointer");
  CF_CHECK_ARG(rank >= 1 && rank <= kMaxRank, "rank %d out of range [1, %d]", rank, kMaxRank);
  CF_CHECK_ARG(iters >= 0 && iters <= 1000, "iters out of range");
  CF_CHECK_ARG(N >= 1 && C >= 1 && N < (int64_t(1) << 31) && C < (int64_t(1) << 31), "bad shape");
  CF_CHECK_ARG(rank <= N && rank <= C, "rank larger than the matrix");
  CF_CHECK_ARG(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0,
               "workspace must be non-null and 256-byte aligned");
  if (lr_mma_eligible(x, base, N, C)) {
    const LrMmaPlan mp = make_lr_mma_plan(N, C, rank);
    if (mp.total > workspace_bytes) {
      set_error("workspace too small: need %zu bytes, got %zu", mp.total, workspace_bytes);
      return CF_ERR_WORKSPACE;
    }
    const __half* xh2 = static_cast<const __half*>(x);
    const __half* bh2 = static_cast<const __half*>(base);
    char* ws2 = static_cast<char*>(workspace);
    cudaStream_t st2 = static_cast<cudaStream_t>(stream);
    const int n2 = static_cast<int>(N), c2 = static_cast<int>(C);
    switch (mp.RP) {
      case 8: return lr_mma_project<8>(xh2, bh2, q0, static_cast<__half*>(U), static_cast<__half*>(V), q_out, n2, c2, rank, iters, ws2, mp, st2);
      case 16: return lr_mma_project<16>(xh2, bh2, q0, static_cast<__half*>(U), static_cast<__half*>(V), q_out, n2, c2, rank, iters, ws2, mp, st2);
      case 32: return lr_mma_project<32>(xh2, bh2, q0, static_cast<__half*>(U), static_cast<__half*>(V), q_out, n2, c2, rank, iters, ws2, mp, st2);
      default: return lr_mma_project<64>(xh2, bh2, q0, static_cast<__half*>(U), static_cast<__half*>(V), q_out, n2, c2, rank, iters, ws2, mp, st2);
    }
  }
  const LrPlan p = make_lr_plan(N, C, rank);
  if (p.total > workspace_bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", p.total, workspace_bytes);
    return CF_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  float* Q = reinterpret_cast<float*>(ws + p.q_off);
  float* Y = reinterpret_cast<float*>(ws + p.y_off);
  float* Zp = reinterpret_cast<float*>(ws + p.zpart_off);
  double* gpart = reinterpret_cast<double*>(ws + p.gpart_off);
  float* rinv = reinterpret_cast<float*>(ws + p.rinv_off);
  const __half* xh = static_cast<const __half*>(x);
  const __half* bh = static_cast<const __half*>(base);
  const int n = static_cast<int>(N), c = static_cast<int>(C), r = rank, RP = p.RP;
  const int small_grid = 2 * sm_count();

  k_lr_pad_copy<<<small_grid, 256, 0, st>>>(q0, Q, c, r, RP, false);
  CF_CHECK_LAUNCH();
  auto AQ = [&](const float* q, float* y) {
    switch (RP) {
      case 8: launch_AQ<8>(xh, bh, q, y, n, c, st); break;
      case 16: launch_AQ<16>(xh, bh, q, y, n, c, st); break;
      case 32: launch_AQ<32>(xh, bh, q, y, n, c, st); break;
      default: launch_AQ<64>(xh, bh, q, y, n, c, st); break;
    }
  };
  auto AtY = [&](const float* y, float* z) {
    switch (RP) {
      case 8: launch_AtY<8>(xh, bh, y, Zp, n, c, p, st); break;
      case 16: launch_AtY<16>(xh, bh, y, Zp, n, c, p, st); break;
      case 32: launch_AtY<32>(xh, bh, y, Zp, n, c, p, st); break;
      default: launch_AtY<64>(xh, bh, y, Zp, n, c, p, st); break;
    }
    k_lr_sum_parts<<<small_grid, 256, 0, st>>>(Zp, z, static_cast<size_t>(c) * RP, p.S);
  };
  for (int it = 0; it < iters; ++it) {
    AQ(Q, Y);                 // Y = A Q
    CF_CHECK_LAUNCH();
    AtY(Y, Q);                // Z = A^T Y   (stored over Q)
    CF_CHECK_LAUNCH();
    if (int rc = orthonormalise(Q, c, RP, r, gpart, rinv, nullptr, st)) return rc;  // Q = orth(Z)
  }
  AQ(Q, Y);                   // U_temp = A Q
  CF_CHECK_LAUNCH();
  if (int rc = orthonormalise(Y, n, RP, r, gpart, rinv, static_cast<__half*>(U), st)) return rc;  // U = orth(A Q)
  if (q_out != nullptr) {
    k_lr_pad_copy<<<small_grid, 256, 0, st>>>(Q, q_out, c, r, RP, true);
    CF_CHECK_LAUNCH();
  }
  AtY(Y, Q);                  // V^T = A^T U  (Q is free now)
  CF_CHECK_LAUNCH();
  k_lr_store_v<<<small_grid, 256, 0, st>>>(Q, static_cast<__half*>(V), c, RP, r);
  CF_CHECK_LAUNCH();
  return CF_OK;
}

int cf_lowrank_q_reconstruct(const void* payload, const void* base, void* recon, int64_t N, int64_t C, int rank,
                             cf_stream_t stream) {
  using namespace cf;
  CF_CHECK_ARG(payload && recon, "null pointer");
  CF_CHECK_ARG(rank >= 1 && rank <= kMaxRank, "rank %d out of range [1, %d]", rank, kMaxRank);
  CF_CHECK_ARG(N >= 2 && N % 2 == 0 && C >= 8 && C % 8 == 0, "N must be even and C a multiple of 8");
  CF_CHECK_ARG(aligned2(payload) && aligned16(recon) && (!base || aligned16(base)), "base/recon must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint8_t* p = static_cast<const uint8_t*>(payload);
  const size_t qu_bytes = static_cast<size_t>(N / 2) * rank, qv_bytes = static_cast<size_t>(C / 2) * rank;
  CF_CHECK_ARG(qu_bytes % 2 == 0 && qv_bytes % 2 == 0, "N * rank and C * rank must be multiples of 4 (fp16 payload)");
  const uint8_t* qU = p;
  const __half* sU = reinterpret_cast<const __half*>(p + qu_bytes);
  const __half* mU = sU + rank;
  const uint8_t* qVt = reinterpret_cast<const uint8_t*>(mU + rank);
  const __half* sV = reinterpret_cast<const __half*>(qVt + qv_bytes);
  const __half* mV = sV + rank;
  const __half* b = static_cast<const __half*>(base);
  __half* o = static_cast<__half*>(recon);
  const int n = static_cast<int>(N), c = static_cast<int>(C);
  const int vec = (rank % 16 == 0 && aligned16(qU) && aligned16(qVt) && std::getenv("CF_LRQ_SCALAR") == nullptr) ? 1 : 0;
  dim3 grid(static_cast<unsigned>((C + 127) / 128), static_cast<unsigned>((N + 63) / 64));
  switch ((rank + 15) / 16) {
    case 1: k_lrq_reconstruct<1><<<grid, 128, 0, st>>>(qU, sU, mU, qVt, sV, mV, b, o, n, c, rank, vec); break;
    case 2: k_lrq_reconstruct<2><<<grid, 128, 0, st>>>(qU, sU, mU, qVt, sV, mV, b, o, n, c, rank, vec); break;
    case 3: k_lrq_reconstruct<3><<<grid, 128, 0, st>>>(qU, sU, mU, qVt, sV, mV, b, o, n, c, rank, vec); break;
    default: k_lrq_reconstruct<4><<<grid, 128, 0, st>>>(qU, sU, mU, qVt, sV, mV, b, o, n, c, rank, vec); break;
  }
  CF_CHECK_LAUNCH();
  return CF_OK;
}

int cf_lowrank_reconstruct(const void* U, const void* V, const void* base, void* recon, int64_t N, int64_t C, int rank,
                           cf_stream_t stream) {
  using namespace cf;
  CF_CHECK_ARG(U && V && recon, "null pointer");
  CF_CHECK_ARG(rank >= 1 && rank <= kMaxRank, "rank %d out of range [1, %d]", rank, kMaxRank);
  CF_CHECK_ARG(N >= 1 && C >= 8 && C % 8 == 0, "C must be a multiple of 8");
  CF_CHECK_ARG(aligned16(V) && aligned16(recon) && (!base || aligned16(base)) && aligned2(U),
               "V/base/recon must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    const char* e = getenv("CF_LEGACY_KERNELS");
    if (!(e && e[0] == '1')) {  // tensor-core path
      dim3 mgrid(static_cast<unsigned>((C + 255) / 256), static_cast<unsigned>((N + 63) / 64));
      const __half* u = static_cast<const __half*>(U);
      const __half* v = static_cast<const __half*>(V);
      const __half* b = static_cast<const __half*>(base);
      __half* o = static_cast<__half*>(recon);
      const int n = static_cast<int>(N), c = static_cast<int>(C);
      static const bool v2 = [] { const char* e2 = getenv("CF_LR_RECON"); return !(e2 && e2[0] == '1'); }();
      if (v2 && rank % 8 == 0 && aligned16(U)) {  // prefetching version (CF_LR_RECON=1: round 1's kernel, A/B)
        dim3 g2(static_cast<unsigned>((C + 127) / 128), static_cast<unsigned>((N + 63) / 64));
        switch ((rank + 15) / 16) {
          case 1: k_lr_reconstruct_v2<1><<<g2, 128, 0, st>>>(u, v, b, o, n, c, rank); break;
          case 2: k_lr_reconstruct_v2<2><<<g2, 128, 0, st>>>(u, v, b, o, n, c, rank); break;
          case 3: k_lr_reconstruct_v2<3><<<g2, 128, 0, st>>>(u, v, b, o, n, c, rank); break;
          default: k_lr_reconstruct_v2<4><<<g2, 128, 0, st>>>(u, v, b, o, n, c, rank); break;
        }
        CF_CHECK_LAUNCH();
        return CF_OK;
      }
      switch ((rank + 15) / 16) {
        case 1: k_lr_reconstruct_mma<1><<<mgrid, 128, 0, st>>>(u, v, b, o, n, c, rank); break;
        case 2: k_lr_reconstruct_mma<2><<<mgrid, 128, 0, st>>>(u, v, b, o, n, c, rank); break;
        case 3: k_lr_reconstruct_mma<3><<<mgrid, 128, 0, st>>>(u, v, b, o, n, c, rank); break;
        default: k_lr_reconstruct_mma<4><<<mgrid, 128, 0, st>>>(u, v, b, o, n, c, rank); break;
      }
      CF_CHECK_LAUNCH();
      return CF_OK;
    }
  }
  dim3 grid(static_cast<unsigned>((C + 255) / 256), static_cast<unsigned>((N + 31) / 32));
  const size_t smem = static_cast<size_t>(rank) * 256 * 2 + 32 * static_cast<size_t>(rank) * 2;
  k_lr_reconstruct<<<grid, 256, smem, st>>>(static_cast<const __half*>(U), static_cast<const __half*>(V),
                                            static_cast<const __half*>(base), static_cast<__half*>(recon),
                                            static_cast<int>(N), static_cast<int>(C), rank);
  CF_CHECK_LAUNCH();
  return CF_OK;
}

}  // extern "C"
