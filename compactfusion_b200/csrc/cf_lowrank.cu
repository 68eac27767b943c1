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
#include "cf_common.cuh"

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
size_t lowrank_workspace_bytes(int64_t N, int64_t C, int rank) {
  if (rank < 1 || rank > kMaxRank) return 256;
  return make_lr_plan(N, C, rank).total;
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
  const LrPlan p = make_lr_plan(N, C, rank);
  CF_CHECK_ARG(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0,
               "workspace must be non-null and 256-byte aligned");
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

int cf_lowrank_reconstruct(const void* U, const void* V, const void* base, void* recon, int64_t N, int64_t C, int rank,
                           cf_stream_t stream) {
  using namespace cf;
  CF_CHECK_ARG(U && V && recon, "null pointer");
  CF_CHECK_ARG(rank >= 1 && rank <= kMaxRank, "rank %d out of range [1, %d]", rank, kMaxRank);
  CF_CHECK_ARG(N >= 1 && C >= 8 && C % 8 == 0, "C must be a multiple of 8");
  CF_CHECK_ARG(aligned16(V) && aligned16(recon) && (!base || aligned16(base)) && aligned2(U),
               "V/base/recon must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(static_cast<unsigned>((C + 255) / 256), static_cast<unsigned>((N + 31) / 32));
  const size_t smem = static_cast<size_t>(rank) * 256 * 2 + 32 * static_cast<size_t>(rank) * 2;
  k_lr_reconstruct<<<grid, 256, smem, st>>>(static_cast<const __half*>(U), static_cast<const __half*>(V),
                                            static_cast<const __half*>(base), static_cast<__half*>(recon),
                                            static_cast<int>(N), static_cast<int>(C), rank);
  CF_CHECK_LAUNCH();
  return CF_OK;
}

}  // extern "C"
