// SIMT fp32 GEMM kernels built on gemm_simt.cuh:
//   * tn_partial_kernel / tn_reduce_kernel : out = A^T [B1 | B2] with a long row contraction,
//     deterministic split-K (fixed-order reduction), optional SYRK symmetry (A == B1);
//   * gemm_kernel<A_ROWC, B_ROWC>          : generic batched C = alpha * A.B (+ beta * C) with
//     lower-triangle-only tiles and triangular k-range hints (Cholesky updates, triangular inverse);
//   * xb_kernel / xb_finalize_kernel       : pass 2, Xb = (X - V W)/vn with the NLL epilogue;
//   * vb_kernel                            : Vb = (v0/vn) L V B^-1 - Xb W^T.
#include "gemm_simt.cuh"
#include "kernels.h"

namespace gpp {

// =====================================================================================
// TN: out(ka x (kb1+kb2)) = sum over rows of A(row, :)^T [B1(row, :) | B2(row, :)]
// =====================================================================================
struct TnParams {
  const float* A;  int64_t lda; int ka;
  const float* B1; int64_t ldb1; int kb1;
  const float* B2; int64_t ldb2; int kb2;
  int64_t n;               // rows contracted
  int symmetric;           // B1 == A: only tiles with tile_n <= tile_m are computed
  int splits;
  int64_t rows_per_split;  // multiple of BK
  float* partial;          // [tile][split][BM*BN]
  int tm, tn1, tn2;        // tile counts
  int ntiles1;             // tiles in the B1 part
};

__device__ __forceinline__ void tn_decode_tile(const TnParams& p, int tile, int& tmi, int& tni, bool& second) {
  if (tile < p.ntiles1) {
    second = false;
    if (p.symmetric) {
      // tile = tmi*(tmi+1)/2 + tni, tni <= tmi
      int t = (int)((sqrtf(8.f * (float)tile + 1.f) - 1.f) * 0.5f);
      while ((t + 1) * (t + 2) / 2 <= tile) ++t;
      while (t * (t + 1) / 2 > tile) --t;
      tmi = t;
      tni = tile - t * (t + 1) / 2;
    } else {
      tmi = tile / p.tn1;
      tni = tile - tmi * p.tn1;
    }
  } else {
    second = true;
    const int t2 = tile - p.ntiles1;
    tmi = t2 / p.tn2;
    tni = t2 - tmi * p.tn2;
  }
}

constexpr int64_t kFlushRows = 1024;  // second-level accumulation period (bounds fp32 running-sum error)

__global__ void __launch_bounds__(kGemmThreads, 2) tn_partial_kernel(TnParams p) {
  __shared__ TileSmem sm;
  const int tile = blockIdx.x / p.splits;
  const int split = blockIdx.x - tile * p.splits;
  int tmi, tni;
  bool second;
  tn_decode_tile(p, tile, tmi, tni, second);

  const int64_t r0 = (int64_t)split * p.rows_per_split;
  int64_t r1 = r0 + p.rows_per_split;
  if (r1 > p.n) r1 = p.n;

  Operand A, B;
  A.ld = p.lda;
  A.mn_valid = min(BM, p.ka - tmi * BM);
  if (!second) {
    B.ld = p.ldb1;
    B.mn_valid = min(BN, p.kb1 - tni * BN);
  } else {
    B.ld = p.ldb2;
    B.mn_valid = min(BN, p.kb2 - tni * BN);
  }
  const float* Bsrc = second ? p.B2 : p.B1;

  float* out = p.partial + ((size_t)tile * p.splits + split) * (size_t)(BM * BN);
  float acc[8][8];
  bool first = true;
  for (int64_t rs = r0; rs < r1 || first; rs += kFlushRows) {
    int64_t len = r1 - rs;
    if (len > kFlushRows) len = kFlushRows;
    if (len < 0) len = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    A.base = p.A + rs * p.lda + (int64_t)tmi * BM;
    B.base = Bsrc + rs * B.ld + (int64_t)tni * BN;
    tile_mainloop<true, true>(A, B, len, sm, acc);
    // flush: the partial tile is private to this CTA, so plain read-modify-write is race free
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float* orow = out + acc_row(i) * BN;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        float4* o4 = reinterpret_cast<float4*>(orow + acc_col(jj * 4));
        float4 v = make_float4(acc[i][jj * 4 + 0], acc[i][jj * 4 + 1], acc[i][jj * 4 + 2], acc[i][jj * 4 + 3]);
        if (!first) {
          const float4 old = *o4;
          v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
        }
        *o4 = v;
      }
    }
    first = false;
  }
}

// out[m][n] = sum_s partial[tile(m,n)][s]; B2 part optionally scaled by v0/vn (scal != nullptr).
struct TnReduceParams {
  const float* partial;
  int splits;
  int ka, kb1, kb2;
  int symmetric;
  int tm, tn1, tn2, ntiles1;
  float* out; int64_t ldo;      // B1 part at column 0
  float* out2; int64_t ldo2;    // B2 part at column 0 of out2
  const double* scal;           // when set: B2 part *= scal[V0]/scal[VN]
};

__global__ void __launch_bounds__(256) tn_reduce_kernel(TnReduceParams p) {
  const int tile = blockIdx.x;
  TnParams q;  // reuse the decoder
  q.symmetric = p.symmetric; q.tn1 = p.tn1; q.tn2 = p.tn2; q.ntiles1 = p.ntiles1;
  int tmi, tni;
  bool second;
  tn_decode_tile(q, tile, tmi, tni, second);
  const int mv = min(BM, p.ka - tmi * BM);
  const int nv = min(BN, (second ? p.kb2 : p.kb1) - tni * BN);
  double scale = 1.0;
  if (second && p.scal) scale = p.scal[GPP_S_V0] / p.scal[GPP_S_VN];
  const float* src = p.partial + (size_t)tile * p.splits * (size_t)(BM * BN);
  for (int e = threadIdx.x; e < BM * BN / 4; e += blockDim.x) {
    const int m = e / (BN / 4), n4 = (e - m * (BN / 4)) * 4;
    if (m >= mv || n4 >= nv) continue;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int s = 0; s < p.splits; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(src + (size_t)s * (BM * BN) + m * BN + n4);
      s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
    }
    const float4 r = make_float4((float)(s0 * scale), (float)(s1 * scale), (float)(s2 * scale), (float)(s3 * scale));
    const int gm = tmi * BM + m, gn = tni * BN + n4;
    if (!second) {
      *reinterpret_cast<float4*>(p.out + (int64_t)gm * p.ldo + gn) = r;
      if (p.symmetric && tmi != tni) {  // mirror into the upper triangle
        p.out[(int64_t)(gn + 0) * p.ldo + gm] = r.x;
        p.out[(int64_t)(gn + 1) * p.ldo + gm] = r.y;
        p.out[(int64_t)(gn + 2) * p.ldo + gm] = r.z;
        p.out[(int64_t)(gn + 3) * p.ldo + gm] = r.w;
      }
    } else {
      *reinterpret_cast<float4*>(p.out2 + (int64_t)gm * p.ldo2 + gn) = r;
    }
  }
}

static void tn_geometry(int64_t n, int ka, int kb1, int kb2, int symmetric, int& tm, int& tn1, int& tn2, int& ntiles1,
                        int& ntiles, int& splits, int64_t& rows_per_split) {
  tm = (int)ceil_div(ka, BM);
  tn1 = (int)ceil_div(kb1, BN);
  tn2 = (int)ceil_div(kb2, BN);
  ntiles1 = symmetric ? tm * (tm + 1) / 2 : tm * tn1;
  ntiles = ntiles1 + tm * tn2;
  // aim for ~3 waves of 2 CTAs/SM, at least 256 rows per split, at most 64 splits
  const int64_t target = (int64_t)sm_count() * 2 * 3;
  int64_t s = ceil_div(target, ntiles > 0 ? ntiles : 1);
  const int64_t smax = n / 256 > 1 ? n / 256 : 1;
  if (s > smax) s = smax;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  rows_per_split = ceil_div(ceil_div(n, s), BK) * BK;
  if (rows_per_split < BK) rows_per_split = BK;
  splits = (int)ceil_div(n > 0 ? n : 1, rows_per_split);
  if (splits < 1) splits = 1;
}

size_t tn_workspace_bytes(int64_t n, int ka, int kb1, int kb2, int symmetric) {
  int tm, tn1, tn2, nt1, nt, splits;
  int64_t rps;
  tn_geometry(n, ka, kb1, kb2, symmetric, tm, tn1, tn2, nt1, nt, splits, rps);
  return (size_t)nt * splits * BM * BN * sizeof(float);
}

int launch_tn(const float* A, int64_t lda, int ka, const float* B1, int64_t ldb1, int kb1, const float* B2,
              int64_t ldb2, int kb2, int64_t n, int symmetric, float* out, int64_t ldo, float* out2, int64_t ldo2,
              const double* scal_for_b2, void* ws, size_t ws_bytes, cudaStream_t st) {
  TnParams p;
  p.A = A; p.lda = lda; p.ka = ka;
  p.B1 = B1; p.ldb1 = ldb1; p.kb1 = kb1;
  p.B2 = B2; p.ldb2 = ldb2; p.kb2 = kb2;
  p.n = n; p.symmetric = symmetric;
  int ntiles;
  tn_geometry(n, ka, kb1, kb2, symmetric, p.tm, p.tn1, p.tn2, p.ntiles1, ntiles, p.splits, p.rows_per_split);
  const size_t need = (size_t)ntiles * p.splits * BM * BN * sizeof(float);
  if (ws_bytes < need || ws == nullptr) {
    set_error("A^T B: workspace too small (%zu < %zu bytes)", ws_bytes, need);
    return GPP_ERR_WORKSPACE;
  }
  if (ntiles == 0) return GPP_OK;
  p.partial = static_cast<float*>(ws);
  tn_partial_kernel<<<ntiles * p.splits, kGemmThreads, 0, st>>>(p);
  GPP_LAUNCH_CHECK();
  TnReduceParams r;
  r.partial = p.partial; r.splits = p.splits; r.ka = ka; r.kb1 = kb1; r.kb2 = kb2; r.symmetric = symmetric;
  r.tm = p.tm; r.tn1 = p.tn1; r.tn2 = p.tn2; r.ntiles1 = p.ntiles1;
  r.out = out; r.ldo = ldo; r.out2 = out2; r.ldo2 = ldo2; r.scal = scal_for_b2;
  tn_reduce_kernel<<<ntiles, 256, 0, st>>>(r);
  GPP_LAUNCH_CHECK();
  return GPP_OK;
}

// =====================================================================================
// Generic batched GEMM: C[b] = alpha * A[b] . B[b] (+ beta * C[b])
// =====================================================================================
template <bool A_ROWC, bool B_ROWC>
__global__ void __launch_bounds__(kGemmThreads, 2) gemm_kernel(GemmParams p) {
  __shared__ TileSmem sm;
  const int tmi = blockIdx.y, tni = blockIdx.x, b = blockIdx.z;
  if (p.lower_only && tni > tmi) return;
  const int M = (b == (int)gridDim.z - 1 && p.M_last >= 0) ? p.M_last : p.M;
  const int m0 = tmi * BM, n0 = tni * BN;
  if (m0 >= M || n0 >= p.N) return;
  int K = p.K;
  if (p.K_is_M) K = M;
  int kbeg = 0, kend = K;
  if (p.tri_a) kend = min(K, m0 + BM);   // A(m, k) == 0 for k > m
  if (p.tri_b) kbeg = n0;                // B(n, k) == 0 for k < n
  const float* Ab = p.A + (int64_t)b * p.strideA;
  const float* Bb = p.B + (int64_t)b * p.strideB;
  float* Cb = p.C + (int64_t)b * p.strideC;

  Operand A, B;
  A.ld = p.lda; A.mn_valid = min(BM, M - m0);
  B.ld = p.ldb; B.mn_valid = min(BN, p.N - n0);
  A.base = A_ROWC ? Ab + (int64_t)kbeg * p.lda + m0 : Ab + (int64_t)m0 * p.lda + kbeg;
  B.base = B_ROWC ? Bb + (int64_t)kbeg * p.ldb + n0 : Bb + (int64_t)n0 * p.ldb + kbeg;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  if (kend > kbeg) tile_mainloop<A_ROWC, B_ROWC>(A, B, kend - kbeg, sm, acc);

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + acc_row(i);
    if (m >= M) continue;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int n = n0 + acc_col(jj * 4);
      if (n >= p.N) continue;
      float4* c4 = reinterpret_cast<float4*>(Cb + (int64_t)m * p.ldc + n);
      float4 v = make_float4(p.alpha * acc[i][jj * 4 + 0], p.alpha * acc[i][jj * 4 + 1], p.alpha * acc[i][jj * 4 + 2],
                             p.alpha * acc[i][jj * 4 + 3]);
      if (p.beta != 0.f) {
        const float4 old = *c4;
        v.x = fmaf(p.beta, old.x, v.x); v.y = fmaf(p.beta, old.y, v.y);
        v.z = fmaf(p.beta, old.z, v.z); v.w = fmaf(p.beta, old.w, v.w);
      }
      *c4 = v;
    }
  }
}

int launch_gemm(const GemmParams& p, bool a_rowc, bool b_rowc, int batches, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0 || batches <= 0) return GPP_OK;
  dim3 grid((unsigned)ceil_div(p.N, BN), (unsigned)ceil_div(p.M, BM), (unsigned)batches);
  if (a_rowc && b_rowc) gemm_kernel<true, true><<<grid, kGemmThreads, 0, st>>>(p);
  else if (a_rowc && !b_rowc) gemm_kernel<true, false><<<grid, kGemmThreads, 0, st>>>(p);
  else if (!a_rowc && b_rowc) gemm_kernel<false, true><<<grid, kGemmThreads, 0, st>>>(p);
  else gemm_kernel<false, false><<<grid, kGemmThreads, 0, st>>>(p);
  GPP_LAUNCH_CHECK();
  return GPP_OK;
}

// =====================================================================================
// Pass 2: Xb = (X - V W) / vn, per-row quad partials, per-CTA sum Xb^2
// =====================================================================================
struct XbParams {
  const float* V; int64_t ldv;
  const float* X; int64_t ldx;
  const float* W; int64_t ldw;
  int64_t n; int Q, L;
  const double* scal;   // reads VN (or alpha_host when scal == nullptr)
  float alpha_host;
  float* Xb; int64_t ldxb;
  float* quad_part;     // [tiles_n][n] (nullptr: skip the NLL epilogue)
  double* xb2_part;     // [tiles_m * tiles_n]
  int tiles_n;          // grid is linear: blockIdx.x = tile_m * tiles_n + tile_n
};

__global__ void __launch_bounds__(kGemmThreads, 2) xb_kernel(XbParams p) {
  __shared__ TileSmem sm;
  __shared__ float red[kGemmThreads / 32];
  const int64_t tmi = blockIdx.x / p.tiles_n;
  const int tni = (int)(blockIdx.x - tmi * p.tiles_n);
  const int64_t m0 = tmi * BM;
  const int n0 = tni * BN;
  Operand A, B;
  A.ld = p.ldv; A.base = p.V + m0 * p.ldv;
  A.mn_valid = (int)min((int64_t)BM, p.n - m0);
  B.ld = p.ldw; B.base = p.W + n0;
  B.mn_valid = min(BN, p.L - n0);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  tile_mainloop<false, true>(A, B, p.Q, sm, acc);

  const float inv_vn = p.scal ? (float)(1.0 / p.scal[GPP_S_VN]) : p.alpha_host;
  float xb2 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + acc_row(i);
    float q = 0.f;
    if (m < p.n) {
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int nn = n0 + acc_col(jj * 4);
        if (nn >= p.L) continue;
        const float4 x = ldg4(p.X + m * p.ldx + nn);
        float4 o;
        o.x = (x.x - acc[i][jj * 4 + 0]) * inv_vn;
        o.y = (x.y - acc[i][jj * 4 + 1]) * inv_vn;
        o.z = (x.z - acc[i][jj * 4 + 2]) * inv_vn;
        o.w = (x.w - acc[i][jj * 4 + 3]) * inv_vn;
        *reinterpret_cast<float4*>(p.Xb + m * p.ldxb + nn) = o;
        q = fmaf(x.x, o.x, q); q = fmaf(x.y, o.y, q); q = fmaf(x.z, o.z, q); q = fmaf(x.w, o.w, q);
        xb2 = fmaf(o.x, o.x, xb2); xb2 = fmaf(o.y, o.y, xb2); xb2 = fmaf(o.z, o.z, xb2); xb2 = fmaf(o.w, o.w, xb2);
      }
    }
    if (p.quad_part) {
      // the 16 threads sharing this row are the 16 lanes of a half-warp (tx = lane & 15)
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      if ((threadIdx.x & 15) == 0 && m < p.n) p.quad_part[(int64_t)tni * p.n + m] = q;
    }
  }
  if (p.xb2_part) {
    xb2 = warp_sum(xb2);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = xb2;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0;
      for (int w = 0; w < kGemmThreads / 32; ++w) s += (double)red[w];
      p.xb2_part[blockIdx.x] = s;
    }
  }
}

// nll_i = 0.5 * sum_t quad_part[t][i] + ROWCONST ; scal[XB2], scal[QUAD] by a fixed-order (deterministic) two-level
// reduction: kXbFinalizeBlocks CTAs each own a contiguous range of rows / partials and leave one double each, a
// single CTA adds those in index order.
__global__ void __launch_bounds__(256) xb_finalize_rows_kernel(const float* __restrict__ quad_part, int tiles_n,
                                                               int64_t n, const double* __restrict__ xb2_part,
                                                               int64_t nparts, const double* __restrict__ scal,
                                                               float* __restrict__ nll, double* __restrict__ fin) {
  __shared__ double red[2][8];
  const double rowconst = scal[GPP_S_ROWCONST];
  const int64_t rows_per = (n + gridDim.x - 1) / gridDim.x, parts_per = (nparts + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per, r1 = min(n, r0 + rows_per);
  const int64_t p0 = (int64_t)blockIdx.x * parts_per, p1 = min(nparts, p0 + parts_per);
  double qsum = 0, xsum = 0;
  for (int64_t i = r0 + threadIdx.x; i < r1; i += blockDim.x) {
    float q = 0.f;
    for (int t = 0; t < tiles_n; ++t) q += quad_part[(int64_t)t * n + i];
    nll[i] = (float)(0.5 * (double)q + rowconst);
    qsum += (double)q;
  }
  for (int64_t i = p0 + threadIdx.x; i < p1; i += blockDim.x) xsum += xb2_part[i];
  qsum = warp_sum(qsum);
  xsum = warp_sum(xsum);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = qsum;
    red[1][threadIdx.x >> 5] = xsum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
    fin[2 * blockIdx.x + 0] = a;
    fin[2 * blockIdx.x + 1] = b;
  }
}

__global__ void xb_finalize_sum_kernel(const double* __restrict__ fin, int nblocks, double* __restrict__ scal) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < nblocks; ++i) { a += fin[2 * i]; b += fin[2 * i + 1]; }
    scal[GPP_S_QUAD] = a;
    scal[GPP_S_XB2] = b;
  }
}

// `fin` = kXbFinalizeBlocks * 2 doubles of workspace (xb_finalize_bytes()).
int launch_xb_finalize(const float* quad_part, int tiles_n, int64_t n, const double* xb2_part, int64_t nparts,
                       double* fin, double* scal, float* nll, cudaStream_t st) {
  xb_finalize_rows_kernel<<<kXbFinalizeBlocks, 256, 0, st>>>(quad_part, tiles_n, n, xb2_part, nparts, scal, nll, fin);
  GPP_LAUNCH_CHECK();
  xb_finalize_sum_kernel<<<1, 32, 0, st>>>(fin, kXbFinalizeBlocks, scal);
  GPP_LAUNCH_CHECK();
  return GPP_OK;
}

size_t xb_workspace_bytes(int64_t n, int L) {
  const int64_t tiles_n = ceil_div(L, BN), tiles_m = ceil_div(n, BM);
  return align_up((size_t)tiles_n * n * sizeof(float), 256) + align_up((size_t)tiles_m * tiles_n * sizeof(double), 256) +
         xb_finalize_bytes();
}

int launch_xb(const float* V, int64_t ldv, const float* X, int64_t ldx, const float* W, int64_t ldw, int64_t n, int Q,
              int L, double* scal, float alpha_host, float* Xb, int64_t ldxb, float* nll, void* ws, size_t ws_bytes,
              cudaStream_t st) {
  XbParams p;
  p.V = V; p.ldv = ldv; p.X = X; p.ldx = ldx; p.W = W; p.ldw = ldw; p.n = n; p.Q = Q; p.L = L;
  p.scal = scal; p.alpha_host = alpha_host; p.Xb = Xb; p.ldxb = ldxb;
  const int64_t tiles_n = ceil_div(L, BN), tiles_m = ceil_div(n, BM);
  if (nll) {
    const size_t need = xb_workspace_bytes(n, L);
    if (ws_bytes < need || !ws) {
      set_error("xb_nll: workspace too small (%zu < %zu bytes)", ws_bytes, need);
      return GPP_ERR_WORKSPACE;
    }
    p.quad_part = static_cast<float*>(ws);
    p.xb2_part = reinterpret_cast<double*>(static_cast<char*>(ws) + align_up((size_t)tiles_n * n * sizeof(float), 256));
  } else {
    p.quad_part = nullptr;
    p.xb2_part = nullptr;
  }
  if (n == 0) return GPP_OK;
  if (tiles_m * tiles_n > 0x7fffffffLL) {
    set_error("xb_nll: too many rows for one launch");
    return GPP_ERR_UNSUPPORTED;
  }
  p.tiles_n = (int)tiles_n;
  xb_kernel<<<(unsigned)(tiles_m * tiles_n), kGemmThreads, 0, st>>>(p);
  GPP_LAUNCH_CHECK();
  if (nll) {
    double* fin = p.xb2_part + align_up((size_t)tiles_m * tiles_n * sizeof(double), 256) / sizeof(double);
    GPP_TRY(launch_xb_finalize(p.quad_part, (int)tiles_n, n, p.xb2_part, tiles_m * tiles_n, fin, scal, nll, st));
  }
  return GPP_OK;
}

// =====================================================================================
// Vb = (v0/vn) L V Binv - Xb W^T
// =====================================================================================
struct VbParams {
  const float* V; int64_t ldv;
  const float* Xb; int64_t ldxb;
  const float* Binv; int64_t ldb;
  const float* W; int64_t ldw;
  const double* scal;
  int64_t n; int Q, L, L_true;
  float* Vb; int64_t ldvb;
  int tiles_n;   // blockIdx.x = tile_m * tiles_n + tile_n
};

__global__ void __launch_bounds__(kGemmThreads, 2) vb_kernel(VbParams p) {
  __shared__ TileSmem sm;
  const int64_t tmi = blockIdx.x / p.tiles_n;
  const int64_t m0 = tmi * BM;
  const int n0 = (int)(blockIdx.x - tmi * p.tiles_n) * BN;
  const int mv = (int)min((int64_t)BM, p.n - m0);
  const int nv = min(BN, p.Q - n0);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  // acc = Xb . W^T   (contraction over the L columns of Xb and of W)
  Operand A, B;
  A.ld = p.ldxb; A.base = p.Xb + m0 * p.ldxb; A.mn_valid = mv;
  B.ld = p.ldw;  B.base = p.W + (int64_t)n0 * p.ldw; B.mn_valid = nv;
  tile_mainloop<false, false>(A, B, p.L, sm, acc);
  const float coef = (float)(p.scal[GPP_S_V0] / p.scal[GPP_S_VN] * (double)p.L_true);
  const float neg_inv = -1.f / coef;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] *= neg_inv;
  // acc += V . Binv   (Binv symmetric; read it row-contracted)
  A.ld = p.ldv; A.base = p.V + m0 * p.ldv;
  B.ld = p.ldb; B.base = p.Binv + n0;
  tile_mainloop<false, true>(A, B, p.Q, sm, acc);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + acc_row(i);
    if (m >= p.n) continue;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int nn = n0 + acc_col(jj * 4);
      if (nn >= p.Q) continue;
      *reinterpret_cast<float4*>(p.Vb + m * p.ldvb + nn) =
          make_float4(coef * acc[i][jj * 4 + 0], coef * acc[i][jj * 4 + 1], coef * acc[i][jj * 4 + 2],
                      coef * acc[i][jj * 4 + 3]);
    }
  }
}

int launch_vb(const float* V, int64_t ldv, const float* Xb, int64_t ldxb, const float* Binv, int64_t ldb,
              const float* W, int64_t ldw, const double* scal, int64_t n, int Q, int L, int L_true, float* Vb,
              int64_t ldvb, cudaStream_t st) {
  if (n == 0) return GPP_OK;
  VbParams p;
  p.V = V; p.ldv = ldv; p.Xb = Xb; p.ldxb = ldxb; p.Binv = Binv; p.ldb = ldb; p.W = W; p.ldw = ldw;
  p.scal = scal; p.n = n; p.Q = Q; p.L = L; p.L_true = L_true; p.Vb = Vb; p.ldvb = ldvb;
  p.tiles_n = (int)ceil_div(Q, BN);
  const int64_t nblk = ceil_div(n, BM) * p.tiles_n;
  if (nblk > 0x7fffffffLL) {
    set_error("vb: too many rows for one launch");
    return GPP_ERR_UNSUPPORTED;
  }
  vb_kernel<<<(unsigned)nblk, kGemmThreads, 0, st>>>(p);
  GPP_LAUNCH_CHECK();
  return GPP_OK;
}

}  // namespace gpp
