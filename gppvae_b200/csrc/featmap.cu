// Vmodel kernels: row normalisation (vmod.py:10-12) and the row-wise Khatri-Rao feature map
// V[i, j*q+k] = xn[d_i, j] * wn[w_i, k] (vmod.py:28-35), forward and backward.
//
// The forward map is a pure HBM-write stream (4*Q bytes per row; tables are L2 resident): one warp
// per row, the two table rows staged in shared memory, 128-bit coalesced stores, persistent
// grid-stride launch sized from the SM count.
#include <math.h>

#include "common.cuh"

namespace gpp {

constexpr int kWarpsPerBlock = 8;

// ---------------------------------------------------------------- normalize_rows
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
normalize_rows_fwd_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const float* xr = x + r * cols;
    float s = 0.f;
    for (int64_t c = lane; c < cols; c += 32) {
      const float v = xr[c];
      s = fmaf(v, v, s);
    }
    s = sqrtf(warp_sum(s));
    for (int64_t c = lane; c < cols; c += 32) y[r * cols + c] = xr[c] / s;
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
normalize_rows_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, int64_t rows, int64_t cols,
                          float* __restrict__ gx) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const float* xr = x + r * cols;
    const float* gr = gy + r * cols;
    float s = 0.f, dot = 0.f;
    for (int64_t c = lane; c < cols; c += 32) {
      const float v = xr[c];
      s = fmaf(v, v, s);
      dot = fmaf(v, gr[c], dot);
    }
    s = warp_sum(s);
    dot = warp_sum(dot);
    const float inv = 1.f / sqrtf(s);
    const float proj = dot / s;  // (y . gy) / |x|  with y = x / |x|
    for (int64_t c = lane; c < cols; c += 32) gx[r * cols + c] = (gr[c] - xr[c] * proj) * inv;
  }
}

// ---------------------------------------------------------------- Khatri-Rao forward
// dynamic smem: kWarpsPerBlock * (p + q) floats.
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
khatri_rao_fwd_kernel(const float* __restrict__ xn, int64_t P, int p, const float* __restrict__ wn, int64_t nviews,
                      int q, const int64_t* __restrict__ d, const int64_t* __restrict__ w, int64_t n,
                      float* __restrict__ V, int64_t ldv) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  float* xs = smem + (size_t)wib * (p + q);
  float* ws = xs + p;
  const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + wib;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  const int Q = p * q;
  const int Q4 = Q >> 2;
  const float qnan = __int_as_float(0x7fc00000);

  int64_t di = 0, wi = 0;
  if (warp < n) {
    di = d[warp];
    wi = w[warp];
  }
  for (int64_t r = warp; r < n; r += nwarps) {
    const int64_t dcur = di, wcur = wi;
    const int64_t rn = r + nwarps;
    if (rn < n) {  // prefetch the next row's indices behind this row's stores
      di = d[rn];
      wi = w[rn];
    }
    const bool ok = (dcur >= 0) & (dcur < P) & (wcur >= 0) & (wcur < nviews);
    __syncwarp();
    if (ok) {
      for (int j = lane; j < p; j += 32) xs[j] = xn[dcur * p + j];
      for (int k = lane; k < q; k += 32) ws[k] = wn[wcur * q + k];
    }
    __syncwarp();
    float4* vrow = reinterpret_cast<float4*>(V + r * ldv);
    const int q4 = q >> 2;
    if ((q & 3) == 0 && (p & 3) == 0 && q4 <= 32 && (32 % q4) == 0) {
      // fast path (q = 4, 8, 16, ..., 128): a lane's four columns share one j and its view quad never changes, so a
      // float4 of the row is one shared-memory read of x, four multiplies and one 128-bit store
      const int kq = lane % q4, jstep = 32 / q4;
      const float4 w4 = ok ? *reinterpret_cast<const float4*>(ws + 4 * kq) : make_float4(qnan, qnan, qnan, qnan);
      int j = lane / q4;
      for (int c4 = lane; c4 < Q4; c4 += 32, j += jstep) {
        const float x = ok ? xs[j] : qnan;
        vrow[c4] = make_float4(x * w4.x, x * w4.y, x * w4.z, x * w4.w);
      }
    } else {
      for (int c4 = lane; c4 < Q4; c4 += 32) {
        const int c = c4 << 2;
        int j = c / q;
        int k = c - j * q;
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          o[e] = ok ? xs[j] * ws[k] : qnan;
          if (++k == q) {
            k = 0;
            ++j;
          }
        }
        vrow[c4] = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// ---------------------------------------------------------------- Khatri-Rao backward
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
khatri_rao_bwd_kernel(const float* __restrict__ gV, int64_t ldg, const float* __restrict__ xn, int64_t P, int p,
                      const float* __restrict__ wn, int64_t nviews, int q, const int64_t* __restrict__ d,
                      const int64_t* __restrict__ w, int64_t n, float* __restrict__ gxn, float* __restrict__ gwn) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t r = warp; r < n; r += nwarps) {
    const int64_t di = d[r], wi = w[r];
    if (di < 0 || di >= P || wi < 0 || wi >= nviews) continue;
    const float* g = gV + r * ldg;
    const float* xr = xn + di * p;
    const float* wr = wn + wi * q;
    for (int j = lane; j < p; j += 32) {  // d/d xn[d_i, j]
      float acc = 0.f;
      for (int k = 0; k < q; ++k) acc = fmaf(g[j * q + k], wr[k], acc);
      atomicAdd(gxn + di * p + j, acc);
    }
    for (int k = lane; k < q; k += 32) {  // d/d wn[w_i, k]
      float acc = 0.f;
      for (int j = 0; j < p; ++j) acc = fmaf(g[j * q + k], xr[j], acc);
      atomicAdd(gwn + wi * q + k, acc);
    }
  }
}

static int persistent_grid(int64_t rows) {
  const int64_t want = ceil_div(rows, kWarpsPerBlock);
  const int64_t cap = (int64_t)sm_count() * 8;  // 8 resident 256-thread CTAs per SM
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace gpp

using namespace gpp;

extern "C" int gpp_normalize_rows_fwd(const float* x, int64_t rows, int64_t cols, float* y, gpp_stream_t stream) {
  GPP_REQUIRE(x && y, "normalize_rows_fwd: null pointer");
  GPP_REQUIRE(rows >= 0 && cols > 0, "normalize_rows_fwd: bad shape %lld x %lld", (long long)rows, (long long)cols);
  if (rows == 0) return GPP_OK;
  normalize_rows_fwd_kernel<<<persistent_grid(rows), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(x, rows, cols, y);
  GPP_LAUNCH_CHECK();
  return GPP_OK;
}

extern "C" int gpp_normalize_rows_bwd(const float* x, const float* gy, int64_t rows, int64_t cols, float* gx,
                                      gpp_stream_t stream) {
  GPP_REQUIRE(x && gy && gx, "normalize_rows_bwd: null pointer");
  GPP_REQUIRE(rows >= 0 && cols > 0, "normalize_rows_bwd: bad shape");
  if (rows == 0) return GPP_OK;
  normalize_rows_bwd_kernel<<<persistent_grid(rows), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(x, gy, rows, cols,
                                                                                                      gx);
  GPP_LAUNCH_CHECK();
  return GPP_OK;
}

extern "C" int gpp_khatri_rao_fwd(const float* xn, int64_t P, int32_t p, const float* wn, int64_t nviews, int32_t q,
                                  const int64_t* d, const int64_t* w, int64_t n, float* V, int64_t ldv,
                                  gpp_stream_t stream) {
  GPP_REQUIRE(xn && wn && d && w && V, "khatri_rao_fwd: null pointer");
  GPP_REQUIRE(P > 0 && p > 0 && nviews > 0 && q > 0 && n >= 0, "khatri_rao_fwd: bad shape");
  const int64_t Q = (int64_t)p * q;
  GPP_REQUIRE(Q % 4 == 0, "khatri_rao_fwd: p*q = %lld must be a multiple of 4 (pad p on the host side)", (long long)Q);
  GPP_REQUIRE(ldv >= Q && ldv % 4 == 0 && aligned16(V), "khatri_rao_fwd: V must be 16-byte aligned with ldv %% 4 == 0");
  const size_t smem = (size_t)kWarpsPerBlock * (p + q) * sizeof(float);
  GPP_REQUIRE(smem <= 48 * 1024, "khatri_rao_fwd: p + q = %d too large", p + q);
  if (n == 0) return GPP_OK;
  khatri_rao_fwd_kernel<<<persistent_grid(n), kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(xn, P, p, wn, nviews,
                                                                                                  q, d, w, n, V, ldv);
  GPP_LAUNCH_CHECK();
  return GPP_OK;
}

extern "C" int gpp_khatri_rao_bwd(const float* gV, int64_t ldg, const float* xn, int64_t P, int32_t p, const float* wn,
                                  int64_t nviews, int32_t q, const int64_t* d, const int64_t* w, int64_t n, float* gxn,
                                  float* gwn, gpp_stream_t stream) {
  GPP_REQUIRE(gV && xn && wn && d && w && gxn && gwn, "khatri_rao_bwd: null pointer");
  GPP_REQUIRE(P > 0 && p > 0 && nviews > 0 && q > 0 && n >= 0 && ldg >= (int64_t)p * q, "khatri_rao_bwd: bad shape");
  if (n == 0) return GPP_OK;
  khatri_rao_bwd_kernel<<<persistent_grid(n), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(gV, ldg, xn, P, p, wn,
                                                                                               nviews, q, d, w, n, gxn,
                                                                                               gwn);
  GPP_LAUNCH_CHECK();
  return GPP_OK;
}
