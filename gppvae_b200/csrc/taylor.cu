// Taylor surrogate of gp.py:127-133 and its backward, fused into one launch each.
//   fwd: out_i = Xb_i . X_i + Vb_i . V_i + <vbs, softmax(lvs)> / n
//   bwd: gX_i = g_i Xb_i ; gV_i = g_i Vb_i ; glvs = J_softmax^T vbs * sum(g) / n
// n is the minibatch (64 rows in train_gppvae.py:293): launch-latency bound, hence the fusion.
#include "common.cuh"

namespace gpp {

constexpr int kTeWarps = 8;

__device__ __forceinline__ float row_dot(const float* __restrict__ a, const float* __restrict__ b, int len, int lane) {
  float s = 0.f;
  const int len4 = len >> 2;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  for (int c = lane; c < len4; c += 32) {
    const float4 x = a4[c], y = b4[c];
    s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
  }
  return s;
}

__global__ void __launch_bounds__(kTeWarps * 32)
taylor_fwd_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ Xb, int64_t ldxb,
                  const float* __restrict__ V, int64_t ldv, const float* __restrict__ Vb, int64_t ldvb, int64_t n, int L,
                  int Q, const float* __restrict__ vbs, const float* __restrict__ lvs, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * kTeWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kTeWarps;
  double v0, vn;
  softmax2(lvs, v0, vn);
  const float cterm = (float)(((double)vbs[0] * v0 + (double)vbs[1] * vn) / (double)n);
  for (int64_t r = warp; r < n; r += nwarps) {
    float s = row_dot(Xb + r * ldxb, X + r * ldx, L, lane);
    if (Q > 0) s += row_dot(Vb + r * ldvb, V + r * ldv, Q, lane);
    s = warp_sum(s);
    if (lane == 0) out[r] = s + cterm;
  }
}

__global__ void __launch_bounds__(kTeWarps * 32)
taylor_bwd_rows_kernel(const float* __restrict__ gout, const float* __restrict__ Xb, int64_t ldxb,
                       const float* __restrict__ Vb, int64_t ldvb, int64_t n, int L, int Q, float* __restrict__ gX,
                       int64_t ldgx, float* __restrict__ gV, int64_t ldgv) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * kTeWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kTeWarps;
  for (int64_t r = warp; r < n; r += nwarps) {
    const float g = gout[r];
    if (gX) {
      const float4* s4 = reinterpret_cast<const float4*>(Xb + r * ldxb);
      float4* d4 = reinterpret_cast<float4*>(gX + r * ldgx);
      for (int c = lane; c < (L >> 2); c += 32) {
        const float4 v = s4[c];
        d4[c] = make_float4(g * v.x, g * v.y, g * v.z, g * v.w);
      }
    }
    if (gV) {
      const float4* s4 = reinterpret_cast<const float4*>(Vb + r * ldvb);
      float4* d4 = reinterpret_cast<float4*>(gV + r * ldgv);
      for (int c = lane; c < (Q >> 2); c += 32) {
        const float4 v = s4[c];
        d4[c] = make_float4(g * v.x, g * v.y, g * v.z, g * v.w);
      }
    }
  }
}

// glvs_b = (sum g / n) * vs_b * (vbs_b - <vbs, vs>)   (softmax Jacobian of gp.py:50 applied to vbs)
__global__ void __launch_bounds__(256)
taylor_bwd_lvs_kernel(const float* __restrict__ gout, int64_t n, const float* __restrict__ vbs,
                      const float* __restrict__ lvs, float* __restrict__ glvs) {
  __shared__ double red[8];
  double s = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += (double)gout[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += red[w];
    double v0, vn;
    softmax2(lvs, v0, vn);
    const double b0 = (double)vbs[0], b1 = (double)vbs[1];
    const double mean = b0 * v0 + b1 * vn;
    const double scale = t / (double)n;
    glvs[0] = (float)(scale * v0 * (b0 - mean));
    glvs[1] = (float)(scale * vn * (b1 - mean));
  }
}

static int rows_grid(int64_t n) {
  const int64_t want = ceil_div(n, kTeWarps);
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace gpp

using namespace gpp;

extern "C" int gpp_taylor_expansion_fwd(const float* X, int64_t ldx, const float* Xb, int64_t ldxb, const float* V,
                                        int64_t ldv, const float* Vb, int64_t ldvb, int64_t n, int32_t L, int32_t Q,
                                        const float* vbs, const float* lvs, float* out, gpp_stream_t stream) {
  GPP_REQUIRE(X && Xb && vbs && lvs && out, "taylor_expansion_fwd: null pointer");
  GPP_REQUIRE(Q == 0 || (V && Vb), "taylor_expansion_fwd: null V / Vb");
  GPP_REQUIRE(n > 0 && L > 0 && L % 4 == 0 && Q >= 0 && Q % 4 == 0, "taylor_expansion_fwd: bad shape");
  GPP_REQUIRE(ldx % 4 == 0 && ldxb % 4 == 0 && ldv % 4 == 0 && ldvb % 4 == 0, "taylor_expansion_fwd: ld %% 4 != 0");
  GPP_REQUIRE(aligned16(X) && aligned16(Xb) && aligned16(V) && aligned16(Vb), "taylor_expansion_fwd: misaligned");
  taylor_fwd_kernel<<<rows_grid(n), kTeWarps * 32, 0, (cudaStream_t)stream>>>(X, ldx, Xb, ldxb, V, ldv, Vb, ldvb, n, L,
                                                                              Q, vbs, lvs, out);
  GPP_LAUNCH_CHECK();
  return GPP_OK;
}

extern "C" int gpp_taylor_expansion_bwd(const float* gout, const float* Xb, int64_t ldxb, const float* Vb, int64_t ldvb,
                                        int64_t n, int32_t L, int32_t Q, const float* vbs, const float* lvs, float* gX,
                                        int64_t ldgx, float* gV, int64_t ldgv, float* glvs, gpp_stream_t stream) {
  GPP_REQUIRE(gout && vbs && lvs, "taylor_expansion_bwd: null pointer");
  GPP_REQUIRE(n > 0 && L > 0 && L % 4 == 0 && Q >= 0 && Q % 4 == 0, "taylor_expansion_bwd: bad shape");
  GPP_REQUIRE(!gX || (Xb && aligned16(Xb) && aligned16(gX) && ldxb % 4 == 0 && ldgx % 4 == 0),
              "taylor_expansion_bwd: bad gX / Xb");
  GPP_REQUIRE(!gV || (Vb && aligned16(Vb) && aligned16(gV) && ldvb % 4 == 0 && ldgv % 4 == 0),
              "taylor_expansion_bwd: bad gV / Vb");
  if (gX || gV) {
    taylor_bwd_rows_kernel<<<rows_grid(n), kTeWarps * 32, 0, (cudaStream_t)stream>>>(gout, Xb, ldxb, Vb, ldvb, n, L, Q,
                                                                                     gX, ldgx, gV, ldgv);
    GPP_LAUNCH_CHECK();
  }
  if (glvs) {
    taylor_bwd_lvs_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(gout, n, vbs, lvs, glvs);
    GPP_LAUNCH_CHECK();
  }
  return GPP_OK;
}
