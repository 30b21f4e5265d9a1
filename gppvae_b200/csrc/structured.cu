// Structured (never-materialised) Khatri-Rao path -- SURVEY.md section 8(f) row 4.
//
// V[i, j q + k] = xn[d_i, j] wn[w_i, k] (vmod.py:28-35) has only P x p + nviews x q free numbers, and every sum over
// the rows of V factors through the (object, view) slots:
//
//   G = V^T V      G[(j,k),(j',k')] = sum_v wn[v,k] wn[v,k'] S_v[j,j'],   S_v = xn^T diag(cnt[:, v]) xn      (p x p)
//   C = V^T X      C[(j,k), l]      = sum_v wn[v,k] T_v[j,l],             T_v = xn^T Xs_v                      (p x L)
//   V W            (V W)[i, :]      = Y[d_i, w_i L ..],   Y = xn [M_0 | M_1 | ...],  M_v[j,l] = sum_k wn[v,k] W[(j,k),l]
//
// with cnt[o, v] = number of rows in slot (o, v) and Xs_v[o, :] = the sum of the rows of X in that slot.  S and T are
// ONE tall-skinny GEMM  xn^T [cnt (x) xn | Xs]  over the P objects, Y is one (P x p)(p x nviews L) GEMM -- both run on
// the tensor-core kernels of gemm_tc.cu -- so the GP term costs O(P p nviews (p + L)) instead of O(N Q (Q + L)) flops
// (c3: 1.3e11 instead of 1.9e13) and V (16 GB at c3) is never written.  What remains per row of X is streaming:
// the slot sums (read X once) and the epilogue Xb = (X - gather(Y)) / vn (read X, read Y, write Xb).
//
// The slot sums are deterministic: the Python layer sorts the rows by slot once per (d, w) (index preparation, the
// data set does not change between epochs, train_gppvae.py:123-126) and kr_slot_sum_kernel adds each slot's rows in
// that fixed order -- no atomics.
#include "common.cuh"
#include "kernels.h"

namespace gpp {

// XZ[o, zcol0 + v L + l] = sum over the rows i of slot s = o nviews + v (order[slot_start[s] .. slot_start[s+1]))
// of X[i, l];  XZ[o, v p + j] = cnt[s] xn[o, j].  One warp per slot.
__global__ void __launch_bounds__(256) kr_slot_sum_kernel(const float* __restrict__ X, int64_t ldx,
                                                          const int64_t* __restrict__ order,
                                                          const int64_t* __restrict__ slot_start,
                                                          const float* __restrict__ xn, int64_t P, int p, int nviews,
                                                          int L, int with_x, float* __restrict__ XZ, int64_t ldxz) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int zcol0 = with_x ? nviews * p : 0;
  for (int64_t s = warp; s < P * nviews; s += nwarps) {
    const int64_t o = s / nviews;
    const int v = (int)(s - o * nviews);
    const int64_t b = slot_start[s], e = slot_start[s + 1];
    float* zrow = XZ + o * ldxz + zcol0 + (int64_t)v * L;
    for (int c = lane * 4; c < L; c += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int64_t t = b; t < e; ++t) {
        const float4 x = *reinterpret_cast<const float4*>(X + order[t] * ldx + c);
        acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
      }
      *reinterpret_cast<float4*>(zrow + c) = acc;
    }
    if (!with_x) continue;
    const float cnt = (float)(e - b);
    float* xrow = XZ + o * ldxz + (int64_t)v * p;
    for (int c = lane * 4; c < p; c += 128) {
      const float4 x = *reinterpret_cast<const float4*>(xn + o * p + c);
      *reinterpret_cast<float4*>(xrow + c) = make_float4(cnt * x.x, cnt * x.y, cnt * x.z, cnt * x.w);
    }
  }
}

// GC (Q x (Q + L), ld = ldgc) from ST (p x nviews (p + L), ld = ldst): see the header of this file.
// One thread per float4 of a GC row; the view weights of the row's k sit in registers.
__global__ void __launch_bounds__(256) kr_assemble_gc_kernel(const float* __restrict__ ST, int64_t ldst,
                                                             const float* __restrict__ wn, int p, int q, int nviews,
                                                             int L, int with_g, float* __restrict__ GC, int64_t ldgc) {
  extern __shared__ float wsm[];   // wn, nviews x q
  for (int e = threadIdx.x; e < nviews * q; e += blockDim.x) wsm[e] = wn[e];
  __syncthreads();
  const int Q = with_g ? p * q : 0;       // columns of the G part in GC (0: GC is C alone, ST holds only T)
  const int r = blockIdx.x;               // GC row (j, k)
  const int j = r / q, k = r - j * q;
  const int zcol0 = with_g ? nviews * p : 0;
  const int ncol4 = (Q + L) >> 2;
  for (int c4 = threadIdx.x; c4 < ncol4; c4 += blockDim.x) {
    const int c = c4 << 2;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < Q) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int cc = c + e, j2 = cc / q, k2 = cc - j2 * q;
        float s = 0.f;
        for (int v = 0; v < nviews; ++v) s = fmaf(wsm[v * q + k] * wsm[v * q + k2], ST[(int64_t)j * ldst + v * p + j2], s);
        o[e] = s;
      }
    } else {
      const int l = c - Q;
      for (int v = 0; v < nviews; ++v) {
        const float4 t = *reinterpret_cast<const float4*>(ST + (int64_t)j * ldst + zcol0 + (int64_t)v * L + l);
        const float wk = wsm[v * q + k];
        o[0] = fmaf(wk, t.x, o[0]); o[1] = fmaf(wk, t.y, o[1]); o[2] = fmaf(wk, t.z, o[2]); o[3] = fmaf(wk, t.w, o[3]);
      }
    }
    *reinterpret_cast<float4*>(GC + (int64_t)r * ldgc + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// M (p x nviews L, ld = ldm):  M[j, v L + l] = sum_k wn[v, k] W[(j, k), l]
__global__ void __launch_bounds__(256) kr_assemble_m_kernel(const float* __restrict__ W, int64_t ldw,
                                                            const float* __restrict__ wn, int p, int q, int nviews,
                                                            int L, float* __restrict__ M, int64_t ldm) {
  const int j = blockIdx.x;
  const int n4 = (nviews * L) >> 2;
  for (int c4 = threadIdx.x; c4 < n4; c4 += blockDim.x) {
    const int c = c4 << 2, v = c / L, l = c - v * L;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < q; ++k) {
      const float wk = wn[v * q + k];
      const float4 x = *reinterpret_cast<const float4*>(W + (int64_t)(j * q + k) * ldw + l);
      acc.x = fmaf(wk, x.x, acc.x); acc.y = fmaf(wk, x.y, acc.y); acc.z = fmaf(wk, x.z, acc.z); acc.w = fmaf(wk, x.w, acc.w);
    }
    *reinterpret_cast<float4*>(M + (int64_t)j * ldm + c) = acc;
  }
}

// Epilogue of the structured pass 2: Xb_i = (X_i - Y[d_i, w_i L ..]) / vn, quad_i = X_i . Xb_i, per-block sum Xb^2.
// One warp per row.
__global__ void __launch_bounds__(256) kr_xb_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ Y,
                                                    int64_t ldy, const int64_t* __restrict__ d,
                                                    const int64_t* __restrict__ w, int64_t n, int64_t P, int nviews, int L,
                                                    const double* __restrict__ scal, float* __restrict__ Xb, int64_t ldxb,
                                                    float* __restrict__ quad, double* __restrict__ xb2_part) {
  __shared__ double red[8];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * 8 + wib, nwarps = (int64_t)gridDim.x * 8;
  const float inv_vn = (float)(1.0 / scal[GPP_S_VN]);
  const float qnan = __int_as_float(0x7fc00000);
  double xb2 = 0;
  for (int64_t i = warp; i < n; i += nwarps) {
    const int64_t di = d[i], wi = w[i];
    const bool ok = (di >= 0) & (di < P) & (wi >= 0) & (wi < nviews);
    const float* y = Y + (ok ? di * ldy + wi * L : 0);
    float qs = 0.f, x2 = 0.f;
    for (int c = lane * 4; c < L; c += 128) {
      const float4 x = *reinterpret_cast<const float4*>(X + i * ldx + c);
      const float4 yy = *reinterpret_cast<const float4*>(y + c);
      float4 o;
      o.x = ok ? (x.x - yy.x) * inv_vn : qnan; o.y = ok ? (x.y - yy.y) * inv_vn : qnan;
      o.z = ok ? (x.z - yy.z) * inv_vn : qnan; o.w = ok ? (x.w - yy.w) * inv_vn : qnan;
      *reinterpret_cast<float4*>(Xb + i * ldxb + c) = o;
      qs = fmaf(x.x, o.x, qs); qs = fmaf(x.y, o.y, qs); qs = fmaf(x.z, o.z, qs); qs = fmaf(x.w, o.w, qs);
      x2 = fmaf(o.x, o.x, x2); x2 = fmaf(o.y, o.y, x2); x2 = fmaf(o.z, o.z, x2); x2 = fmaf(o.w, o.w, x2);
    }
    qs = warp_sum(qs);
    x2 = warp_sum(x2);
    if (lane == 0) quad[i] = qs;
    xb2 += (double)x2;
  }
  if (lane == 0) red[wib] = xb2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int k = 0; k < 8; ++k) t += red[k];
    xb2_part[blockIdx.x] = t;
  }
}

constexpr int kKrXbBlocks = 1184;   // 8 x 148

}  // namespace gpp

using namespace gpp;

extern "C" int gpp_kr_slot_sums(const float* X, int64_t ldx, const int64_t* order, const int64_t* slot_start,
                                const float* xn, int64_t P, int32_t p, int32_t nviews, int32_t L, int32_t with_x,
                                float* XZ, int64_t ldxz, gpp_stream_t stream) {
  GPP_REQUIRE(X && order && slot_start && xn && XZ, "kr_slot_sums: null pointer");
  GPP_REQUIRE(P > 0 && p > 0 && nviews > 0 && L > 0 && p % 4 == 0 && L % 4 == 0, "kr_slot_sums: p and L must be multiples of 4");
  GPP_REQUIRE(ldx >= L && ldx % 4 == 0 && ldxz >= (int64_t)nviews * ((with_x ? p : 0) + L) && ldxz % 4 == 0 && aligned16(X) &&
                  aligned16(XZ) && aligned16(xn),
              "kr_slot_sums: bad leading dimension / alignment");
  const int64_t slots = P * nviews;
  const int grid = (int)(ceil_div(slots, 8) < 8192 ? ceil_div(slots, 8) : 8192);
  kr_slot_sum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, ldx, order, slot_start, xn, P, p, nviews, L, with_x, XZ,
                                                             ldxz);
  GPP_LAUNCH_CHECK();
  return GPP_OK;
}

extern "C" int gpp_kr_assemble_gc(const float* ST, int64_t ldst, const float* wn, int32_t p, int32_t q, int32_t nviews,
                                  int32_t L, int32_t with_g, float* GC, int64_t ldgc, gpp_stream_t stream) {
  GPP_REQUIRE(ST && wn && GC, "kr_assemble_gc: null pointer");
  const int64_t Q = (int64_t)p * q;
  GPP_REQUIRE(p > 0 && q > 0 && nviews > 0 && L >= 0 && Q % 4 == 0 && L % 4 == 0 && p % 4 == 0,
              "kr_assemble_gc: p, p*q and L must be multiples of 4");
  GPP_REQUIRE(ldst >= (int64_t)nviews * ((with_g ? p : 0) + L) && ldst % 4 == 0 && ldgc >= (with_g ? Q : 0) + L &&
                  ldgc % 4 == 0 && aligned16(ST) &&
                  aligned16(GC),
              "kr_assemble_gc: bad leading dimension / alignment");
  const size_t smem = (size_t)nviews * q * sizeof(float);
  GPP_REQUIRE(smem <= 48 * 1024, "kr_assemble_gc: view table too large");
  kr_assemble_gc_kernel<<<(unsigned)Q, 256, smem, (cudaStream_t)stream>>>(ST, ldst, wn, p, q, nviews, L, with_g, GC, ldgc);
  GPP_LAUNCH_CHECK();
  return GPP_OK;
}

extern "C" int gpp_kr_assemble_m(const float* W, int64_t ldw, const float* wn, int32_t p, int32_t q, int32_t nviews,
                                 int32_t L, float* M, int64_t ldm, gpp_stream_t stream) {
  GPP_REQUIRE(W && wn && M, "kr_assemble_m: null pointer");
  GPP_REQUIRE(p > 0 && q > 0 && nviews > 0 && L > 0 && L % 4 == 0, "kr_assemble_m: L must be a multiple of 4");
  GPP_REQUIRE(ldw >= L && ldw % 4 == 0 && ldm >= (int64_t)nviews * L && ldm % 4 == 0 && aligned16(W) && aligned16(M),
              "kr_assemble_m: bad leading dimension / alignment");
  kr_assemble_m_kernel<<<(unsigned)p, 256, 0, (cudaStream_t)stream>>>(W, ldw, wn, p, q, nviews, L, M, ldm);
  GPP_LAUNCH_CHECK();
  return GPP_OK;
}

extern "C" size_t gpp_kr_xb_workspace_bytes(int64_t n) {
  return align_up((size_t)(n > 0 ? n : 1) * sizeof(float), 256) + align_up((size_t)kKrXbBlocks * sizeof(double), 256) +
         xb_finalize_bytes();
}

extern "C" int gpp_kr_xb_nll(const float* X, int64_t ldx, const float* Y, int64_t ldy, const int64_t* d,
                             const int64_t* w, int64_t n, int64_t P, int32_t nviews, int32_t L, double* scal, float* Xb,
                             int64_t ldxb, float* nll, void* workspace, size_t workspace_bytes, gpp_stream_t stream) {
  GPP_REQUIRE(X && Y && d && w && scal && Xb && nll, "kr_xb_nll: null pointer");
  GPP_REQUIRE(n >= 0 && P > 0 && nviews > 0 && L > 0 && L % 4 == 0, "kr_xb_nll: bad shape");
  GPP_REQUIRE(ldx >= L && ldx % 4 == 0 && ldxb >= L && ldxb % 4 == 0 && ldy >= (int64_t)nviews * L && ldy % 4 == 0 &&
                  aligned16(X) && aligned16(Y) && aligned16(Xb),
              "kr_xb_nll: bad leading dimension / alignment");
  const size_t need = gpp_kr_xb_workspace_bytes(n);
  if (!workspace || workspace_bytes < need) {
    set_error("kr_xb_nll: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return GPP_ERR_WORKSPACE;
  }
  char* ws = static_cast<char*>(workspace);
  float* quad = reinterpret_cast<float*>(ws);
  double* part = reinterpret_cast<double*>(ws + align_up((size_t)(n > 0 ? n : 1) * sizeof(float), 256));
  double* fin = part + align_up((size_t)kKrXbBlocks * sizeof(double), 256) / sizeof(double);
  cudaStream_t st = (cudaStream_t)stream;
  kr_xb_kernel<<<kKrXbBlocks, 256, 0, st>>>(X, ldx, Y, ldy, d, w, n, P, nviews, L, scal, Xb, ldxb, quad, part);
  GPP_LAUNCH_CHECK();
  return launch_xb_finalize(quad, 1, n, part, kKrXbBlocks, fin, scal, nll, st);
}
