// fp32 SIMT tile engine shared by the Q-space kernels (Cholesky trailing updates, triangular
// inverse, small GEMMs) and -- until the tcgen05 path replaces them for the two N-long sweeps --
// by pass 1 / pass 2.  Classic 128x128x16 register-tiled SGEMM: 256 threads, 8x8 outputs per
// thread as a 2x2 arrangement of 4x4 blocks, double-buffered shared memory with register
// prefetch, 128-bit global and shared accesses.
//
// Operand addressing (elements): an operand is "row-contracted" when the contraction index walks
// the rows of a row-major matrix (element(mn, k) = base[k * ld + mn], contiguous along the output
// index -- V in V^T V) and "column-contracted" when it walks the columns
// (element(mn, k) = base[mn * ld + k] -- V in V W).
#pragma once
#include "common.cuh"

namespace gpp {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int kGemmThreads = 256;
constexpr int kLdsPad = 4;               // keeps rows 16-byte aligned, breaks the worst store conflicts
constexpr int kLds = BM + kLdsPad;       // BM == BN

struct Operand {
  const float* base;   // points at element (mn = 0, k = 0) of this tile
  int64_t ld;          // leading dimension in elements
  int mn_valid;        // rows/cols of the tile that exist (rest reads as zero); multiple of 4 if row-contracted
};

__device__ __forceinline__ float4 ldg4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// Fetch this thread's two float4 of a (BK x 128) operand slab starting at contraction index k0.
// ROWC (row-contracted): slab rows are k, 32 float4 per row.  k_valid = number of k that exist from k0.
template <bool ROWC>
__device__ __forceinline__ void fetch_slab(const Operand& op, int64_t k0, int k_valid, float4 (&r)[2]) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = tid + i * kGemmThreads;  // 0..511
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ROWC) {
      const int k = idx >> 5, m4 = (idx & 31) << 2;
      if (k < k_valid && m4 < op.mn_valid) v = ldg4(op.base + (k0 + k) * op.ld + m4);
    } else {
      const int m = idx >> 2, k4 = (idx & 3) << 2;
      if (m < op.mn_valid && k4 < k_valid) v = ldg4(op.base + (int64_t)m * op.ld + k0 + k4);
    }
    r[i] = v;
  }
}

template <bool ROWC>
__device__ __forceinline__ void stash_slab(float (*s)[kLds], const float4 (&r)[2]) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = tid + i * kGemmThreads;
    if (ROWC) {
      const int k = idx >> 5, m4 = (idx & 31) << 2;
      *reinterpret_cast<float4*>(&s[k][m4]) = r[i];
    } else {
      const int m = idx >> 2, k4 = (idx & 3) << 2;
      s[k4 + 0][m] = r[i].x;
      s[k4 + 1][m] = r[i].y;
      s[k4 + 2][m] = r[i].z;
      s[k4 + 3][m] = r[i].w;
    }
  }
}

struct TileSmem {
  float a[2][BK][kLds];
  float b[2][BK][kLds];
};

// acc[i][j] += sum_{k in [0, klen)} A(row_i, k) * B(col_j, k), where
//   row_i = (i < 4 ? ty*4 + i : 64 + ty*4 + i - 4),  col_j likewise with tx;  tx = tid & 15, ty = tid >> 4.
// When the contraction runs over columns (not ROWC) klen must be a multiple of 4.
template <bool A_ROWC, bool B_ROWC>
__device__ __forceinline__ void tile_mainloop(const Operand& A, const Operand& B, int64_t klen, TileSmem& sm,
                                              float (&acc)[8][8]) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float4 ra[2], rb[2];
  const int64_t ntiles = (klen + BK - 1) / BK;
  if (ntiles == 0) return;
  {
    const int kv = (int)(klen < BK ? klen : BK);
    fetch_slab<A_ROWC>(A, 0, kv, ra);
    fetch_slab<B_ROWC>(B, 0, kv, rb);
  }
  stash_slab<A_ROWC>(sm.a[0], ra);
  stash_slab<B_ROWC>(sm.b[0], rb);
  __syncthreads();
  for (int64_t t = 0; t < ntiles; ++t) {
    const int cur = (int)(t & 1);
    if (t + 1 < ntiles) {
      const int64_t k0 = (t + 1) * BK;
      const int64_t left = klen - k0;
      const int kv = (int)(left < BK ? left : BK);
      fetch_slab<A_ROWC>(A, k0, kv, ra);
      fetch_slab<B_ROWC>(B, k0, kv, rb);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sm.a[cur][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sm.a[cur][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&sm.b[cur][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&sm.b[cur][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (t + 1 < ntiles) {
      stash_slab<A_ROWC>(sm.a[cur ^ 1], ra);
      stash_slab<B_ROWC>(sm.b[cur ^ 1], rb);
    }
    __syncthreads();
  }
}

// tile-local row / column of accumulator slot i
__device__ __forceinline__ int acc_row(int i) { return (i < 4 ? 0 : 60) + (threadIdx.x >> 4) * 4 + i; }
__device__ __forceinline__ int acc_col(int j) { return (j < 4 ? 0 : 60) + (threadIdx.x & 15) * 4 + j; }

}  // namespace gpp
