// Pieces shared by the two tensor-core implementations of pass 1 (gemm_tc.cu: fp32 operands converted in the kernel;
// gemm_planes.cu: pre-split fp16 planes): the tile / split-K geometry, the partial-tile reduction and the fp16 split.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace gpp {

constexpr int kTileM = 256, kTileN = 256;   // output tile of a CTA pair

// fp16 split of an fp32 operand: hi = the value rounded to 11 significant bits (exactly an fp16 number), lo = the
// remainder rounded to fp16, BOTH scaled by the same power of two per operand, 2^(kF16Top - e) with max|x| < 2^e, so
// that hi.hi, hi.lo and lo.hi share one scale and one accumulator.  kF16Top = 7 puts the largest magnitude below 128
// and keeps the remainder a normal fp16 number for every element above 2^-10 of the maximum (below that the error is
// bounded by 2^-32 of the maximum).
constexpr int kF16Top = 7;

// binary exponent e with 2^(e-1) <= max|x| < 2^e from the bit pattern of max|x| (0 when unknown / all zero)
__host__ __device__ __forceinline__ int exp_of_bits_value(uint32_t b) {
  if (b == 0) return 0;
  const int e = (int)((b >> 23) & 0xFF) - 126;
  return e < -50 ? -50 : (e > 50 ? 50 : e);
}
__device__ __forceinline__ int exp_of_bits(const uint32_t* p) { return p ? exp_of_bits_value(*p) : 0; }

// the 11 leading significant bits of an fp32 number given as its bit pattern, rounded half away from zero (the
// remainder is then zero-mean, so the dropped lo.lo term is no coherent bias)
__device__ __forceinline__ float hi11_round(uint32_t bits) { return __uint_as_float((bits + 0x1000u) & 0xFFFFE000u); }

struct Pass1Params {
  int64_t n;
  int Q, L;
  int tm_count;      // ceil(Q / 256)
  int tiles_g;       // lower-triangular tiles of G: (tm, tn) with tn <= tm
  int tn_c;          // ceil(L / 256)
  int tiles;         // tiles_g + tm_count * tn_c
  int splits;
  int64_t rows_per_split;  // multiple of the kernel's k-block
  float* partial;          // [tile][split][256 * 256]
  float* G; int64_t ldg;   // V^T V  (not touched when tiles_g == 0)
  float* C; int64_t ldc;   // V^T X
  const double* scal_c;    // when set: C *= scal[V0] / scal[VN]
  const double* diag;      // when set: G[i][i] = diag[i], the exactly accumulated column sums of squares of V
  const uint32_t* amax;    // device: [0] bits of max|V|, [1] bits of max|X| (fp16 scales); may be null
  const uint32_t* amax_x;  // planes kernel: bits of max|X| live with X's planes (amax then holds V's alone)
  unsigned int* wave_ctr;  // device, zeroed before the launch: producer-units issued so far (wave alignment); may be null
};

__device__ __forceinline__ void decode_tile(const Pass1Params& p, int tile, int& tm, int& tn, bool& is_c) {
  if (tile < p.tiles_g) {
    is_c = false;
    int t = 0;
    while ((t + 1) * (t + 2) / 2 <= tile) ++t;
    tm = t;
    tn = tile - t * (t + 1) / 2;
  } else {
    is_c = true;
    const int r = tile - p.tiles_g;
    tm = r / p.tn_c;
    tn = r - tm * p.tn_c;
  }
}

template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// ---- host side (gemm_tc.cu) ----
// tile list + split count for `pairs` co-resident CTA pairs; rows_per_split is a multiple of kblock
void pass1_geometry(int64_t n, int Q, int L, bool skip_g, int pairs, int kblock, Pass1Params& p);
// GC <- fixed-order fp64 sum of the partial tiles (+ exact diagonal, + C scale), then the upper triangle of G
int launch_pass1_reduce(const Pass1Params& p, cudaStream_t st);
// 2-D row-major tensor map: elem_bytes 4 (fp32) or 2 (fp16); box = {box_cols, box_rows}; out-of-bounds reads are zero
int make_tensor_map_2d(CUtensorMap* m, const void* ptr, int elem_bytes, int64_t rows, int64_t cols, int64_t ld_elems,
                       int box_cols, int box_rows, CUtensorMapSwizzle swz);

}  // namespace gpp
