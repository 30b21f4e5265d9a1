// How fast can all SMs stream pass-1-shaped TMA boxes ({32 floats, 16 rows}, SWIZZLE_128B_ATOM_32B) out of L2 / HBM
// when nothing consumes them?  Every CTA walks the rows of an n x Q fp32 matrix reading two 128-column blocks (its
// "A" and "B" halves of a 256 x 256 tile), ring of `S` 16 KB stages, one consumer thread that frees a stage as soon as
// it has landed.  Tiles are assigned like tc_pass1_kernel does (lower-triangular tile list, consecutive CTA pairs on
// consecutive tiles), so the L2 sharing pattern is the same.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o exp2_tma_stream exp2_tma_stream.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../gppvae_b200/csrc/tc_common.cuh"
using namespace gpp::tc;

constexpr int BK = 16, S = 10, kStage = 16384;

__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, int64_t n, int tiles_t, int boxes_per_stage) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(base + S * kStage);
  uint64_t* empty = full + S;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int pair = blockIdx.x >> 1, rank = blockIdx.x & 1;
  const int tile = pair % tiles_t;
  int t = 0;
  while ((t + 1) * (t + 2) / 2 <= tile) ++t;
  const int tmi = t, tni = tile - t * (t + 1) / 2;
  const int nst = (int)(n / BK);
  if (threadIdx.x == 0) {
    for (int st = 0; st < nst; ++st) {
      const int s = st % S;
      mbar_wait(&empty[s], ((st / S) & 1) ^ 1);
      mbar_arrive_expect_tx(&full[s], boxes_per_stage * 2048);
      for (int g = 0; g < boxes_per_stage; ++g) {
        const int col = (g < 4 ? tmi : tni) * 256 + rank * 128 + (g & 3) * 32;
        tma_load_2d(base + s * kStage + g * 2048, &tm, col, st * BK, &full[s]);
      }
    }
  } else if (threadIdx.x == 32) {
    for (int st = 0; st < nst; ++st) {
      const int s = st % S;
      mbar_wait(&full[s], (st / S) & 1);
      mbar_arrive(&empty[s]);
    }
  }
}

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 65536;
  const int Q = argc > 2 ? atoi(argv[2]) : 4096;
  float* V;
  cudaMalloc(&V, (size_t)n * Q * 4);
  cudaMemset(V, 0, (size_t)n * Q * 4);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fn);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int T = Q / 256, tiles_t = T * (T + 1) / 2;
  for (int promo = 0; promo < 3; ++promo) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)Q, (cuuint64_t)n}, strides[1] = {(cuuint64_t)Q * 4};
    cuuint32_t box[2] = {32, BK}, es[2] = {1, 1};
    const CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                     : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, V, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 1;
    const int smem = S * kStage + 1024 + 256;
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int boxes = 8; boxes >= 4; boxes -= 4) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      stream_kernel<<<sms, 64, smem>>>(tm, n, tiles_t, boxes);
      cudaEventRecord(e0);
      for (int r = 0; r < 3; ++r) stream_kernel<<<sms, 64, smem>>>(tm, n, tiles_t, boxes);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      ms /= 3;
      const double bytes = (double)sms * (n / BK) * boxes * 2048.0;
      printf("promotion %d, %d boxes/stage: %.3f ms, %.2f TB/s into shared memory (%s)\n", promo, boxes, ms, bytes / ms / 1e9,
             cudaGetErrorString(err));
    }
  }
  return 0;
}
