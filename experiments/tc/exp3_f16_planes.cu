// Round-2 design experiment for pass 1 on pre-split fp16 planes (V = hi + lo, each an [n][Q] fp16 matrix).
//
// Question: what bounds a 256 x 256 (per CTA pair) Gram tile when both MMA operands come straight from TMA-written
// shared memory -- the L2 -> shared-memory stream, the tensor pipe, or shared-memory bandwidth -- and does TMA
// multicast inside a cluster of 2 / 4 pairs move the ceiling?
//
//   mode 0  stream only: TMA boxes {64 halfs, BR rows} (SWIZZLE_128B), consumer thread frees a stage when it landed
//   mode 1  stream + MMA: the pair leader issues, per 16 k-rows, the three kind::f16 M=256 N=256 K=16 SS MMAs
//           (Ah.Bl, Al.Bh, Ah.Bh) into alternating accumulators (no drain: measures the operand pipeline alone)
//   mode 2  MMA only (no loads): the tensor pipe + shared-memory operand fetch ceiling of that instruction mix
//   mode 3  MMA only, kind::tf32 K=8 on the same shared memory (dense tf32 rate of the SS form)
// Cluster size CS: 2 = one pair (unicast), 4 = two pairs sharing the M slab (A multicast), 8 = 2 x 2 pairs (A and B
// multicast).  Each pair's loads signal the full barrier of the pair LEADER (cta_group::2 form of the TMA load).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o exp3_f16_planes exp3_f16_planes.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../gppvae_b200/csrc/tc_common.cuh"
using namespace gpp::tc;

constexpr int kRing = 196608;   // bytes of operand ring per CTA

__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* m, int x, int y, uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_mc(uint32_t dst, const CUtensorMap* m, int x, int y, uint32_t bar_cluster,
                                                    uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, "
      "{%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(x), "r"(y), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct Bars {
  uint64_t full[16], empty[16], done;
  uint32_t tmem_base;
};

// BR: rows per box / stage.  Stage layout per CTA: [Ah g0][Ah g1][Al g0][Al g1][Bh g0][Bh g1][Bl g0][Bl g1], BR * 128 B each.
template <int CS>
__global__ void __launch_bounds__(128, 1)
planes_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmL, int64_t n, int BR,
              int slabs, int mode, int stages) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  Bars* bars = reinterpret_cast<Bars*>(base + kRing);
  const uint32_t crank = cluster_ctarank();          // rank in the cluster
  const uint32_t r = crank & 1, pi = crank >> 1;     // rank in the pair, pair in the cluster
  const int cluster_id = blockIdx.x / CS;
  const int box_bytes = BR * 128, stage_bytes = 8 * box_bytes;
  const int S = stages;
  constexpr int kPairs = CS / 2;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], kPairs);
    }
    mbar_init(&bars->done, 1);
    fence_mbar_init();
  }
  if ((threadIdx.x >> 5) == 1) tmem_alloc_pair(&bars->tmem_base, 512);
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const int nst = (int)(n / BR);

  // slab assignment (synthetic, wraps around)
  int a_slab, b_slab;
  if (CS == 2) { a_slab = (cluster_id * 3) % slabs; b_slab = (cluster_id * 5 + 1) % slabs; }
  else if (CS == 4) { a_slab = cluster_id % slabs; b_slab = (2 * cluster_id + (int)pi) % slabs; }
  else { a_slab = (2 * cluster_id + (int)(pi >> 1)) % slabs; b_slab = (2 * cluster_id + 5 + (int)(pi & 1)) % slabs; }
  const int acol = a_slab * 256 + (int)r * 128, bcol = b_slab * 256 + (int)r * 128;

  if (threadIdx.x == 0 && mode != 2 && mode != 3) {
    // ------------------------------------------------ producer
    // barrier of the pair leader: the shared::cta address with the peer bit cleared (what CUTLASS' 2-SM TMA loads do)
    const uint32_t full0_leader = smem_u32(&bars->full[0]) & 0xFEFFFFFFu;
    for (int st = 0; st < nst; ++st) {
      const int s = st % S;
      mbar_wait(&bars->empty[s], ((st / S) & 1) ^ 1);
      const uint32_t bar = full0_leader + 8u * s;
      if (r == 0) mbar_arrive_expect_tx(&bars->full[s], 2 * stage_bytes);
      const uint32_t dst = smem_u32(base + s * stage_bytes);
      const int row = st * BR;
      for (int b = 0; b < 8; ++b) {
        const bool is_a = b < 4;
        const CUtensorMap* m = (b & 2) ? &tmL : &tmH;
        const int col = (is_a ? acol : bcol) + (b & 1) * 64;
        if (CS == 2) {
          tma_load_2d_pair(dst + b * box_bytes, m, col, row, bar);
        } else if (CS == 4) {
          if (is_a) {
            if ((uint32_t)((b & 3) >> 1) == pi)   // pair 0 issues the hi boxes, pair 1 the lo boxes
              tma_load_2d_pair_mc(dst + b * box_bytes, m, col, row, bar, (uint16_t)((1u << r) | (1u << (r + 2))));
          } else {
            tma_load_2d_pair(dst + b * box_bytes, m, col, row, bar);
          }
        } else {
          const uint32_t mrow = pi >> 1, ncol = pi & 1;
          if (is_a) {   // shared by the pairs (mrow, 0) and (mrow, 1)
            if ((uint32_t)((b & 3) >> 1) == ncol)
              tma_load_2d_pair_mc(dst + b * box_bytes, m, col, row, bar,
                                  (uint16_t)((1u << ((mrow * 2 + 0) * 2 + r)) | (1u << ((mrow * 2 + 1) * 2 + r))));
          } else {      // shared by the pairs (0, ncol) and (1, ncol)
            if ((uint32_t)((b & 3) >> 1) == mrow)
              tma_load_2d_pair_mc(dst + b * box_bytes, m, col, row, bar,
                                  (uint16_t)((1u << ((0 * 2 + ncol) * 2 + r)) | (1u << ((1 * 2 + ncol) * 2 + r))));
          }
        }
      }
    }
  } else if (threadIdx.x == 32 && r == 0) {
    // ------------------------------------------------ consumer (pair leader)
    uint32_t empty_addr[CS];
    for (int c = 0; c < CS; ++c) empty_addr[c] = mapa_u32(&bars->empty[0], c);
    if (mode == 0) {
      for (int st = 0; st < nst; ++st) {
        const int s = st % S;
        mbar_wait(&bars->full[s], (st / S) & 1);
        for (int c = 0; c < CS; ++c) mbar_arrive_cluster(empty_addr[c] + 8u * s);
      }
    } else {
      const bool tf32 = mode == 3;
      const uint32_t idesc16 = umma_idesc_f16(256, 256, true, true);
      const uint32_t idesc32 = umma_idesc_tf32(256, 256, false, false);
      const uint32_t lbo = (uint32_t)box_bytes;   // between 64-column groups
      uint16_t mask = 0;
      for (int c = 0; c < CS; ++c) mask |= (uint16_t)(1u << c);
      for (int st = 0; st < nst; ++st) {
        const int s = st % S;
        if (mode == 1) {
          mbar_wait(&bars->full[s], (st / S) & 1);
          tcgen05_fence_after();
        }
        const uint32_t sb = smem_u32(base + s * stage_bytes);
        const uint32_t ah = sb, al = sb + 2 * box_bytes, bh = sb + 4 * box_bytes, bl = sb + 6 * box_bytes;
        const uint32_t d = tmem + ((st & 1) ? 256u : 0u);
        for (int kk = 0; kk < BR / 16; ++kk) {
          const uint32_t o = kk * 2048;
          if (!tf32) {
            umma_f16_pair_ss(d, umma_desc(ah + o, lbo, 1024, kLayoutSw128), umma_desc(bl + o, lbo, 1024, kLayoutSw128), idesc16, 1);
            umma_f16_pair_ss(d, umma_desc(al + o, lbo, 1024, kLayoutSw128), umma_desc(bh + o, lbo, 1024, kLayoutSw128), idesc16, 1);
            umma_f16_pair_ss(d, umma_desc(ah + o, lbo, 1024, kLayoutSw128), umma_desc(bh + o, lbo, 1024, kLayoutSw128), idesc16, 1);
          } else {
            // K-major SWIZZLE_128B operands (128 rows x 32 tf32 per 16 KB): three K = 8 steps, 32 B apart
            for (int j = 0; j < 3; ++j) {
              const uint32_t o8 = ((kk * 3 + j) & 3) * 32;
              umma_tf32_pair(d, umma_desc(sb + o8, 16, 1024, kLayoutSw128), umma_desc(sb + 16384 + o8, 16, 1024, kLayoutSw128), idesc32, 1);
            }
          }
        }
        if (mode == 1) umma_commit_pair(&bars->empty[s], mask);
      }
      umma_commit_pair(&bars->done, (uint16_t)(1u << crank));
      mbar_wait(&bars->done, 0);
    }
  }
  __syncthreads();
  tcgen05_fence_before();
  cluster_sync_all();
  if ((threadIdx.x >> 5) == 1) tmem_dealloc_pair(tmem, 512);
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CS>
float run(const CUtensorMap& tmH, const CUtensorMap& tmL, int64_t n, int BR, int slabs, int mode, int sms, int reps) {
  const int smem = kRing + 1024 + (int)sizeof(Bars);
  cudaFuncSetAttribute(planes_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(planes_kernel<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int stages = kRing / (8 * BR * 128) > 16 ? 16 : kRing / (8 * BR * 128);
  cudaLaunchConfig_t cfg{};
  const int grid = (sms / CS) * CS;
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at{};
  at.id = cudaLaunchAttributeClusterDimension;
  at.val.clusterDim.x = CS; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
  cfg.attrs = &at; cfg.numAttrs = 1;
  int maxc = 0;
  cudaOccupancyMaxActiveClusters(&maxc, planes_kernel<CS>, &cfg);
  if (maxc * CS < grid) cfg.gridDim = dim3(maxc * CS);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaLaunchKernelEx(&cfg, planes_kernel<CS>, tmH, tmL, n, BR, slabs, mode, stages);
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) cudaLaunchKernelEx(&cfg, planes_kernel<CS>, tmH, tmL, n, BR, slabs, mode, stages);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= reps;
  const double ctas = cfg.gridDim.x;
  const double bytes = ctas * (double)(n / BR) * 8.0 * BR * 128.0;
  const double flops = (ctas / 2) * (double)(n / 16) * 3.0 * 2.0 * 256 * 256 * (mode == 3 ? 8 : 16);
  printf("CS=%d BR=%2d mode=%d ctas=%3d stages=%2d: %8.3f ms", CS, BR, mode, (int)ctas, stages, ms);
  if (mode <= 1) printf("  %6.2f TB/s into smem", bytes / ms / 1e9);
  if (mode >= 1) printf("  %7.1f TFLOP/s (%s)", flops / ms / 1e9, mode == 3 ? "tf32" : "fp16");
  printf("  [%s]\n", cudaGetErrorString(err));
  fflush(stdout);
  return ms;
}

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 131072;
  const int Q = argc > 2 ? atoi(argv[2]) : 4096;
  __half *H, *L;
  cudaMalloc(&H, (size_t)n * Q * 2);
  cudaMalloc(&L, (size_t)n * Q * 2);
  cudaMemset(H, 0, (size_t)n * Q * 2);
  cudaMemset(L, 0, (size_t)n * Q * 2);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncFn enc = reinterpret_cast<EncFn>(fn);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int slabs = Q / 256;
  for (int BR : {16, 32, 64}) {
    CUtensorMap tmH, tmL;
    cuuint64_t dims[2] = {(cuuint64_t)Q, (cuuint64_t)n}, strides[1] = {(cuuint64_t)Q * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)BR}, es[2] = {1, 1};
    if (enc(&tmH, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, H, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return 1;
    if (enc(&tmL, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, L, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return 1;
    for (int mode : {0, 1}) {
      run<2>(tmH, tmL, n, BR, slabs, mode, sms, 3);
      run<4>(tmH, tmL, n, BR, slabs, mode, sms, 3);
      run<8>(tmH, tmL, n, BR, slabs, mode, sms, 3);
    }
    if (BR == 32) {
      run<2>(tmH, tmL, n, BR, slabs, 2, sms, 3);
      run<2>(tmH, tmL, n, BR, slabs, 3, sms, 3);
    }
  }
  return 0;
}
