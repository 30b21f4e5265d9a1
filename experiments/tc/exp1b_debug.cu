// Debug: (1) TMEM st/ld round trip, (2) one K=8 MMA from hand-filled smem, K-major no-swizzle and MN-major SW128.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define GPP_MBAR_SPIN_LIMIT (1u << 22)
#include "../../gppvae_b200/csrc/tc_common.cuh"
using namespace gpp::tc;

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// variant 0: K-major, no swizzle (INTERLEAVE): core matrix = 8 rows x 16 bytes (4 tf32), contiguous 128 B.
//   element (r, k): ((r/8) * SBO) + ((k/4) * LBO) + (r%8)*16 + (k%4)*4   [bytes]
// variant 1: MN-major SW128: element (mn, k): (mn/32)*LBO + (k/8)*SBO + (k%8)*128 + swz(((mn%32)/4), k%8)*16 + (mn%4)*4
__global__ void __launch_bounds__(128, 1) dbg_kernel(int variant, float* out, float* rt) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int M = 128, N = 64;
  const int K = variant == 2 ? 16 : 8;
  float* A = reinterpret_cast<float*>(base);            // 128 x 8
  float* B = reinterpret_cast<float*>(base + 8192);     // 64 x 8
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 64);
  // fill operands: A[m][k] = (m % 7) + 0.25 * k ; B[n][k] = (n % 5) - 0.5 * k  (exact in tf32)
  for (int e = threadIdx.x; e < M * K; e += 128) {
    const int m = e / K, k = e % K;
    const float v = (float)(m % 7) + 0.25f * k;
    uint32_t off;
    if (variant == 0) off = (m / 8) * 256 + (k / 4) * 128 + (m % 8) * 16 + (k % 4) * 4;   // SBO = 256, LBO = 128
    else if (variant == 2) { const int chunk = (k / 4) ^ ((m >> 1) & 3); off = m * 64 + chunk * 16 + (k % 4) * 4; }   // K-major SW64
    else { const int chunk = ((m % 32) / 8) ^ (k % 4); off = (m / 32) * 1024 + k * 128 + chunk * 32 + (m % 8) * 4; }
    *reinterpret_cast<float*>(base + off) = v;
  }
  for (int e = threadIdx.x; e < N * K; e += 128) {
    const int n = e / K, k = e % K;
    const float v = (float)(n % 5) - 0.5f * k;
    uint32_t off;
    if (variant == 0) off = (n / 8) * 256 + (k / 4) * 128 + (n % 8) * 16 + (k % 4) * 4;
    else if (variant == 2) { const int chunk = ((n % 32) / 8) ^ (k % 4); off = (n / 32) * 2048 + k * 128 + chunk * 32 + (n % 8) * 4; }  // MN-major, 16 k-rows per group
    else { const int chunk = ((n % 32) / 8) ^ (k % 4); off = (n / 32) * 1024 + k * 128 + chunk * 32 + (n % 8) * 4; }
    *reinterpret_cast<float*>(base + 8192 + off) = v;
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base;
  // (1) st/ld round trip
  {
    float v[32], r[32];
    for (int j = 0; j < 32; ++j) v[j] = (float)(threadIdx.x * 100 + j);
    tmem_st_32x32(tmem + ((uint32_t)(warp * 32) << 16), v);
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), r);
    for (int j = 0; j < 32; ++j) rt[threadIdx.x * 32 + j] = r[j];
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (threadIdx.x == 0) {
    uint64_t da, db;
    uint32_t idesc;
    if (variant == 0) {
      // no swizzle: layout_type 0
      auto mk = [](uint32_t addr, uint32_t lbo, uint32_t sbo) {
        uint64_t d = 0; d |= (uint64_t)((addr >> 4) & 0x3FFF); d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
        d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32; d |= (uint64_t)1 << 46; return d; };
      da = mk(smem_u32(base), 128, 256); db = mk(smem_u32(base + 8192), 128, 256);
      idesc = umma_idesc_tf32(M, N, false, false);
    } else {
      da = umma_desc(smem_u32(base), 1024, 512, kLayoutSw128Base32); db = umma_desc(smem_u32(base + 8192), 1024, 512, kLayoutSw128Base32);
      idesc = umma_idesc_tf32(M, N, true, true);
    }
    if (variant == 2) {
      idesc = umma_idesc_tf32(M, N, false, true);
      for (int kk = 0; kk < 2; ++kk) {
        da = umma_desc(smem_u32(base) + kk * 32, 16, 512, 4 /*SWIZZLE_64B*/);
        db = umma_desc(smem_u32(base + 8192) + kk * 1024, 2048, 512, kLayoutSw128Base32);
        umma_tf32(tmem, da, db, idesc, kk);
      }
    } else {
      umma_tf32(tmem, da, db, idesc, 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  for (int c = 0; c < N; c += 32) {
    float v[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    for (int j = 0; j < 32; ++j) out[threadIdx.x * N + c + j] = v[j];
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)
int main() {
  float *d_out, *d_rt;
  CK(cudaMalloc(&d_out, 128 * 64 * 4)); CK(cudaMalloc(&d_rt, 128 * 32 * 4));
  CK(cudaFuncSetAttribute(dbg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  for (int variant = 0; variant < 3; ++variant) {
    CK(cudaMemset(d_out, 0xff, 128 * 64 * 4));
    dbg_kernel<<<1, 128, 32768>>>(variant, d_out, d_rt);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    static float out[128 * 64], rt[128 * 32];
    CK(cudaMemcpy(out, d_out, sizeof(out), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rt, d_rt, sizeof(rt), cudaMemcpyDeviceToHost));
    int bad_rt = 0, bad = 0;
    for (int t = 0; t < 128; ++t) for (int j = 0; j < 32; ++j) if (rt[t * 32 + j] != (float)(t * 100 + j)) ++bad_rt;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
      float ref = 0; for (int k = 0; k < (variant == 2 ? 16 : 8); ++k) ref += ((m % 7) + 0.25f * k) * ((n % 5) - 0.5f * k);
      if (out[m * 64 + n] != ref) { if (bad < 5) printf("  variant %d mismatch (%d,%d): got %g ref %g\n", variant, m, n, out[m * 64 + n], ref); ++bad; }
    }
    printf("variant %d: tmem st/ld mismatches %d ; mma mismatches %d / %d ; out[0..3]= %g %g %g %g\n", variant, bad_rt, bad, 128 * 64, out[0], out[1], out[2], out[3]);
  }
  return 0;
}
