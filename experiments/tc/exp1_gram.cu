// Experiment 1: single-CTA tcgen05 kind::tf32 Gram tile, MN-major operands straight from a row-major matrix.
//   D(128 x 256) = sum_k V[k, 0:128]^T V[k, 0:256]
// Checks the UMMA descriptors / TMA swizzle layout, whether kind::tf32 truncates or rounds fp32 inputs, the
// accuracy of the 3xTF32 split and the behaviour of long fp32 accumulation in TMEM.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define GPP_MBAR_SPIN_LIMIT (1u << 24)
#include "../../gppvae_b200/csrc/tc_common.cuh"

using namespace gpp::tc;

constexpr int TM = 128, TN = 256, BK = 16, STAGES = 4;
constexpr int A_BYTES = TM * BK * 4, B_BYTES = TN * BK * 4, RAW_BYTES = A_BYTES + B_BYTES;  // 8K + 16K = 24K
constexpr int STAGE_BYTES = 2 * RAW_BYTES;                                                   // raw + lo
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;

struct Smem {
  uint64_t full[STAGES], conv[STAGES], empty[STAGES], accum;
  uint32_t tmem_base;
};

template <int MODE>
__global__ void __launch_bounds__(192, 1) gram_tile_kernel(const __grid_constant__ CUtensorMap tmV, int nrows,
                                                           float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Smem* sm = reinterpret_cast<Smem*>(base + STAGES * STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = nrows / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&sm->full[s], 1);
      mbar_init(&sm->conv[s], 128);
      mbar_init(&sm->empty[s], 1);
    }
    mbar_init(&sm->accum, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&sm->tmem_base, 256);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = sm->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmV);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&sm->empty[s], ph ^ 1);
        uint8_t* st = base + s * STAGE_BYTES;
        mbar_arrive_expect_tx(&sm->full[s], RAW_BYTES);
        for (int g = 0; g < TM / 32; ++g) tma_load_2d(st + g * (BK * 128), &tmV, g * 32, kb * BK, &sm->full[s]);
        for (int g = 0; g < TN / 32; ++g)
          tma_load_2d(st + A_BYTES + g * (BK * 128), &tmV, g * 32, kb * BK, &sm->full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(TM, TN, true, true);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(MODE == 0 ? &sm->full[s] : &sm->conv[s], ph);
        tcgen05_fence_after();
        const uint32_t a_hi = smem_u32(base + s * STAGE_BYTES), b_hi = a_hi + A_BYTES;
        const uint32_t a_lo = a_hi + RAW_BYTES, b_lo = a_lo + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk) {
          const uint32_t o = kk * 1024;
          const uint64_t dah = umma_desc(a_hi + o, BK * 128, 512, kLayoutSw128Base32), dbh = umma_desc(b_hi + o, BK * 128, 512, kLayoutSw128Base32);
          umma_tf32(tmem, dah, dbh, idesc, (kb | kk) != 0);
          if (MODE != 0) {
            const uint64_t dal = umma_desc(a_lo + o, BK * 128, 512, kLayoutSw128Base32), dbl = umma_desc(b_lo + o, BK * 128, 512, kLayoutSw128Base32);
            umma_tf32(tmem, dah, dbl, idesc, 1);
            umma_tf32(tmem, dal, dbh, idesc, 1);
          }
        }
        umma_commit(&sm->empty[s]);
      }
      umma_commit(&sm->accum);
    }
  } else {
    const int t = threadIdx.x - 64;  // 0..127
    if (MODE != 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&sm->full[s], ph);
        float4* raw = reinterpret_cast<float4*>(base + s * STAGE_BYTES);
        float4* lo = reinterpret_cast<float4*>(base + s * STAGE_BYTES + RAW_BYTES);
#pragma unroll 4
        for (int i = t; i < RAW_BYTES / 16; i += 128) {
          const float4 v = raw[i];
          float4 h, l;
          if (MODE == 1) {
            split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
            raw[i] = h;
          } else {
            h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = tf32_rn(v.x - h.x);
            h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = tf32_rn(v.y - h.y);
            h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = tf32_rn(v.z - h.z);
            h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = tf32_rn(v.w - h.w);
          }
          lo[i] = l;
        }
        fence_proxy_async_smem();
        mbar_arrive(&sm->conv[s]);
      }
    }
    // epilogue: TMEM -> global (row m = TMEM lane, 256 fp32 columns)
    mbar_wait(&sm->accum, 0);
    tcgen05_fence_after();
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    for (int c = 0; c < TN; c += 32) {
      float v[32];
      tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + c, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) out[row * TN + c + j] = v[j];
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float trunc_tf32(float a) { uint32_t u; memcpy(&u, &a, 4); u &= 0xFFFFE000u; memcpy(&a, &u, 4); return a; }
static float rn_tf32(float a) { uint32_t u; memcpy(&u, &a, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&a, &u, 4); return a; }

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

template <int MODE>
static void run(const char* name, const CUtensorMap& tm, int n, float* d_out, const std::vector<float>& V, int Q,
                bool positive) {
  CK(cudaFuncSetAttribute(gram_tile_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  CK(cudaMemset(d_out, 0, TM * TN * 4));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  gram_tile_kernel<MODE><<<1, 192, SMEM_BYTES>>>(tm, n, d_out);
  cudaEventRecord(e1);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<float> out(TM * TN);
  CK(cudaMemcpy(out.data(), d_out, TM * TN * 4, cudaMemcpyDeviceToHost));
  // references
  double emax = 0, etr = 0, ern = 0, ref_max = 0, bias = 0;
  for (int m = 0; m < TM; m += 7)
    for (int c = 0; c < TN; c += 5) {
      double ex = 0, tr = 0, rn = 0;
      for (int k = 0; k < n; ++k) {
        const float a = V[(size_t)k * Q + m], b = V[(size_t)k * Q + c];
        ex += (double)a * b; tr += (double)trunc_tf32(a) * trunc_tf32(b); rn += (double)rn_tf32(a) * rn_tf32(b);
      }
      const double got = out[m * TN + c];
      ref_max = fmax(ref_max, fabs(ex));
      emax = fmax(emax, fabs(got - ex)); etr = fmax(etr, fabs(got - tr)); ern = fmax(ern, fabs(got - rn));
      bias += (got - ex) / (fabs(ex) + 1e-30);
    }
  const int cnt = ((TM + 6) / 7) * ((TN + 4) / 5);
  printf("%-28s n=%6d %s: %.3f ms  max|err|/max|ref| exact %.3e  vs trunc-inputs %.3e  vs rn-inputs %.3e  mean signed rel %.3e\n",
         name, n, positive ? "pos" : "sgn", ms, emax / ref_max, etr / ref_max, ern / ref_max, bias / cnt);
}

int main(int argc, char** argv) {
  const int Q = 512;
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  float* d_out; CK(cudaMalloc(&d_out, TM * TN * 4));
  for (int pass = 0; pass < 3; ++pass) {
    const int n = pass == 0 ? 256 : (pass == 1 ? 4096 : 65536);
    for (int positive = 0; positive < 2; ++positive) {
      std::vector<float> V((size_t)n * Q);
      srand(123 + pass);
      for (auto& x : V) { float u = rand() / (float)RAND_MAX; x = positive ? 0.5f + 0.5f * u : 2.f * u - 1.f; }
      float* d_V; CK(cudaMalloc(&d_V, V.size() * 4));
      CK(cudaMemcpy(d_V, V.data(), V.size() * 4, cudaMemcpyHostToDevice));
      CUtensorMap tm;
      cuuint64_t dims[2] = {(cuuint64_t)Q, (cuuint64_t)n};
      cuuint64_t strides[1] = {(cuuint64_t)Q * 4};
      cuuint32_t box[2] = {32, BK};
      cuuint32_t es[2] = {1, 1};
      CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_V, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      run<0>("1xTF32 raw fp32", tm, n, d_out, V, Q, positive);
      run<1>("3xTF32 rn-split (hi stored)", tm, n, d_out, V, Q, positive);
      run<2>("3xTF32 trunc-split (raw hi)", tm, n, d_out, V, Q, positive);
      cudaFree(d_V);
    }
  }
  printf("done\n");
  return 0;
}
