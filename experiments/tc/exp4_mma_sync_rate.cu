// Legacy warp-level mma.sync rates on sm_100a (dense TF32 m16n8k8, FP16 m16n8k16 with fp32 accumulation), per SM and
// clock, for 4 / 8 / 16 / 32 resident warps per SM: decides whether a 3xTF32 mma.sync tile is worth having for the
// Cholesky's trailing updates (fp32 SIMT peak: 128 FMA / clk / SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp4_mma_sync_rate exp4_mma_sync_rate.cu && ./exp4_mma_sync_rate
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <int KIND>
__global__ void __launch_bounds__(1024) rate_kernel(int iters, float* out, long long* cycles) {
  float c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 5, a3 = 7, b0 = 11, b1 = 13;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456f) out[0] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(1024) ffma_kernel(int iters, float* out, long long* cycles) {
  float c[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) c[i] = (float)i;
  const float a = 1.0001f + threadIdx.x * 1e-9f, b = 0.5f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) c[i] = fmaf(c[i], a, b);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += c[i];
  if (s == 123.456f) out[0] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}


// ---- the Cholesky's 3xTF32 step, as in qspace.cu (mma_step), on operands in shared memory: LDS + split + 24 MMAs in three
// rounds of independent ones + the delayed adds.  VAR 0: everything; 1: no split (lo = hi); 2: MMAs only (operands loaded
// once outside the loop)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
constexpr int PITCH = 36;   // (static shared memory: 48 KB; the kernel uses pitch 36 for global tiles, 68 for the panel)
template <int VAR>
__global__ void __launch_bounds__(256, 1) step_kernel(int iters, float* out, long long* cycles) {
  __shared__ float As[128][PITCH];
  __shared__ float Bs[64][PITCH];
  for (int e = threadIdx.x; e < 128 * PITCH; e += blockDim.x) (&As[0][0])[e] = 1.f + 1e-3f * (e % 97);
  for (int e = threadIdx.x; e < 64 * PITCH; e += blockDim.x) (&Bs[0][0])[e] = 0.5f + 1e-3f * (e % 89);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wm = warp >> 1, wn = warp & 1, g = lane >> 2, t = lane & 3;
  const float* ap = &As[wm * 32][0] + g * PITCH + t;
  const float* bp = &Bs[wn * 32][0] + g * PITCH + t;
  float h[2][4][4], l[2][4][4], p[2][4][4], c[2][4][4];
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) for (int e = 0; e < 4; ++e) h[i][j][e] = l[i][j][e] = p[i][j][e] = c[i][j][e] = 0.f;
  uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
  auto load = [&](int kk) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if (VAR == 0) { split_tf32(bp[nt * 8 * PITCH + kk], bh[nt][0], bl[nt][0]); split_tf32(bp[nt * 8 * PITCH + kk + 4], bh[nt][1], bl[nt][1]); }
      else { bh[nt][0] = bl[nt][0] = __float_as_uint(bp[nt * 8 * PITCH + kk]); bh[nt][1] = bl[nt][1] = __float_as_uint(bp[nt * 8 * PITCH + kk + 4]); }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      if (VAR == 0) {
        split_tf32(ap[(mt * 16) * PITCH + kk], ah[mt][0], al[mt][0]); split_tf32(ap[(mt * 16 + 8) * PITCH + kk], ah[mt][1], al[mt][1]);
        split_tf32(ap[(mt * 16) * PITCH + kk + 4], ah[mt][2], al[mt][2]); split_tf32(ap[(mt * 16 + 8) * PITCH + kk + 4], ah[mt][3], al[mt][3]);
      } else {
        ah[mt][0] = al[mt][0] = __float_as_uint(ap[(mt * 16) * PITCH + kk]); ah[mt][1] = al[mt][1] = __float_as_uint(ap[(mt * 16 + 8) * PITCH + kk]);
        ah[mt][2] = al[mt][2] = __float_as_uint(ap[(mt * 16) * PITCH + kk + 4]); ah[mt][3] = al[mt][3] = __float_as_uint(ap[(mt * 16 + 8) * PITCH + kk + 4]);
      }
    }
  };
  auto mmas = [&](float (&cc)[2][4][4]) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) cc[mt][nt][e] = 0.f;
        mma_tf32(cc[mt][nt], ah[mt], bh[nt]);
      }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma_tf32(l[mt][nt], al[mt], bh[nt]);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma_tf32(l[mt][nt], ah[mt], bl[nt]);
  };
  auto add = [&](float (&cc)[2][4][4]) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) h[mt][nt][e] += cc[mt][nt][e];
  };
  if (VAR == 2) load(0);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int ks = 0; ks < 4; ks += 2) {
      if (VAR != 2) load(ks * 8);
      mmas(c);
      add(p);
      if (VAR != 2) load(ks * 8 + 8);
      mmas(p);
      add(c);
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) for (int e = 0; e < 4; ++e) s += h[i][j][e] + l[i][j][e] + p[i][j][e];
  if (s == 123.456f) out[0] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  long long* cyc;
  cudaMalloc(&out, 4);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    for (int kind = 0; kind < 3; ++kind) {
      for (int rep = 0; rep < 2; ++rep) {
        if (kind == 0) rate_kernel<0><<<sms, warps * 32>>>(iters, out, cyc);
        else if (kind == 1) rate_kernel<1><<<sms, warps * 32>>>(iters, out, cyc);
        else ffma_kernel<<<sms, warps * 32>>>(iters, out, cyc);
        cudaDeviceSynchronize();
      }
      long long h[256];
      cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < sms; ++i) avg += (double)h[i];
      avg /= sms;
      const double per_mma = kind == 0 ? 16.0 * 8 * 8 : 16.0 * 8 * 16;
      const double macs = kind == 2 ? (double)iters * 32 * warps * 32 : (double)iters * 8 * warps * per_mma;
      printf("%-28s warps/SM %2d : %8.1f MAC/clk/SM (%.0f cycles)\n",
             kind == 0 ? "mma.sync m16n8k8 tf32" : kind == 1 ? "mma.sync m16n8k16 f16->f32" : "ffma", warps, macs / avg, avg);
    }
  }
  for (int var = 0; var < 3; ++var) {
    const int it2 = 2000;
    for (int rep = 0; rep < 2; ++rep) {
      if (var == 0) step_kernel<0><<<sms, 256>>>(it2, out, cyc);
      else if (var == 1) step_kernel<1><<<sms, 256>>>(it2, out, cyc);
      else step_kernel<2><<<sms, 256>>>(it2, out, cyc);
      cudaDeviceSynchronize();
    }
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += (double)h[i];
    avg /= sms;
    printf("3xTF32 step from shared memory, variant %d (0 full, 1 no split, 2 MMAs only): %.0f cycles per k8 step of a 128x64 CTA tile "
           "(tensor floor 384), %.1f useful MAC/clk/SM\n", var, avg / (it2 * 4.0), 128.0 * 64 * 8 * it2 * 4 / avg);
  }
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
