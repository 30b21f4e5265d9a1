// How a streaming WRITE of 16 GB leaves the SMs on B200, on a cool chip and at the clock a power-capped tensor-core burst
// leaves behind: (a) per-thread 128-bit stores (what the Khatri-Rao producer does), (b) bulk copies shared -> global issued
// by one thread per CTA (cp.async.bulk, the TMA store path), (c) per-thread 64-bit stores (the fp16 planes).  Shared library
// driven from experiments/bench/store_paths.py, which supplies the burst (a pass-1 launch of the product library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -o libexp5_store_paths.so exp5_store_paths.cu
#include <cuda_runtime.h>
#include <cstdint>

__global__ void __launch_bounds__(256) stg128_kernel(float4* __restrict__ out, int64_t n4, float v) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float4 val = make_float4(v, v + 1.f, v + 2.f, v + 3.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) out[i] = val;
}

__global__ void __launch_bounds__(256) stg64_kernel(uint2* __restrict__ out, int64_t n2, uint32_t v) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const uint2 val = make_uint2(v, v + 1u);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) out[i] = val;
}

// every CTA owns a contiguous range of `chunk`-byte pieces; the tile in shared memory is filled once (the point is the
// store path, not the producer), then thread 0 keeps `depth` bulk stores in flight
template <int CHUNK>
__global__ void __launch_bounds__(256) bulk_kernel(uint8_t* __restrict__ out, int64_t nchunks, int depth) {
  extern __shared__ __align__(128) uint8_t tile[];
  for (int i = threadIdx.x; i < CHUNK / 16; i += blockDim.x) reinterpret_cast<uint4*>(tile)[i] = make_uint4(i, 1, 2, 3);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x != 0) return;
  const uint32_t src = (uint32_t)__cvta_generic_to_shared(tile);
  int inflight = 0;
  for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * CHUNK), "r"(src), "n"(CHUNK)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (++inflight >= depth) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // (all of them: simplest correct bound)
      inflight = 0;
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

extern "C" int exp5_stg128(void* out, long long bytes, int ctas, void* stream) {
  stg128_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>((float4*)out, bytes / 16, 1.f);
  return (int)cudaGetLastError();
}
extern "C" int exp5_stg64(void* out, long long bytes, int ctas, void* stream) {
  stg64_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>((uint2*)out, bytes / 8, 1u);
  return (int)cudaGetLastError();
}
extern "C" int exp5_bulk(void* out, long long bytes, int ctas, int chunk, int depth, void* stream) {
  if (chunk == 4096)
    bulk_kernel<4096><<<ctas, 256, 4096, (cudaStream_t)stream>>>((uint8_t*)out, bytes / 4096, depth);
  else if (chunk == 16384)
    bulk_kernel<16384><<<ctas, 256, 16384, (cudaStream_t)stream>>>((uint8_t*)out, bytes / 16384, depth);
  else
    bulk_kernel<32768><<<ctas, 256, 32768, (cudaStream_t)stream>>>((uint8_t*)out, bytes / 32768, depth);
  return (int)cudaGetLastError();
}
