// Microbenchmark: latency of tcgen05.commit -> mbarrier completion, with and without MMAs in
// flight, and of an mbarrier ping-pong between two warps.  Build: see tools/ubench/run.sh
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../osu-diffusion_b200/csrc/ptx.cuh"
using namespace osudit;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(128, 1) k(long long* out, int n_mma, int ncols) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, ping, pong;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&ping, 1); mbar_init(&pong, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc<512>(&slot); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 32768 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ncols >> 3) << 17) | ((128u >> 4) << 24);
  if (warp == 1 && lane == 0) {
    long long tot = 0, mx = 0;
    for (int it = 0; it < 64; ++it) {
      long long t0 = clock64();
      for (int m = 0; m < n_mma; ++m)
        umma_bf16(tm, desc_sw128(smem_u32(smem)), desc_sw128(smem_u32(smem) + 16384), idesc, m != 0);
      umma_commit(&bar);
      mbar_wait(&bar, it & 1);
      long long dt = clock64() - t0;
      if (it >= 8) { tot += dt; mx = dt > mx ? dt : mx; }
    }
    out[0] = tot / 56; out[1] = mx;
  }
  // ping-pong between warp 2 and warp 3 (single lanes)
  if (warp == 2 && lane == 0) {
    long long t0 = clock64();
    for (int it = 0; it < 64; ++it) { mbar_arrive(&ping); mbar_wait(&pong, it & 1); }
    out[2] = (clock64() - t0) / 64;
  }
  if (warp == 3 && lane == 0) {
    for (int it = 0; it < 64; ++it) { mbar_wait(&ping, it & 1); mbar_arrive(&pong); }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tm);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int ncols : {64, 128, 256}) for (int n : {0, 1, 4, 12, 24}) {
    k<<<1, 128, 65536>>>(d, n, ncols);
    long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("N=%3d n_mma=%2d: issue+commit+wait avg %lld max %lld cycles | mbarrier ping-pong round trip %lld cycles (err %s)\n",
           ncols, n, h[0], h[1], h[2], cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
