// Throughput microbenchmarks that bound the attention softmax on sm_100a (one CTA on one SM, clock64 inside):
//   tmem_ld  : back-to-back tcgen05.ld.32x32b.x32 (4 KB per warp instruction) with W warps, each on its own lane quadrant
//   mufu     : MUFU.EX2 issue rate with W warps
//   ffma/ffma2/fadd2 : FMA-pipe issue rate, scalar vs packed
// Prints cycles per warp instruction (SM-wide) and the implied bytes or elements per clock.  Not part of the library.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(addr));
}

__global__ void __launch_bounds__(1024, 1) k_tmem(int iters, int depth, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t a[32], b[32];
    tmem_ld32(base + ((i * 64) & 511 & ~63), a);
    if (depth > 1) tmem_ld32(base + (((i * 64) & 511 & ~63) + 32), b);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < 32; k += 8) acc ^= a[k];
    if (depth > 1) {
#pragma unroll
      for (int k = 0; k < 32; k += 8) acc ^= b[k];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  sink[threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot));
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_alu(int iters, long long* out, float* sink, float seed) {
  float x[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = seed * (k + 1) + threadIdx.x * 1e-6f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {  // 16 MUFU.EX2
#pragma unroll
      for (int k = 0; k < 16; ++k) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[k]));
    } else if (MODE == 1) {  // 16 FFMA
#pragma unroll
      for (int k = 0; k < 16; ++k) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[k]) : "f"(seed));
    } else if (MODE == 2) {  // 8 FFMA2 (16 elements)
#pragma unroll
      for (int k = 0; k < 16; k += 2) {
        unsigned long long v, s;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[k]), "f"(x[k + 1]));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(s) : "f"(seed));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v) : "l"(s));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x[k]), "=f"(x[k + 1]) : "l"(v));
      }
    } else if (MODE == 3) {  // 8 MUFU + 8 FFMA interleaved
#pragma unroll
      for (int k = 0; k < 16; k += 2) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[k]));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[k + 1]) : "f"(seed));
      }
    } else if (MODE == 4) {  // 16 F2FP (pack) — which pipe?
#pragma unroll
      for (int k = 0; k < 16; k += 2) {
        uint32_t pk;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(x[k]), "f"(x[k + 1]));
        x[k] = __uint_as_float(pk);
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(x[k + 1]), "f"(x[k]));
        x[k + 1] = __uint_as_float(pk);
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  float s = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += x[k];
  sink[threadIdx.x] = s;
}

int main() {
  long long* out;
  uint32_t* sink;
  cudaMalloc(&out, 64);
  cudaMalloc(&sink, 4096 * 4);
  const int iters = 4096;
  for (int depth = 1; depth <= 2; ++depth)
    for (int warps : {1, 4, 8, 16, 32}) {
      k_tmem<<<1, warps * 32>>>(iters, depth, out, sink);
      long long c;
      if (cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("tmem: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      const double n = (double)iters * depth * warps;
      printf("tmem_ld.x32 depth %d warps %2d: %8.1f cycles per warp-load SM-wide, %6.1f B/clk/SM\n", depth, warps, c / n, n * 4096.0 / c);
    }
  const char* names[] = {"MUFU.EX2 x16", "FFMA x16", "FFMA2 x8 (16 el)", "MUFU x8 + FFMA x8", "F2FP x16"};
  for (int mode = 0; mode < 5; ++mode)
    for (int warps : {4, 8, 16}) {
      void (*k)(int, long long*, float*, float) = mode == 0 ? k_alu<0> : mode == 1 ? k_alu<1> : mode == 2 ? k_alu<2> : mode == 3 ? k_alu<3> : k_alu<4>;
      k<<<1, warps * 32>>>(iters, out, (float*)sink, 0.5f);
      long long c;
      if (cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("alu: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      const double n = (double)iters * 16 * warps;  // warp-level element-instructions
      printf("%-20s warps %2d: %6.2f SM-cycles per 32 elements, %6.1f elements/clk/SM\n", names[mode], warps, c / n, n * 32 / c);
    }
  return 0;
}
