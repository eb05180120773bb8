// Fused adaLN-Zero residual update + LayerNorm + modulate, and the final layer.
//
// Reference: DiTBlock.forward / modulate / FinalLayer.forward (models.py:12-13,151-175,192-196):
//     x = x + gate[b] * branch            (branch = attention or MLP output of the previous half-block)
//     h = LN(x) * (1 + scale[b]) + shift[b]   (LayerNorm eps 1e-6, no affine)
// The residual stream x stays fp32 in HBM (a bf16 residual costs ~9e-3 rel-L2); the
// branch output arrives as bf16 from the GEMM epilogue and h leaves as the bf16 A-operand of the next
// GEMM.  HBM-bound: 12*D bytes per token with the residual update (read x 4D + y 2D, write x 4D +
// h 2D), 6*D without.  One warp per token row, 128-bit loads of x, 64-bit of y/h with the same
// lane->column mapping, statistics by warp shuffles, everything else in registers.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"
#include "ptx.cuh"

namespace osudit {

constexpr float kLnEps = 1e-6f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// NV = D / 128: float4 vectors per lane.
// The read-only variant (no pending residual update) streams x once and h once: loads bypass L1 allocation, stores are
// marked streaming.  Measured at config 2 (tools/resid_bench.py): 0.232 -> 0.205 ms after the out-projection, 0.231 ->
// 0.227 ms after fc2.  LN_HINTS=0 builds the plain loads / stores.
#ifndef LN_HINTS
#define LN_HINTS 1
#endif
__device__ __forceinline__ float4 ld_stream(const float* p) {
#if LN_HINTS
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
#else
  return *reinterpret_cast<const float4*>(p);
#endif
}
__device__ __forceinline__ void st_stream(__nv_bfloat16* p, uint2 v) {
#if LN_HINTS
  asm volatile("st.global.cs.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
#else
  *reinterpret_cast<uint2*>(p) = v;
#endif
}

template <int NV, bool kHasBranch>
__device__ __forceinline__ void load_row_update(float4 (&v)[NV], float* __restrict__ xrow,
                                                const __nv_bfloat16* __restrict__ yrow,
                                                const float* __restrict__ gate, int lane) {
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = kHasBranch ? *reinterpret_cast<const float4*>(xrow + (lane + 32 * i) * 4) : ld_stream(xrow + (lane + 32 * i) * 4);
  if (kHasBranch) {
    uint2 yv[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i)
      yv[i] = *reinterpret_cast<const uint2*>(yrow + (lane + 32 * i) * 4);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gate + (lane + 32 * i) * 4));
      const float2 y01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yv[i].x));
      const float2 y23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yv[i].y));
      v[i].x = fmaf(g.x, y01.x, v[i].x);
      v[i].y = fmaf(g.y, y01.y, v[i].y);
      v[i].z = fmaf(g.z, y23.x, v[i].z);
      v[i].w = fmaf(g.w, y23.y, v[i].w);
    }
  }
}

template <int NV>
__device__ __forceinline__ void row_stats(const float4 (&v)[NV], float& mean, float& rstd) {
  constexpr float inv_d = 1.0f / (NV * 128);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  mean = warp_sum(s) * inv_d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  rstd = rsqrtf(warp_sum(q) * inv_d + kLnEps);
}

template <int NV, bool kHasBranch>
__global__ void __launch_bounds__(256)
ln_modulate_kernel(float* __restrict__ x, const __nv_bfloat16* __restrict__ y,
                   const float* __restrict__ gate, const float* __restrict__ shift,
                   const float* __restrict__ scale, int64_t mod_ld, int64_t rows, int T,
                   __nv_bfloat16* __restrict__ h, float* __restrict__ x_out) {
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int64_t b = row / T;
  float* xrow = x + row * D;
  float4 v[NV];
  load_row_update<NV, kHasBranch>(v, xrow, kHasBranch ? y + row * D : nullptr,
                                  kHasBranch ? gate + b * mod_ld : nullptr, lane);
  if (kHasBranch) {  // updated residual: in place, or to x_out when the caller keeps x (training)
    float* xw = x_out != nullptr ? x_out + row * D : xrow;
#pragma unroll
    for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(xw + (lane + 32 * i) * 4) = v[i];
  }
  float mean, rstd;
  row_stats<NV>(v, mean, rstd);
  const float* sh = shift + b * mod_ld;
  const float* sc = scale + b * mod_ld;
  __nv_bfloat16* hrow = h + row * D;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float4 s4 = __ldg(reinterpret_cast<const float4*>(sc + c));
    const float4 t4 = __ldg(reinterpret_cast<const float4*>(sh + c));
    const float o0 = fmaf((v[i].x - mean) * rstd, 1.0f + s4.x, t4.x);
    const float o1 = fmaf((v[i].y - mean) * rstd, 1.0f + s4.y, t4.y);
    const float o2 = fmaf((v[i].z - mean) * rstd, 1.0f + s4.z, t4.z);
    const float o3 = fmaf((v[i].w - mean) * rstd, 1.0f + s4.w, t4.w);
    uint2 o;
    o.x = pack_bf16(o0, o1);
    o.y = pack_bf16(o2, o3);
    if (kHasBranch) *reinterpret_cast<uint2*>(hrow + c) = o; else st_stream(hrow + c, o);
  }
}

// FinalLayer: (pending residual update) -> LN -> modulate -> Linear D -> 4, written channel-major
// (B, 4, T) like DiT.forward's output (models.py:323-324).  All fp32 (SURVEY F16 keeps the final
// projection out of bf16).
template <int NV, bool kHasBranch>
__global__ void __launch_bounds__(256)
final_layer_kernel(float* __restrict__ x, const __nv_bfloat16* __restrict__ y,
                   const float* __restrict__ gate, const float* __restrict__ shift,
                   const float* __restrict__ scale, int64_t mod_ld, int64_t rows, int T,
                   const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ out) {
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int64_t b = row / T;
  const int64_t t = row - b * T;
  float4 v[NV];
  load_row_update<NV, kHasBranch>(v, x + row * D, kHasBranch ? y + row * D : nullptr,
                                  kHasBranch ? gate + b * mod_ld : nullptr, lane);
  float mean, rstd;
  row_stats<NV>(v, mean, rstd);
  const float* sh = shift + b * mod_ld;
  const float* sc = scale + b * mod_ld;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float4 s4 = __ldg(reinterpret_cast<const float4*>(sc + c));
    const float4 t4 = __ldg(reinterpret_cast<const float4*>(sh + c));
    const float h0 = fmaf((v[i].x - mean) * rstd, 1.0f + s4.x, t4.x);
    const float h1 = fmaf((v[i].y - mean) * rstd, 1.0f + s4.y, t4.y);
    const float h2 = fmaf((v[i].z - mean) * rstd, 1.0f + s4.z, t4.z);
    const float h3 = fmaf((v[i].w - mean) * rstd, 1.0f + s4.w, t4.w);
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + o * D + c));
      acc[o] = fmaf(h0, w4.x, fmaf(h1, w4.y, fmaf(h2, w4.z, fmaf(h3, w4.w, acc[o]))));
    }
  }
#pragma unroll
  for (int o = 0; o < 4; ++o) acc[o] = warp_sum(acc[o]);
  if (lane < 4) {
    const float r = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
    out[(b * 4 + lane) * T + t] = r + __ldg(bias + lane);
  }
}

template <template <int, bool> class Launcher, typename... Args>
static int dispatch_nv(int D, bool has_branch, Args... args) {
  if (D % 128 != 0 || D < 128 || D > 1536)
    return set_error(-1, "hidden size must be a multiple of 128 in [128, 1536]");
  const int nv = D / 128;
#define OSUDIT_NV_CASE(N)                                                    \
  case N:                                                                    \
    return has_branch ? Launcher<N, true>::run(args...) : Launcher<N, false>::run(args...);
  switch (nv) {
    OSUDIT_NV_CASE(1) OSUDIT_NV_CASE(2) OSUDIT_NV_CASE(3) OSUDIT_NV_CASE(4) OSUDIT_NV_CASE(5)
    OSUDIT_NV_CASE(6) OSUDIT_NV_CASE(7) OSUDIT_NV_CASE(8) OSUDIT_NV_CASE(9) OSUDIT_NV_CASE(10)
    OSUDIT_NV_CASE(11) OSUDIT_NV_CASE(12)
  }
#undef OSUDIT_NV_CASE
  return set_error(-1, "unreachable");
}

template <int NV, bool HB>
struct LnLauncher {
  static int run(float* x, const __nv_bfloat16* y, const float* gate, const float* shift,
                 const float* scale, int64_t mod_ld, int64_t rows, int T, __nv_bfloat16* h,
                 float* x_out, cudaStream_t st) {
    const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
    ln_modulate_kernel<NV, HB><<<grid, 256, 0, st>>>(x, y, gate, shift, scale, mod_ld, rows, T, h, x_out);
    OSUDIT_CHECK_LAUNCH();
    return 0;
  }
};

template <int NV, bool HB>
struct FinalLauncher {
  static int run(float* x, const __nv_bfloat16* y, const float* gate, const float* shift,
                 const float* scale, int64_t mod_ld, int64_t rows, int T, const float* w,
                 const float* bias, float* out, cudaStream_t st) {
    const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
    final_layer_kernel<NV, HB><<<grid, 256, 0, st>>>(x, y, gate, shift, scale, mod_ld, rows, T, w,
                                                     bias, out);
    OSUDIT_CHECK_LAUNCH();
    return 0;
  }
};

}  // namespace osudit

using namespace osudit;

extern "C" int osudit_ln_modulate(float* x, const void* branch, const float* gate,
                                  const float* shift, const float* scale, int64_t mod_ld,
                                  int64_t rows, int T, int D, void* h, float* x_out, void* stream) {
  if (rows <= 0 || T <= 0) return set_error(-1, "ln_modulate: bad shape");
  if ((branch == nullptr) != (gate == nullptr))
    return set_error(-1, "ln_modulate: branch and gate must be given together");
  return dispatch_nv<LnLauncher>(D, branch != nullptr, x, static_cast<const __nv_bfloat16*>(branch),
                                 gate, shift, scale, mod_ld, rows, T,
                                 static_cast<__nv_bfloat16*>(h), x_out, static_cast<cudaStream_t>(stream));
}

extern "C" int osudit_final_layer(float* x, const void* branch, const float* gate,
                                  const float* shift, const float* scale, int64_t mod_ld,
                                  int64_t rows, int T, int D, const float* w, const float* bias,
                                  int out_channels, float* out, void* stream) {
  if (rows <= 0 || T <= 0) return set_error(-1, "final_layer: bad shape");
  if (out_channels != 4) return set_error(-1, "final_layer: out_channels must be 4 (in_channels=2, learn_sigma)");
  if ((branch == nullptr) != (gate == nullptr))
    return set_error(-1, "final_layer: branch and gate must be given together");
  return dispatch_nv<FinalLauncher>(D, branch != nullptr, x,
                                    static_cast<const __nv_bfloat16*>(branch), gate, shift, scale,
                                    mod_ld, rows, T, w, bias, out,
                                    static_cast<cudaStream_t>(stream));
}
