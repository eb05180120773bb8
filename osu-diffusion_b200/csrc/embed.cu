// Sinusoidal feature producers, emitted directly as split-bf16 GEMM operands.
//
// Reference: positional_embedding.py:29-77 (timestep_embedding / offset_ / position_sequence_
// embedding: args = float32(v) * freqs, output [cos | sin]), FirstLayer.forward (models.py:227-233:
// columns [sincos(x*512) | sincos(y*384) | sincos(o/10) | c]), TimestepEmbedder (models.py:35-38),
// LabelEmbedder + "b = t + y" + the SiLU in front of every adaLN Linear (models.py:69-74,320,148,189).
//
// `freqs` (exp(-ln(1e4) k / half), fp32) is computed once on the host exactly as the reference
// computes it and passed in, so the fp32 product is the reference's; sincosf (accurate range
// reduction; arguments reach ~3e4 rad) must not be replaced by the fast intrinsics.
//
// Every output is written twice: hi = bf16(v) and lo = bf16(v - hi).  The first-layer / adaLN /
// t-MLP GEMMs then run as hi*Whi + lo*Whi + hi*Wlo on the bf16 tensor-core kernel, which keeps
// these small but precision-critical products at ~fp32 accuracy.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#include "common.h"
#include "ptx.cuh"

namespace osudit {

// lo == nullptr: `hi` is really a float* (fp32 mode, fp32_mode.cu) and receives v unsplit.
__device__ __forceinline__ void split_store(__nv_bfloat16* hi, __nv_bfloat16* lo, int64_t idx, float v) {
  if (lo == nullptr) {
    reinterpret_cast<float*>(hi)[idx] = v;
    return;
  }
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[idx] = h;
  lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
}

constexpr int kTokPerCta = 32;

// x (xrows, 2, T) f32, o (B, T) f32, c (B, E, T) f32 -> a_hi/a_lo (B*T, 384 + E) bf16.
// Row b reads x[b % xrows] (classifier-free guidance feeds both halves from the first half,
// models.py:332-333).
__global__ void __launch_bounds__(256)
embed_xoc_kernel(const float* __restrict__ x, const float* __restrict__ o,
                 const float* __restrict__ c, const float* __restrict__ freqs, float pf_x, float pf_y,
                 int xrows, int T, int E, __nv_bfloat16* __restrict__ a_hi,
                 __nv_bfloat16* __restrict__ a_lo, int x_only) {
  extern __shared__ float s_c[];  // [E][kTokPerCta + 1]
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kTokPerCta;
  const int Kout = 384 + E;
  const int tid = threadIdx.x;
  // x_only: the offset and context columns (272 of 528) do not change between the denoising steps of one sampling
  // loop (models.py:227-233: only x is the diffusion state); they are left as a previous full call wrote them
  const int ngrp = x_only ? 2 : 3;

  // context: coalesced read along T, transposed through shared memory
  for (int idx = tid; !x_only && idx < E * kTokPerCta; idx += 256) {
    const int e = idx / kTokPerCta;
    const int tt = idx % kTokPerCta;
    const int t = t0 + tt;
    s_c[e * (kTokPerCta + 1) + tt] =
        t < T ? c[(static_cast<int64_t>(b) * E + e) * T + t] : 0.f;
  }

  // sin/cos features: 3 groups (x, y, o) x 64 frequencies per token
  const float* xb = x + static_cast<int64_t>(b % xrows) * 2 * T;
  const float* ob = o + static_cast<int64_t>(b) * T;
  for (int idx = tid; idx < kTokPerCta * ngrp * 64; idx += 256) {
    const int tt = idx / (ngrp * 64);
    const int r = idx % (ngrp * 64);
    const int grp = r >> 6;
    const int k = r & 63;
    const int t = t0 + tt;
    if (t >= T) continue;
    float base;
    if (grp == 0) base = __fmul_rn(xb[t], pf_x);
    else if (grp == 1) base = __fmul_rn(xb[T + t], pf_y);
    else base = __fdiv_rn(ob[t], 10.0f);
    const float arg = __fmul_rn(base, __ldg(freqs + k));
    float sn, cs;
    sincosf(arg, &sn, &cs);
    const int64_t row = (static_cast<int64_t>(b) * T + t) * Kout;
    split_store(a_hi, a_lo, row + grp * 128 + k, cs);
    split_store(a_hi, a_lo, row + grp * 128 + 64 + k, sn);
  }
  __syncthreads();
  for (int idx = tid; !x_only && idx < kTokPerCta * E; idx += 256) {
    const int tt = idx / E;
    const int e = idx % E;
    const int t = t0 + tt;
    if (t >= T) continue;
    const int64_t row = (static_cast<int64_t>(b) * T + t) * Kout;
    split_store(a_hi, a_lo, row + 384 + e, s_c[e * (kTokPerCta + 1) + tt]);
  }
}

// t (rows,) int64 -> [cos | sin] (rows, 256) split-bf16.  freqs has 128 entries.
__global__ void timestep_features_kernel(const int64_t* __restrict__ t,
                                         const float* __restrict__ freqs, int rows,
                                         __nv_bfloat16* __restrict__ hi,
                                         __nv_bfloat16* __restrict__ lo) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * 128) return;
  const int r = idx >> 7;
  const int k = idx & 127;
  const float arg = __fmul_rn(static_cast<float>(t[r]), __ldg(freqs + k));
  float sn, cs;
  sincosf(arg, &sn, &cs);
  split_store(hi, lo, static_cast<int64_t>(r) * 256 + k, cs);
  split_store(hi, lo, static_cast<int64_t>(r) * 256 + 128 + k, sn);
}

__device__ __forceinline__ float silu(float v) { return v / (1.0f + expf(-v)); }

// out[r] = SiLU(a[ia[r]] (+ table[y[r]])) as split-bf16; ia == nullptr means ia[r] = r.
__global__ void silu_split_kernel(const float* __restrict__ a, const int32_t* __restrict__ ia,
                                  const float* __restrict__ table, const int64_t* __restrict__ y,
                                  int64_t rows, int D, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= rows * D) return;
  const int64_t r = idx / D;
  const int d = static_cast<int>(idx - r * D);
  const int64_t ar = ia ? ia[r] : r;
  float v = a[ar * D + d];
  if (table) v += table[y[r] * D + d];
  split_store(hi, lo, idx, silu(v));
}

// Device-side assertion on the class labels (what the reference's embedding lookup does on CUDA, models.py:73):
// a label outside [0, table_rows) stops the launch instead of reading / scattering outside the table.
__global__ void check_labels_kernel(const int64_t* __restrict__ y, int64_t n, int64_t table_rows) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n && (y[i] < 0 || y[i] >= table_rows)) {
    printf("osudit: class label %lld at row %lld is outside the embedding table [0, %lld)\n",
           static_cast<long long>(y[i]), static_cast<long long>(i), static_cast<long long>(table_rows));
    __trap();
  }
}

// fp32 -> split-bf16 (weights are packed once per parameter version with this).
__global__ void split_kernel(const float* __restrict__ a, int64_t n, __nv_bfloat16* __restrict__ hi,
                             __nv_bfloat16* __restrict__ lo) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const float v = a[idx];
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[idx] = h;
  if (lo) lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
}

}  // namespace osudit

using namespace osudit;

extern "C" int osudit_embed_xoc(const float* x, const float* o, const float* c, const float* freqs64,
                                float pf_x, float pf_y, int B, int xrows, int T, int E, void* a_hi,
                                void* a_lo, void* stream) {
  if (B <= 0 || T <= 0 || E < 0 || xrows <= 0) return set_error(-1, "embed_xoc: bad shape");
  if (B > 65535) return set_error(-1, "embed_xoc: batch too large for one launch");
  const size_t smem = static_cast<size_t>(E) * (kTokPerCta + 1) * sizeof(float);
  if (smem > 48 * 1024) return set_error(-1, "embed_xoc: context_size too large");
  dim3 grid((T + kTokPerCta - 1) / kTokPerCta, B);
  embed_xoc_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      x, o, c, freqs64, pf_x, pf_y, xrows, T, E, static_cast<__nv_bfloat16*>(a_hi),
      static_cast<__nv_bfloat16*>(a_lo), 0);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

extern "C" int osudit_embed_x(const float* x, const float* freqs64, float pf_x, float pf_y, int B, int xrows, int T,
                              int E, void* a_hi, void* a_lo, void* stream) {
  if (B <= 0 || T <= 0 || E < 0 || xrows <= 0) return set_error(-1, "embed_x: bad shape");
  if (B > 65535) return set_error(-1, "embed_x: batch too large for one launch");
  dim3 grid((T + kTokPerCta - 1) / kTokPerCta, B);
  embed_xoc_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, nullptr, nullptr, freqs64, pf_x, pf_y, xrows, T, E, static_cast<__nv_bfloat16*>(a_hi),
      static_cast<__nv_bfloat16*>(a_lo), 1);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

extern "C" int osudit_timestep_features(const int64_t* t, const float* freqs128, int rows, void* hi,
                                        void* lo, void* stream) {
  if (rows <= 0) return set_error(-1, "timestep_features: bad shape");
  const int n = rows * 128;
  timestep_features_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      t, freqs128, rows, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo));
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

extern "C" int osudit_embed_xoc_f32(const float* x, const float* o, const float* c, const float* freqs64,
                                    float pf_x, float pf_y, int B, int xrows, int T, int E, float* a,
                                    void* stream) {
  if (a == nullptr) return set_error(-1, "embed_xoc_f32: null output");
  return osudit_embed_xoc(x, o, c, freqs64, pf_x, pf_y, B, xrows, T, E, a, nullptr, stream);
}

extern "C" int osudit_timestep_features_f32(const int64_t* t, const float* freqs128, int rows, float* out,
                                            void* stream) {
  if (out == nullptr) return set_error(-1, "timestep_features_f32: null output");
  return osudit_timestep_features(t, freqs128, rows, out, nullptr, stream);
}

extern "C" int osudit_silu_split(const float* a, const int32_t* a_index, const float* table,
                                 const int64_t* y, int64_t rows, int D, void* hi, void* lo,
                                 void* stream) {
  if (rows <= 0 || D <= 0) return set_error(-1, "silu_split: bad shape");
  if ((table == nullptr) != (y == nullptr))
    return set_error(-1, "silu_split: table and y must be given together");
  const int64_t n = rows * D;
  silu_split_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0,
                      static_cast<cudaStream_t>(stream)>>>(
      a, a_index, table, y, rows, D, static_cast<__nv_bfloat16*>(hi),
      static_cast<__nv_bfloat16*>(lo));
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

extern "C" int osudit_check_labels(const int64_t* y, int64_t n, int64_t table_rows, void* stream) {
  if (n <= 0 || table_rows <= 0 || y == nullptr) return set_error(-1, "check_labels: bad arguments");
  check_labels_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      y, n, table_rows);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

extern "C" int osudit_split_bf16(const float* a, int64_t n, void* hi, void* lo, void* stream) {
  if (n <= 0) return set_error(-1, "split_bf16: bad shape");
  split_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0,
                 static_cast<cudaStream_t>(stream)>>>(a, n, static_cast<__nv_bfloat16*>(hi),
                                                      static_cast<__nv_bfloat16*>(lo));
  OSUDIT_CHECK_LAUNCH();
  return 0;
}
