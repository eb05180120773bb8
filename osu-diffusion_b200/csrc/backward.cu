// Backward-pass kernels of the DiT block that are not GEMMs or attention: transposes (the weight- and
// data-gradient GEMMs reuse gemm_tcgen05_kernel on transposed operands), the gated-residual and
// adaLN LayerNorm-modulate backward with their per-sample (shift / scale / gate) reductions, GELU,
// bias column sums, the final layer and the label-embedding scatter.
//
// Reference semantics being differentiated: DiTBlock.forward / modulate / FinalLayer.forward
// (models.py:12-13,151-175,192-196), nn.GELU(approximate="tanh") (models.py:138), LabelEmbedder
// (models.py:69-74).  The reference gets these gradients from autograd; train.py:257 is the call site.
//
// Per-sample adaLN gradients (d shift, d scale, d gate: one row of the [B, depth*6D+2D] modulation
// matrix per batch element) are reduced over a CTA's rows in shared memory and flushed with one
// global atomicAdd per column per CTA; the host zeroes the dmod matrix once per step.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"
#include "ptx.cuh"

namespace osudit {

constexpr float kLnEpsB = 1e-6f;

__device__ __forceinline__ float warp_sum_b(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------ transpose
// out[C, R] = in[R, C]^T (bf16), 64x64 tiles through shared memory.
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t R,
                      int64_t C, int64_t LDO) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * 64, c0 = static_cast<int64_t>(blockIdx.x) * 64;
  for (int i = threadIdx.x; i < 64 * 32; i += 256) {
    const int r = i / 32, c2 = (i % 32) * 2;
    __nv_bfloat162 v = __floats2bfloat162_rn(0.f, 0.f);
    if (r0 + r < R) {
      if (c0 + c2 + 1 < C) v = *reinterpret_cast<const __nv_bfloat162*>(in + (r0 + r) * C + c0 + c2);
      else if (c0 + c2 < C) v.x = in[(r0 + r) * C + c0 + c2];
    }
    tile[r][c2] = v.x;
    tile[r][c2 + 1] = v.y;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 32; i += 256) {
    const int c = i / 32, r2 = (i % 32) * 2;  // output row = input column
    if (c0 + c < C) {
      if (r0 + r2 + 1 < R && (LDO % 2) == 0)
        *reinterpret_cast<__nv_bfloat162*>(out + (c0 + c) * LDO + r0 + r2) =
            __halves2bfloat162(tile[r2][c], tile[r2 + 1][c]);
      else {
        if (r0 + r2 < R) out[(c0 + c) * LDO + r0 + r2] = tile[r2][c];
        if (r0 + r2 + 1 < R) out[(c0 + c) * LDO + r0 + r2 + 1] = tile[r2 + 1][c];
      }
    }
  }
}

// fp32 -> bf16 transpose (for gradients held in fp32, e.g. the residual-stream gradient feeding the
// first-layer weight gradient, or dmod feeding the adaLN weight gradient).
__global__ void __launch_bounds__(256)
transpose_f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t R,
                             int64_t C, int64_t LDO) {
  __shared__ float tile[64][65];
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * 64, c0 = static_cast<int64_t>(blockIdx.x) * 64;
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    const int r = i / 64, c = i % 64;
    tile[r][c] = (r0 + r < R && c0 + c < C) ? in[(r0 + r) * C + c0 + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    const int c = i / 64, r = i % 64;
    if (c0 + c < C && r0 + r < R) out[(c0 + c) * LDO + r0 + r] = __float2bfloat16_rn(tile[r][c]);
  }
}

// ------------------------------------------------------------------ batched weight re-pack
// Every GEMM-ready copy of every fp32 nn.Linear weight in ONE launch (the training step re-packs all of them after each
// optimizer update, train.py:258-261): same-layout bf16 (forward operand), transposed bf16 (data-gradient operand),
// split-bf16 hi / lo (the precision-critical small GEMMs).  One CTA per 64 x 64 tile of some matrix; the per-matrix
// launches it replaces (50 transposes + 48 casts + 21 splits of 10 us each) were launch-latency bound at ~1 TB/s.
struct RepackSeg {       // mirrored by osudit/train.py (numpy structured dtype) and include/osudit.h
  const float* src;      // fp32 [rows, cols] row-major
  __nv_bfloat16* copy;   // bf16 [rows, cols] or null
  __nv_bfloat16* trans;  // bf16 [cols, ld_trans] (element (c, r) = src[r][c]) or null
  __nv_bfloat16* hi;     // split-bf16 [rows, cols] or null
  __nv_bfloat16* lo;
  int64_t ld_trans;
  int32_t rows, cols;
  int32_t tile0, tiles_x;  // first tile index of this segment in the launch, tiles per row of tiles
};

__global__ void __launch_bounds__(256) repack_weights_kernel(const RepackSeg* __restrict__ segs, int nseg) {
  __shared__ float tile[64][65];
  __shared__ RepackSeg sg;
  if (threadIdx.x == 0) {
    int lo = 0, hi = nseg - 1;  // last segment whose tile0 <= blockIdx.x
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (segs[mid].tile0 <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid - 1;
    }
    sg = segs[lo];
  }
  __syncthreads();
  const int t = static_cast<int>(blockIdx.x) - sg.tile0;
  const int r0 = (t / sg.tiles_x) * 64, c0 = (t % sg.tiles_x) * 64;
  const bool vec = (sg.cols & 3) == 0 && (sg.rows & 3) == 0 && (sg.ld_trans & 3) == 0;  // every nn.Linear of the model
  if (vec) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {  // 16-byte loads, 8-byte bf16 stores
      const int r = i >> 4, c = (i & 15) * 4;
      const bool in = r0 + r < sg.rows && c0 + c < sg.cols;
      const int64_t idx = static_cast<int64_t>(r0 + r) * sg.cols + c0 + c;
      const float4 v = in ? *reinterpret_cast<const float4*>(sg.src + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
      tile[r][c] = v.x; tile[r][c + 1] = v.y; tile[r][c + 2] = v.z; tile[r][c + 3] = v.w;
      if (in) {
        const uint2 h = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
        if (sg.copy) *reinterpret_cast<uint2*>(sg.copy + idx) = h;
        if (sg.hi) {
          *reinterpret_cast<uint2*>(sg.hi + idx) = h;
          const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h.x));
          const float2 b2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h.y));
          *reinterpret_cast<uint2*>(sg.lo + idx) = make_uint2(pack_bf16(v.x - a.x, v.y - a.y), pack_bf16(v.z - b2.x, v.w - b2.y));
        }
      }
    }
    if (sg.trans == nullptr) return;
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int c = i >> 4, r = (i & 15) * 4;
      if (c0 + c < sg.cols && r0 + r < sg.rows)
        *reinterpret_cast<uint2*>(sg.trans + static_cast<int64_t>(c0 + c) * sg.ld_trans + r0 + r) =
            make_uint2(pack_bf16(tile[r][c], tile[r + 1][c]), pack_bf16(tile[r + 2][c], tile[r + 3][c]));
    }
    return;
  }
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    const int r = i >> 6, c = i & 63;
    const bool in = r0 + r < sg.rows && c0 + c < sg.cols;
    const int64_t idx = static_cast<int64_t>(r0 + r) * sg.cols + c0 + c;
    const float v = in ? sg.src[idx] : 0.f;
    tile[r][c] = v;
    if (in) {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      if (sg.copy) sg.copy[idx] = h;
      if (sg.hi) {
        sg.hi[idx] = h;
        sg.lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
      }
    }
  }
  if (sg.trans == nullptr) return;
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    const int c = i >> 6, r = i & 63;
    if (c0 + c < sg.cols && r0 + r < sg.rows)
      sg.trans[static_cast<int64_t>(c0 + c) * sg.ld_trans + r0 + r] = __float2bfloat16_rn(tile[r][c]);
  }
}

// ---------------------------------------------------------------------------------- GELU
__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.0f + tanhf(u));
}
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  const float x2 = x * x;
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x2);
  const float t = tanhf(u);
  const float du = 0.7978845608028654f * (1.0f + 3.0f * 0.044715f * x2);
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * du;
}

// mode 0: out = gelu(pre); mode 1: out = dy * gelu'(pre)
__global__ void __launch_bounds__(256)
gelu_kernel(const __nv_bfloat16* __restrict__ pre, const __nv_bfloat16* __restrict__ dy,
            __nv_bfloat16* __restrict__ out, int64_t n8, int mode) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  uint4 p = reinterpret_cast<const uint4*>(pre)[i];
  uint4 g = mode ? reinterpret_cast<const uint4*>(dy)[i] : make_uint4(0, 0, 0, 0);
  uint32_t* pp = reinterpret_cast<uint32_t*>(&p);
  uint32_t* gg = reinterpret_cast<uint32_t*>(&g);
  uint4 o;
  uint32_t* oo = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 x = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&pp[k]));
    if (mode == 0) {
      oo[k] = pack_bf16(gelu_tanh_f(x.x), gelu_tanh_f(x.y));
    } else {
      const float2 d = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&gg[k]));
      oo[k] = pack_bf16(d.x * gelu_tanh_grad(x.x), d.y * gelu_tanh_grad(x.y));
    }
  }
  reinterpret_cast<uint4*>(out)[i] = o;
}

// Two-launch path of osudit_gemm_bf16_aux (shapes the CTA-pair kernel does not take):
// mode 0: out = gelu(aux), aux = gelu'(aux) in place;  mode 1: out = out * aux.
__global__ void __launch_bounds__(256)
gelu_aux_kernel(__nv_bfloat16* __restrict__ aux, __nv_bfloat16* __restrict__ out, int64_t n8, int mode) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  uint4 a = reinterpret_cast<const uint4*>(aux)[i];
  uint4 o = mode ? reinterpret_cast<const uint4*>(out)[i] : make_uint4(0, 0, 0, 0);
  uint32_t* aa = reinterpret_cast<uint32_t*>(&a);
  uint32_t* oo = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 x = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&aa[k]));
    if (mode == 0) {
      oo[k] = pack_bf16(gelu_tanh_f(x.x), gelu_tanh_f(x.y));
      aa[k] = pack_bf16(gelu_tanh_grad(x.x), gelu_tanh_grad(x.y));
    } else {
      const float2 y = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&oo[k]));
      oo[k] = pack_bf16(y.x * x.x, y.y * x.y);
    }
  }
  reinterpret_cast<uint4*>(out)[i] = o;
  if (mode == 0) reinterpret_cast<uint4*>(aux)[i] = a;
}

int gelu_aux_launch(void* aux, void* out, int64_t n, int mode, cudaStream_t st) {
  if (n <= 0 || (n % 8) != 0) return set_error(-1, "gelu_aux: element count must be a positive multiple of 8");
  gelu_aux_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, st>>>(
      static_cast<__nv_bfloat16*>(aux), static_cast<__nv_bfloat16*>(out), n / 8, mode);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

// out[r, c] = dy[r, c] * gelu'(pre[r, c]) and dbias[c] += sum_r out[r, c] (the fc1 bias gradient):
// thread = 8 consecutive columns x kGeluRows rows, one atomic per column per CTA.
constexpr int kGeluRows = 64;
__global__ void __launch_bounds__(256)
gelu_bwd_colsum_kernel(const __nv_bfloat16* __restrict__ pre, const __nv_bfloat16* __restrict__ dy,
                       __nv_bfloat16* __restrict__ out, float* __restrict__ dbias, int64_t rows, int N) {
  const int c = (blockIdx.x * 256 + threadIdx.x) * 8;
  if (c >= N) return;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * kGeluRows;
  const int64_t r1 = r0 + kGeluRows < rows ? r0 + kGeluRows : rows;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t rb = r0; rb < r1; rb += 4) {
    uint4 p[4], g[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // 8 independent 16-byte loads in flight per thread
      if (rb + j < r1) {
        p[j] = *reinterpret_cast<const uint4*>(pre + (rb + j) * N + c);
        g[j] = *reinterpret_cast<const uint4*>(dy + (rb + j) * N + c);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (rb + j >= r1) break;
      uint32_t* pp = reinterpret_cast<uint32_t*>(&p[j]);
      uint32_t* gg = reinterpret_cast<uint32_t*>(&g[j]);
      uint4 o;
      uint32_t* oo = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 x = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&pp[k]));
        const float2 d = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&gg[k]));
        const float v0 = d.x * gelu_tanh_grad(x.x), v1 = d.y * gelu_tanh_grad(x.y);
        acc[2 * k] += v0;
        acc[2 * k + 1] += v1;
        oo[k] = pack_bf16(v0, v1);
      }
      *reinterpret_cast<uint4*>(out + (rb + j) * N + c) = o;
    }
  }
  if (dbias != nullptr) {
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(dbias + c + k, acc[k]);
  }
}

// ------------------------------------------------------------------------------- col sums
// out[N] (fp32, accumulated with atomics; host zeroes) += sum over rows of in[rows, N].
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ in, float* __restrict__ out, int64_t rows, int N, int rows_per_cta) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= N) return;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
  float acc = 0.f;
  for (int64_t r = r0; r < r1; ++r) acc += static_cast<float>(in[r * N + c]);
  atomicAdd(out + c, acc);
}

// -------------------------------------------------------------------- gated residual backward
// forward: x_out = x + gate[b] * y.   dy (bf16) = gate[b] * dx;  dgate[b] += sum_t dx * y.
// One CTA = up to 32 rows of one batch element; thread = 4 consecutive columns.
__global__ void __launch_bounds__(384)
gate_residual_bwd_kernel(const float* __restrict__ dx, const __nv_bfloat16* __restrict__ y,
                         const float* __restrict__ gate, float* __restrict__ dgate, int64_t mod_ld,
                         int T, int D, __nv_bfloat16* __restrict__ dy, float* __restrict__ dbias) {
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * 32;
  const int c = threadIdx.x * 4;
  if (c >= D) return;
  const float4 g = *reinterpret_cast<const float4*>(gate + static_cast<int64_t>(b) * mod_ld + c);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);  // column sums of dy: the bias gradient of the Linear before
  const int t1 = t0 + 32 < T ? t0 + 32 : T;
  for (int t = t0; t < t1; ++t) {
    const int64_t off = (static_cast<int64_t>(b) * T + t) * D + c;
    const float4 d = *reinterpret_cast<const float4*>(dx + off);
    const uint2 yv = *reinterpret_cast<const uint2*>(y + off);
    const float2 y01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yv.x));
    const float2 y23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yv.y));
    acc.x += d.x * y01.x; acc.y += d.y * y01.y; acc.z += d.z * y23.x; acc.w += d.w * y23.y;
    const float4 gd = make_float4(g.x * d.x, g.y * d.y, g.z * d.z, g.w * d.w);
    bsum.x += gd.x; bsum.y += gd.y; bsum.z += gd.z; bsum.w += gd.w;
    uint2 o;
    o.x = pack_bf16(gd.x, gd.y);
    o.y = pack_bf16(gd.z, gd.w);
    *reinterpret_cast<uint2*>(dy + off) = o;
  }
  if (dbias != nullptr) {
    atomicAdd(dbias + c + 0, bsum.x); atomicAdd(dbias + c + 1, bsum.y);
    atomicAdd(dbias + c + 2, bsum.z); atomicAdd(dbias + c + 3, bsum.w);
  }
  float* dg = dgate + static_cast<int64_t>(b) * mod_ld + c;
  atomicAdd(dg + 0, acc.x); atomicAdd(dg + 1, acc.y); atomicAdd(dg + 2, acc.z); atomicAdd(dg + 3, acc.w);
}

// ------------------------------------------------------------- LayerNorm-modulate backward
// forward: h = LN(x) * (1 + scale[b]) + shift[b].  Given dh (bf16):
//   dshift[b] += sum_t dh;  dscale[b] += sum_t dh * LN(x);  g = dh * (1 + scale[b]);
//   dx_acc += rstd * (g - mean(g) - n * mean(g * n)),  n = (x - mean) * rstd.
// One warp per row (NV float4 per lane); 8 rows of ONE batch element per CTA so the per-sample
// reductions go through shared-memory atomics and one global atomic per column per CTA.
template <int NV>
__global__ void __launch_bounds__(256)
ln_modulate_bwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dh,
                       const float* __restrict__ scale, float* __restrict__ dshift,
                       float* __restrict__ dscale, int64_t mod_ld, int T, float* __restrict__ dx_acc,
                       int accumulate) {
  constexpr int D = NV * 128;
  __shared__ float s_dshift[D];
  __shared__ float s_dscale[D];
  for (int i = threadIdx.x; i < D; i += 256) { s_dshift[i] = 0.f; s_dscale[i] = 0.f; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (t < T) {
    const int64_t row = static_cast<int64_t>(b) * T + t;
    const float* xrow = x + row * D;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = *reinterpret_cast<const float4*>(xrow + (lane + 32 * i) * 4);
    constexpr float inv_d = 1.0f / D;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum_b(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + bb * bb) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum_b(q) * inv_d + kLnEpsB);
    const float* sc = scale + static_cast<int64_t>(b) * mod_ld;
    const __nv_bfloat16* dhrow = dh + row * D;
    float4 g[NV];
    float sg = 0.f, sgn = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + 32 * i) * 4;
      const uint2 dv = *reinterpret_cast<const uint2*>(dhrow + c);
      const float2 d01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dv.x));
      const float2 d23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dv.y));
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(sc + c));
      v[i].x = (v[i].x - mean) * rstd; v[i].y = (v[i].y - mean) * rstd;
      v[i].z = (v[i].z - mean) * rstd; v[i].w = (v[i].w - mean) * rstd;
      atomicAdd(&s_dshift[c + 0], d01.x); atomicAdd(&s_dshift[c + 1], d01.y);
      atomicAdd(&s_dshift[c + 2], d23.x); atomicAdd(&s_dshift[c + 3], d23.y);
      atomicAdd(&s_dscale[c + 0], d01.x * v[i].x); atomicAdd(&s_dscale[c + 1], d01.y * v[i].y);
      atomicAdd(&s_dscale[c + 2], d23.x * v[i].z); atomicAdd(&s_dscale[c + 3], d23.y * v[i].w);
      g[i] = make_float4(d01.x * (1.f + s4.x), d01.y * (1.f + s4.y), d23.x * (1.f + s4.z),
                         d23.y * (1.f + s4.w));
      sg += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      sgn += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
    }
    const float mg = warp_sum_b(sg) * inv_d;
    const float mgn = warp_sum_b(sgn) * inv_d;
    float* drow = dx_acc + row * D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + 32 * i) * 4;
      float4 o = make_float4(rstd * (g[i].x - mg - v[i].x * mgn), rstd * (g[i].y - mg - v[i].y * mgn),
                             rstd * (g[i].z - mg - v[i].z * mgn), rstd * (g[i].w - mg - v[i].w * mgn));
      if (accumulate) {
        const float4 old = *reinterpret_cast<const float4*>(drow + c);
        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
      }
      *reinterpret_cast<float4*>(drow + c) = o;
    }
  }
  __syncthreads();
  float* gs = dshift + static_cast<int64_t>(b) * mod_ld;
  float* gc = dscale + static_cast<int64_t>(b) * mod_ld;
  for (int i = threadIdx.x; i < D; i += 256) {
    atomicAdd(gs + i, s_dshift[i]);
    atomicAdd(gc + i, s_dscale[i]);
  }
}

// ------------------------------------------- LayerNorm-modulate backward + gated residual backward
// The two steps that follow each other along the residual stream in the backward pass, in ONE pass over dx:
//   dx_new = dx_old + LN-modulate-backward(x, dh)                     (as ln_modulate_bwd_kernel)
//   dy     = gate[b] * dx_new,  dgate[b] += sum_t dx_new * y,  dbias += sum_rows dy   (as gate_residual_bwd)
// where y / gate belong to the branch whose output was added to the residual just BEFORE this LayerNorm in
// the forward.  18*D B/token instead of 14*D + 12*D for the two separate kernels.  One warp owns kRows
// consecutive rows; the four per-column sums are kept in a warp-private shared-memory strip (plain
// read-modify-write, no atomics), reduced over the CTA's 4 warps at the end: one global atomic per column
// per CTA per sum.
constexpr int kLgWarps = 4;
constexpr int kLgRows = 8;  // rows per warp

template <int NV, bool kGate>
__global__ void __launch_bounds__(kLgWarps * 32)
ln_gate_bwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dh,
                   const float* __restrict__ scale, float* __restrict__ dshift, float* __restrict__ dscale,
                   int64_t mod_ld, int T, float* __restrict__ dx_acc, int accumulate,
                   const __nv_bfloat16* __restrict__ y, const float* __restrict__ gate,
                   float* __restrict__ dgate, __nv_bfloat16* __restrict__ dy, float* __restrict__ dbias) {
  constexpr int D = NV * 128;
  constexpr int kArr = kGate ? 4 : 2;
  extern __shared__ float s_strip[];  // [kLgWarps][kArr][D]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  float* mine = s_strip + warp * kArr * D;
  for (int i = lane; i < kArr * D; i += 32) mine[i] = 0.f;
  __syncwarp();
  const float* sc = scale + static_cast<int64_t>(b) * mod_ld;
  const float* gt = kGate ? gate + static_cast<int64_t>(b) * mod_ld : nullptr;
  constexpr float inv_d = 1.0f / D;
  const int t_begin = (blockIdx.x * kLgWarps + warp) * kLgRows;
  for (int t = t_begin; t < t_begin + kLgRows && t < T; ++t) {
    const int64_t row = static_cast<int64_t>(b) * T + t;
    const float* xrow = x + row * D;
    const __nv_bfloat16* dhrow = dh + row * D;
    float* drow = dx_acc + row * D;
    float4 v[NV], old[NV];
    uint2 dv[NV], yv[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {  // all loads of the row in flight before the first reduction
      const int c = (lane + 32 * i) * 4;
      v[i] = *reinterpret_cast<const float4*>(xrow + c);
      dv[i] = *reinterpret_cast<const uint2*>(dhrow + c);
      if (accumulate) old[i] = *reinterpret_cast<const float4*>(drow + c);
      if (kGate) yv[i] = *reinterpret_cast<const uint2*>(y + row * D + c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum_b(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + bb * bb) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum_b(q) * inv_d + kLnEpsB);
    float4 g[NV];
    float sg = 0.f, sgn = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + 32 * i) * 4;
      const float2 d01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dv[i].x));
      const float2 d23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dv[i].y));
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(sc + c));
      v[i].x = (v[i].x - mean) * rstd; v[i].y = (v[i].y - mean) * rstd;
      v[i].z = (v[i].z - mean) * rstd; v[i].w = (v[i].w - mean) * rstd;
      float4 a = *reinterpret_cast<float4*>(mine + c);  // dshift
      a.x += d01.x; a.y += d01.y; a.z += d23.x; a.w += d23.y;
      *reinterpret_cast<float4*>(mine + c) = a;
      a = *reinterpret_cast<float4*>(mine + D + c);  // dscale
      a.x += d01.x * v[i].x; a.y += d01.y * v[i].y; a.z += d23.x * v[i].z; a.w += d23.y * v[i].w;
      *reinterpret_cast<float4*>(mine + D + c) = a;
      g[i] = make_float4(d01.x * (1.f + s4.x), d01.y * (1.f + s4.y), d23.x * (1.f + s4.z),
                         d23.y * (1.f + s4.w));
      sg += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      sgn += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
    }
    const float mg = warp_sum_b(sg) * inv_d;
    const float mgn = warp_sum_b(sgn) * inv_d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + 32 * i) * 4;
      float4 o = make_float4(rstd * (g[i].x - mg - v[i].x * mgn), rstd * (g[i].y - mg - v[i].y * mgn),
                             rstd * (g[i].z - mg - v[i].z * mgn), rstd * (g[i].w - mg - v[i].w * mgn));
      if (accumulate) { o.x += old[i].x; o.y += old[i].y; o.z += old[i].z; o.w += old[i].w; }
      *reinterpret_cast<float4*>(drow + c) = o;
      if (kGate) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gt + c));
        const float2 y01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yv[i].x));
        const float2 y23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yv[i].y));
        const float4 gd = make_float4(g4.x * o.x, g4.y * o.y, g4.z * o.z, g4.w * o.w);
        uint2 pk;
        pk.x = pack_bf16(gd.x, gd.y);
        pk.y = pack_bf16(gd.z, gd.w);
        *reinterpret_cast<uint2*>(dy + row * D + c) = pk;
        float4 a = *reinterpret_cast<float4*>(mine + 2 * D + c);  // dgate
        a.x += o.x * y01.x; a.y += o.y * y01.y; a.z += o.z * y23.x; a.w += o.w * y23.y;
        *reinterpret_cast<float4*>(mine + 2 * D + c) = a;
        a = *reinterpret_cast<float4*>(mine + 3 * D + c);  // dbias
        a.x += gd.x; a.y += gd.y; a.z += gd.z; a.w += gd.w;
        *reinterpret_cast<float4*>(mine + 3 * D + c) = a;
      }
    }
  }
  __syncthreads();
  float* dst[4] = {dshift + static_cast<int64_t>(b) * mod_ld, dscale + static_cast<int64_t>(b) * mod_ld,
                   kGate ? dgate + static_cast<int64_t>(b) * mod_ld : nullptr, kGate ? dbias : nullptr};
#pragma unroll
  for (int a = 0; a < kArr; ++a) {
    if (dst[a] == nullptr) continue;
    for (int c = threadIdx.x; c < D; c += kLgWarps * 32) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < kLgWarps; ++w) sum += s_strip[(w * kArr + a) * D + c];
      atomicAdd(dst[a] + c, sum);
    }
  }
}

// ------------------------------------------------------------------- final layer backward
// forward: out[b, o, t] = sum_d hmod[row, d] * W[o, d] + bias[o], hmod = LN(x)*(1+scale)+shift
// (x here already includes the last gated residual).  Given dout fp32 [B,4,T]:
//   dW[o, :] += dout * hmod;  dbias[o] += dout;  dhmod = sum_o dout[o] * W[o, :]  -> LN-modulate
//   backward -> dx (written), dshift, dscale.
template <int NV>
__global__ void __launch_bounds__(256)
final_layer_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dout,
                       const float* __restrict__ shift, const float* __restrict__ scale,
                       float* __restrict__ dshift, float* __restrict__ dscale, int64_t mod_ld, int T,
                       const float* __restrict__ w, float* __restrict__ dw, float* __restrict__ dbias,
                       float* __restrict__ dx) {
  constexpr int D = NV * 128;
  extern __shared__ float s_acc[];  // dshift[D] | dscale[D] | dW[4][D] | dbias[4]
  float* s_dshift = s_acc;
  float* s_dscale = s_acc + D;
  float* s_dw = s_acc + 2 * D;
  float* s_db = s_acc + 6 * D;
  for (int i = threadIdx.x; i < 6 * D + 4; i += 256) s_acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (t < T) {
    const int64_t row = static_cast<int64_t>(b) * T + t;
    const float* xrow = x + row * D;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = *reinterpret_cast<const float4*>(xrow + (lane + 32 * i) * 4);
    constexpr float inv_d = 1.0f / D;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum_b(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + bb * bb) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum_b(q) * inv_d + kLnEpsB);
    float dout4[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) dout4[o] = dout[(static_cast<int64_t>(b) * 4 + o) * T + t];
    if (lane < 4) atomicAdd(&s_db[lane], dout4[lane]);
    const float* sc = scale + static_cast<int64_t>(b) * mod_ld;
    const float* sh = shift + static_cast<int64_t>(b) * mod_ld;
    float4 g[NV];
    float sg = 0.f, sgn = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + 32 * i) * 4;
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(sc + c));
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(sh + c));
      v[i].x = (v[i].x - mean) * rstd; v[i].y = (v[i].y - mean) * rstd;
      v[i].z = (v[i].z - mean) * rstd; v[i].w = (v[i].w - mean) * rstd;
      const float4 hm = make_float4(fmaf(v[i].x, 1.f + s4.x, t4.x), fmaf(v[i].y, 1.f + s4.y, t4.y),
                                    fmaf(v[i].z, 1.f + s4.z, t4.z), fmaf(v[i].w, 1.f + s4.w, t4.w));
      float4 dh = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + o * D + c));
        dh.x = fmaf(dout4[o], w4.x, dh.x); dh.y = fmaf(dout4[o], w4.y, dh.y);
        dh.z = fmaf(dout4[o], w4.z, dh.z); dh.w = fmaf(dout4[o], w4.w, dh.w);
        atomicAdd(&s_dw[o * D + c + 0], dout4[o] * hm.x); atomicAdd(&s_dw[o * D + c + 1], dout4[o] * hm.y);
        atomicAdd(&s_dw[o * D + c + 2], dout4[o] * hm.z); atomicAdd(&s_dw[o * D + c + 3], dout4[o] * hm.w);
      }
      atomicAdd(&s_dshift[c + 0], dh.x); atomicAdd(&s_dshift[c + 1], dh.y);
      atomicAdd(&s_dshift[c + 2], dh.z); atomicAdd(&s_dshift[c + 3], dh.w);
      atomicAdd(&s_dscale[c + 0], dh.x * v[i].x); atomicAdd(&s_dscale[c + 1], dh.y * v[i].y);
      atomicAdd(&s_dscale[c + 2], dh.z * v[i].z); atomicAdd(&s_dscale[c + 3], dh.w * v[i].w);
      g[i] = make_float4(dh.x * (1.f + s4.x), dh.y * (1.f + s4.y), dh.z * (1.f + s4.z), dh.w * (1.f + s4.w));
      sg += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      sgn += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
    }
    const float mg = warp_sum_b(sg) * inv_d;
    const float mgn = warp_sum_b(sgn) * inv_d;
    float* drow = dx + row * D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + 32 * i) * 4;
      *reinterpret_cast<float4*>(drow + c) =
          make_float4(rstd * (g[i].x - mg - v[i].x * mgn), rstd * (g[i].y - mg - v[i].y * mgn),
                      rstd * (g[i].z - mg - v[i].z * mgn), rstd * (g[i].w - mg - v[i].w * mgn));
    }
  }
  __syncthreads();
  float* gs = dshift + static_cast<int64_t>(b) * mod_ld;
  float* gc = dscale + static_cast<int64_t>(b) * mod_ld;
  for (int i = threadIdx.x; i < D; i += 256) {
    atomicAdd(gs + i, s_dshift[i]);
    atomicAdd(gc + i, s_dscale[i]);
  }
  for (int i = threadIdx.x; i < 4 * D; i += 256) atomicAdd(dw + i, s_dw[i]);
  if (threadIdx.x < 4) atomicAdd(dbias + threadIdx.x, s_db[threadIdx.x]);
}

// -------------------------------------------------------------- SiLU backward, label scatter
// cond = a[ia[r]] + table[y[r]];  ds given (bf16 or fp32 as float) for s = SiLU(cond):
//   dcond[r] = ds[r] * silu'(cond[r]);  (dtable[y[r]] += dcond[r] when dtable != nullptr)
__global__ void __launch_bounds__(256)
silu_bwd_kernel(const float* __restrict__ a, const float* __restrict__ table,
                const int64_t* __restrict__ y, const float* __restrict__ ds, int64_t rows, int D,
                float* __restrict__ dcond, float* __restrict__ dtable) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= rows * D) return;
  const int64_t r = idx / D;
  const int d = static_cast<int>(idx - r * D);
  float v = a[idx];
  if (table) v += table[y[r] * D + d];
  const float sg = 1.0f / (1.0f + expf(-v));
  const float g = ds[idx] * (sg * (1.0f + v * (1.0f - sg)));
  dcond[idx] = g;
  if (dtable) atomicAdd(dtable + y[r] * D + d, g);
}

template <template <int> class L, typename... Args>
static int dispatch_nv1(int D, Args... args) {
  if (D % 128 != 0 || D < 128 || D > 1536)
    return set_error(-1, "hidden size must be a multiple of 128 in [128, 1536]");
#define OSUDIT_NV1(N) case N: return L<N>::run(args...);
  switch (D / 128) {
    OSUDIT_NV1(1) OSUDIT_NV1(2) OSUDIT_NV1(3) OSUDIT_NV1(4) OSUDIT_NV1(5) OSUDIT_NV1(6)
    OSUDIT_NV1(7) OSUDIT_NV1(8) OSUDIT_NV1(9) OSUDIT_NV1(10) OSUDIT_NV1(11) OSUDIT_NV1(12)
  }
#undef OSUDIT_NV1
  return set_error(-1, "unreachable");
}

template <int NV>
struct LnBwdLauncher {
  static int run(const float* x, const __nv_bfloat16* dh, const float* scale, float* dshift,
                 float* dscale, int64_t mod_ld, int B, int T, float* dx, int accumulate, cudaStream_t st) {
    dim3 grid((T + 7) / 8, B);
    ln_modulate_bwd_kernel<NV><<<grid, 256, 0, st>>>(x, dh, scale, dshift, dscale, mod_ld, T, dx, accumulate);
    OSUDIT_CHECK_LAUNCH();
    return 0;
  }
};

template <int NV>
struct LnGateBwdLauncher {
  static int run(const float* x, const __nv_bfloat16* dh, const float* scale, float* dshift, float* dscale,
                 int64_t mod_ld, int B, int T, float* dx, int accumulate, const __nv_bfloat16* y,
                 const float* gate, float* dgate, __nv_bfloat16* dy, float* dbias, cudaStream_t st) {
    const bool with_gate = y != nullptr;
    const int smem = kLgWarps * (with_gate ? 4 : 2) * NV * 128 * static_cast<int>(sizeof(float));
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(ln_gate_bwd_kernel<NV, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, kLgWarps * 4 * NV * 128 * 4);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(ln_gate_bwd_kernel<NV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kLgWarps * 2 * NV * 128 * 4);
      if (e != cudaSuccess) return set_error(-5, cudaGetErrorString(e));
      configured = true;
    }
    dim3 grid((T + kLgWarps * kLgRows - 1) / (kLgWarps * kLgRows), B);
    if (with_gate)
      ln_gate_bwd_kernel<NV, true><<<grid, kLgWarps * 32, smem, st>>>(x, dh, scale, dshift, dscale, mod_ld, T, dx,
                                                                     accumulate, y, gate, dgate, dy, dbias);
    else
      ln_gate_bwd_kernel<NV, false><<<grid, kLgWarps * 32, smem, st>>>(x, dh, scale, dshift, dscale, mod_ld, T, dx,
                                                                      accumulate, nullptr, nullptr, nullptr, nullptr,
                                                                      nullptr);
    OSUDIT_CHECK_LAUNCH();
    return 0;
  }
};

template <int NV>
struct FinalBwdLauncher {
  static int run(const float* x, const float* dout, const float* shift, const float* scale,
                 float* dshift, float* dscale, int64_t mod_ld, int B, int T, const float* w, float* dw,
                 float* dbias, float* dx, cudaStream_t st) {
    constexpr int smem = (6 * NV * 128 + 4) * sizeof(float);
    dim3 grid((T + 7) / 8, B);
    final_layer_bwd_kernel<NV><<<grid, 256, smem, st>>>(x, dout, shift, scale, dshift, dscale, mod_ld, T, w,
                                                       dw, dbias, dx);
    OSUDIT_CHECK_LAUNCH();
    return 0;
  }
};

}  // namespace osudit

using namespace osudit;

extern "C" int osudit_transpose_bf16(const void* in, void* out, int64_t rows, int64_t cols, int64_t out_ld, int in_is_f32,
                                     void* stream) {
  if (rows <= 0 || cols <= 0 || out_ld < rows) return set_error(-1, "transpose: bad shape");
  dim3 grid(static_cast<unsigned>((cols + 63) / 64), static_cast<unsigned>((rows + 63) / 64));
  if (grid.y > 65535) return set_error(-1, "transpose: too many rows for one launch");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (in_is_f32)
    transpose_f32_to_bf16_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(in),
                                                      static_cast<__nv_bfloat16*>(out), rows, cols, out_ld);
  else
    transpose_bf16_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(in),
                                               static_cast<__nv_bfloat16*>(out), rows, cols, out_ld);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

extern "C" int osudit_repack_weights(const void* segments, int nseg, int total_tiles, void* stream) {
  if (segments == nullptr || nseg <= 0 || total_tiles <= 0) return set_error(-1, "repack_weights: bad arguments");
  repack_weights_kernel<<<total_tiles, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const RepackSeg*>(segments), nseg);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

extern "C" int osudit_gelu(const void* pre, const void* dy, void* out, int64_t n, int backward,
                           void* stream) {
  if (n <= 0 || (n % 8) != 0) return set_error(-1, "gelu: element count must be a positive multiple of 8");
  if (backward && dy == nullptr) return set_error(-1, "gelu: backward needs dy");
  gelu_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(pre), static_cast<const __nv_bfloat16*>(dy),
      static_cast<__nv_bfloat16*>(out), n / 8, backward);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

extern "C" int osudit_gelu_bwd(const void* pre, const void* dy, void* out, int64_t rows, int N, float* dbias,
                               void* stream) {
  if (rows <= 0 || N <= 0 || (N % 8) != 0) return set_error(-1, "gelu_bwd: need rows > 0 and N a multiple of 8");
  dim3 grid((N / 8 + 255) / 256, static_cast<unsigned>((rows + kGeluRows - 1) / kGeluRows));
  gelu_bwd_colsum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(pre), static_cast<const __nv_bfloat16*>(dy),
      static_cast<__nv_bfloat16*>(out), dbias, rows, N);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

extern "C" int osudit_colsum(const void* in, int in_is_f32, int64_t rows, int N, float* out, void* stream) {
  if (rows <= 0 || N <= 0) return set_error(-1, "colsum: bad shape");
  const int rpc = 256;
  dim3 grid((N + 255) / 256, static_cast<unsigned>((rows + rpc - 1) / rpc));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (in_is_f32) colsum_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(in), out, rows, N, rpc);
  else colsum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(in), out, rows, N, rpc);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

extern "C" int osudit_gate_residual_bwd(const float* dx, const void* y, const float* gate, float* dgate,
                                        int64_t mod_ld, int B, int T, int D, void* dy, float* dbias,
                                        void* stream) {
  if (B <= 0 || T <= 0 || D % 4 != 0 || D / 4 > 384) return set_error(-1, "gate_residual_bwd: bad shape");
  dim3 grid((T + 31) / 32, B);
  gate_residual_bwd_kernel<<<grid, 384, 0, static_cast<cudaStream_t>(stream)>>>(
      dx, static_cast<const __nv_bfloat16*>(y), gate, dgate, mod_ld, T, D, static_cast<__nv_bfloat16*>(dy),
      dbias);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

extern "C" int osudit_ln_modulate_bwd(const float* x, const void* dh, const float* scale, float* dshift,
                                      float* dscale, int64_t mod_ld, int B, int T, int D, float* dx,
                                      int accumulate, void* stream) {
  if (B <= 0 || T <= 0 || B > 65535) return set_error(-1, "ln_modulate_bwd: bad shape");
  return dispatch_nv1<LnBwdLauncher>(D, x, static_cast<const __nv_bfloat16*>(dh), scale, dshift, dscale,
                                     mod_ld, B, T, dx, accumulate, static_cast<cudaStream_t>(stream));
}

extern "C" int osudit_ln_gate_bwd(const float* x, const void* dh, const float* scale, float* dshift,
                                  float* dscale, int64_t mod_ld, int B, int T, int D, float* dx, int accumulate,
                                  const void* y, const float* gate, float* dgate, void* dy, float* dbias,
                                  void* stream) {
  if (B <= 0 || T <= 0 || B > 65535) return set_error(-1, "ln_gate_bwd: bad shape");
  if (y != nullptr && (gate == nullptr || dgate == nullptr || dy == nullptr))
    return set_error(-1, "ln_gate_bwd: y needs gate, dgate and dy");
  return dispatch_nv1<LnGateBwdLauncher>(D, x, static_cast<const __nv_bfloat16*>(dh), scale, dshift, dscale,
                                         mod_ld, B, T, dx, accumulate, static_cast<const __nv_bfloat16*>(y), gate,
                                         dgate, static_cast<__nv_bfloat16*>(dy), dbias,
                                         static_cast<cudaStream_t>(stream));
}

extern "C" int osudit_final_layer_bwd(const float* x, const float* dout, const float* shift,
                                      const float* scale, float* dshift, float* dscale, int64_t mod_ld,
                                      int B, int T, int D, const float* w, float* dw, float* dbias,
                                      float* dx, void* stream) {
  if (B <= 0 || T <= 0 || B > 65535) return set_error(-1, "final_layer_bwd: bad shape");
  return dispatch_nv1<FinalBwdLauncher>(D, x, dout, shift, scale, dshift, dscale, mod_ld, B, T, w, dw, dbias,
                                        dx, static_cast<cudaStream_t>(stream));
}

extern "C" int osudit_silu_bwd(const float* a, const float* table, const int64_t* y, const float* ds,
                               int64_t rows, int D, float* dcond, float* dtable, void* stream) {
  if (rows <= 0 || D <= 0) return set_error(-1, "silu_bwd: bad shape");
  if ((table == nullptr) != (y == nullptr)) return set_error(-1, "silu_bwd: table and y go together");
  const int64_t n = rows * D;
  silu_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      a, table, y, ds, rows, D, dcond, dtable);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}
