// "fp32 mode" of the denoiser (north star: predicted epsilon within 1e-5 relative L2 of the fp32
// reference).  Accuracy first, speed second: this is the mode a user selects to validate a checkpoint
// against the reference, not the one that is benchmarked.
//
// Every activation stays fp32 in HBM.  A GEMM operand is handed to the bf16 tensor-core kernel as a
// three-way split  v = hi + mid + lo  (each bf16, the residuals are exact in fp32, 24 mantissa bits in
// total) stored side by side as one row [hi(K) | mid(K) | lo(K)]; the weights as [hi hi hi mid mid lo].
// Three K-segments of osudit_gemm_bf16 then accumulate the six significant products
//   (hi+mid+lo)·Whi + (hi+mid)·Wmid + hi·Wlo
// in one fp32 TMEM accumulator.  The kernels here are the fp32 producers / consumers around those
// GEMMs: split (+GELU / SiLU / label-embedding add), residual + LayerNorm + modulate, the final layer,
// and an fp32 CUDA-core banded attention.  Reference lines as in ln_modulate.cu / attn_band.cu /
// embed.cu: models.py:12-13,112-119,151-175,192-196,164-170,320.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "common.h"

namespace osudit {

__device__ __forceinline__ void split3_store(__nv_bfloat16* __restrict__ row3, int K, int k, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(h);
  const __nv_bfloat16 m = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(m);
  row3[k] = h;
  row3[K + k] = m;
  row3[2 * K + k] = __float2bfloat16_rn(r2);
}

__device__ __forceinline__ float gelu_tanh_f32(float x) {  // nn.GELU(approximate="tanh"), models.py:138
  const float x3 = x * x * x;
  return 0.5f * x * (1.0f + tanhf(0.7978845608028654f * (x + 0.044715f * x3)));
}

__device__ __forceinline__ float silu_f32(float v) { return v / (1.0f + expf(-v)); }

// out3[r] = split3(act(in[r] (+ table[y[r]])));  act: 0 identity, 1 GELU(tanh), 2 SiLU.
__global__ void split3_kernel(const float* __restrict__ in, int64_t rows, int K, int act,
                              const float* __restrict__ table, const int64_t* __restrict__ y,
                              __nv_bfloat16* __restrict__ out3) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= rows * K) return;
  const int64_t r = idx / K;
  const int k = static_cast<int>(idx - r * K);
  float v = in[idx];
  if (table != nullptr) v = __fadd_rn(v, table[y[r] * K + k]);
  if (act == 1) v = gelu_tanh_f32(v);
  else if (act == 2) v = silu_f32(v);
  split3_store(out3 + r * 3 * K, K, k, v);
}

__device__ __forceinline__ float warp_sum_f32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per token row, any D % 4 == 0.  x += gate*branch (in place, fp32 branch), two-pass
// LayerNorm statistics, modulate with the reference's operation order (no contraction).
// kFinal: project to 4 channels instead of writing the split operand.
template <bool kFinal>
__global__ void __launch_bounds__(256)
ln_f32_kernel(float* __restrict__ x, const float* __restrict__ branch, const float* __restrict__ gate,
              const float* __restrict__ shift, const float* __restrict__ scale, int64_t mod_ld,
              int64_t rows, int T, int D, __nv_bfloat16* __restrict__ h3, const float* __restrict__ w,
              const float* __restrict__ bias, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int64_t b = row / T;
  float* xr = x + row * D;
  float s = 0.f;
  for (int c = lane * 4; c < D; c += 128) {
    float4 v = *reinterpret_cast<const float4*>(xr + c);
    if (branch != nullptr) {
      const float4 yv = *reinterpret_cast<const float4*>(branch + row * D + c);
      const float4 g = *reinterpret_cast<const float4*>(gate + b * mod_ld + c);
      v.x = __fadd_rn(v.x, __fmul_rn(g.x, yv.x));
      v.y = __fadd_rn(v.y, __fmul_rn(g.y, yv.y));
      v.z = __fadd_rn(v.z, __fmul_rn(g.z, yv.z));
      v.w = __fadd_rn(v.w, __fmul_rn(g.w, yv.w));
      *reinterpret_cast<float4*>(xr + c) = v;
    }
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = warp_sum_f32(s) / static_cast<float>(D);
  float q = 0.f;
  for (int c = lane * 4; c < D; c += 128) {  // re-reads this lane's own stores
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const float a0 = v.x - mean, a1 = v.y - mean, a2 = v.z - mean, a3 = v.w - mean;
    q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
  }
  const float rstd = 1.0f / sqrtf(warp_sum_f32(q) / static_cast<float>(D) + 1e-6f);
  const float* sh = shift + b * mod_ld;
  const float* sc = scale + b * mod_ld;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = lane * 4; c < D; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const float4 s4 = *reinterpret_cast<const float4*>(sc + c);
    const float4 t4 = *reinterpret_cast<const float4*>(sh + c);
    float hv[4];
    hv[0] = __fadd_rn(__fmul_rn(__fmul_rn(v.x - mean, rstd), __fadd_rn(1.0f, s4.x)), t4.x);
    hv[1] = __fadd_rn(__fmul_rn(__fmul_rn(v.y - mean, rstd), __fadd_rn(1.0f, s4.y)), t4.y);
    hv[2] = __fadd_rn(__fmul_rn(__fmul_rn(v.z - mean, rstd), __fadd_rn(1.0f, s4.z)), t4.z);
    hv[3] = __fadd_rn(__fmul_rn(__fmul_rn(v.w - mean, rstd), __fadd_rn(1.0f, s4.w)), t4.w);
    if (kFinal) {
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const float4 w4 = *reinterpret_cast<const float4*>(w + static_cast<int64_t>(o) * D + c);
        acc[o] = fmaf(hv[0], w4.x, fmaf(hv[1], w4.y, fmaf(hv[2], w4.z, fmaf(hv[3], w4.w, acc[o]))));
      }
    } else {
      __nv_bfloat16* hr = h3 + row * 3 * D;
#pragma unroll
      for (int j = 0; j < 4; ++j) split3_store(hr, D, c + j, hv[j]);
    }
  }
  if (kFinal) {
#pragma unroll
    for (int o = 0; o < 4; ++o) acc[o] = warp_sum_f32(acc[o]);
    if (lane < 4) {
      const float r = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
      const int64_t t = row - b * T;
      out[(b * 4 + lane) * T + t] = r + bias[lane];
    }
  }
}

// fp32 banded / full / generically masked attention on the CUDA cores: one thread per query row,
// 64 queries per CTA, K/V tiles of 64 keys staged in shared memory and read as warp broadcasts,
// online softmax with expf.  qkv [B*T, 3*H*HD] fp32 (rows [q | k | v], heads contiguous), out [B*T, H*HD].
template <int HD>
__global__ void __launch_bounds__(64)
attn_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out, int T, int H, int w_left,
                int w_right, const uint8_t* __restrict__ mask) {
  __shared__ __align__(16) float ks[64][HD];
  __shared__ __align__(16) float vs[64][HD];
  const int tid = threadIdx.x;
  const int q0 = blockIdx.x * 64;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int D = H * HD;
  const int64_t ld = 3 * static_cast<int64_t>(D);
  const int qi = q0 + tid;
  const bool valid = qi < T;
  const float* base = qkv + static_cast<int64_t>(b) * T * ld + h * HD;
  float q[HD], o[HD];
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    const float4 v = valid ? *reinterpret_cast<const float4*>(base + qi * ld + d) : make_float4(0, 0, 0, 0);
    q[d] = v.x; q[d + 1] = v.y; q[d + 2] = v.z; q[d + 3] = v.w;
    o[d] = o[d + 1] = o[d + 2] = o[d + 3] = 0.f;
  }
  const float inv_sqrt = 1.0f / sqrtf(static_cast<float>(HD));
  float m = -INFINITY, l = 0.f;
  const int klo = w_left < 0 ? 0 : max(0, q0 - w_left);
  const int khi = w_right < 0 ? T : min(T, q0 + 63 + w_right + 1);
  for (int kt = klo; kt < khi; kt += 64) {
    const int nk = min(64, khi - kt);
    __syncthreads();
    for (int idx = tid; idx < nk * (HD / 4); idx += 64) {
      const int j = idx / (HD / 4);
      const int c = (idx - j * (HD / 4)) * 4;
      const float* src = base + static_cast<int64_t>(kt + j) * ld + c;
      *reinterpret_cast<float4*>(&ks[j][c]) = *reinterpret_cast<const float4*>(src + D);
      *reinterpret_cast<float4*>(&vs[j][c]) = *reinterpret_cast<const float4*>(src + 2 * D);
    }
    __syncthreads();
    if (!valid) continue;
    for (int j = 0; j < nk; ++j) {
      const int kj = kt + j;
      if (w_left >= 0 && kj < qi - w_left) continue;
      if (w_right >= 0 && kj > qi + w_right) continue;
      if (mask != nullptr && mask[static_cast<int64_t>(qi) * T + kj]) continue;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 kv = *reinterpret_cast<const float4*>(&ks[j][d]);
        s0 = fmaf(q[d], kv.x, s0);
        s1 = fmaf(q[d + 1], kv.y, s1);
        s2 = fmaf(q[d + 2], kv.z, s2);
        s3 = fmaf(q[d + 3], kv.w, s3);
      }
      const float s = ((s0 + s1) + (s2 + s3)) * inv_sqrt;
      float p = 1.0f;
      if (s > m) {
        const float corr = expf(m - s);  // 0 on the first key (m = -inf)
        l *= corr;
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] *= corr;
        m = s;
      } else {
        p = expf(s - m);
      }
      l += p;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(&vs[j][d]);
        o[d] = fmaf(p, vv.x, o[d]);
        o[d + 1] = fmaf(p, vv.y, o[d + 1]);
        o[d + 2] = fmaf(p, vv.z, o[d + 2]);
        o[d + 3] = fmaf(p, vv.w, o[d + 3]);
      }
    }
  }
  if (!valid) return;
  float* dst = out + (static_cast<int64_t>(b) * T + qi) * D + h * HD;
#pragma unroll
  for (int d = 0; d < HD; d += 4)  // a row with no allowed key gives 0/0 = NaN, as softmax over all -inf does
    *reinterpret_cast<float4*>(dst + d) = make_float4(o[d] / l, o[d + 1] / l, o[d + 2] / l, o[d + 3] / l);
}

}  // namespace osudit

using namespace osudit;

extern "C" int osudit_split3_bf16(const float* in, int64_t rows, int K, int act, const float* table,
                                  const int64_t* y, void* out3, void* stream) {
  if (rows <= 0 || K <= 0) return set_error(-1, "split3: bad shape");
  if (act < 0 || act > 2) return set_error(-1, "split3: unknown activation");
  if ((table == nullptr) != (y == nullptr)) return set_error(-1, "split3: table and y must be given together");
  const int64_t n = rows * K;
  split3_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, rows, K, act, table, y, static_cast<__nv_bfloat16*>(out3));
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

extern "C" int osudit_ln_modulate_f32(float* x, const float* branch, const float* gate, const float* shift,
                                      const float* scale, int64_t mod_ld, int64_t rows, int T, int D,
                                      void* h3, void* stream) {
  if (rows <= 0 || T <= 0 || D <= 0 || (D % 4) != 0) return set_error(-1, "ln_modulate_f32: bad shape");
  if ((branch == nullptr) != (gate == nullptr))
    return set_error(-1, "ln_modulate_f32: branch and gate must be given together");
  ln_f32_kernel<false><<<static_cast<unsigned>((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, branch, gate, shift, scale, mod_ld, rows, T, D, static_cast<__nv_bfloat16*>(h3), nullptr, nullptr,
      nullptr);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

extern "C" int osudit_final_layer_f32(float* x, const float* branch, const float* gate, const float* shift,
                                      const float* scale, int64_t mod_ld, int64_t rows, int T, int D,
                                      const float* w, const float* bias, int out_channels, float* out,
                                      void* stream) {
  if (rows <= 0 || T <= 0 || D <= 0 || (D % 4) != 0) return set_error(-1, "final_layer_f32: bad shape");
  if (out_channels != 4) return set_error(-1, "final_layer_f32: out_channels must be 4");
  if ((branch == nullptr) != (gate == nullptr))
    return set_error(-1, "final_layer_f32: branch and gate must be given together");
  ln_f32_kernel<true><<<static_cast<unsigned>((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, branch, gate, shift, scale, mod_ld, rows, T, D, nullptr, w, bias, out);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

extern "C" int osudit_attn_band_f32(const float* qkv, float* out, int B, int T, int H, int head_dim,
                                    int w_left, int w_right, const uint8_t* mask, void* stream) {
  if (B <= 0 || T <= 0 || H <= 0) return set_error(-1, "attn_band_f32: bad shape");
  if (B > 65535 || H > 65535) return set_error(-1, "attn_band_f32: batch / heads too large for one launch");
  if ((w_left < 0) != (w_right < 0)) return set_error(-1, "attn_band_f32: give both window sides or neither");
  dim3 grid((T + 63) / 64, H, B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (head_dim == 64) attn_f32_kernel<64><<<grid, 64, 0, st>>>(qkv, out, T, H, w_left, w_right, mask);
  else if (head_dim == 72) attn_f32_kernel<72><<<grid, 64, 0, st>>>(qkv, out, T, H, w_left, w_right, mask);
  else return set_error(-1, "attn_band_f32: head_dim must be 64 or 72");
  OSUDIT_CHECK_LAUNCH();
  return 0;
}
