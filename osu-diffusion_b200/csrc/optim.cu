// Multi-tensor AdamW + EMA (+ gradient unscale, + skip on non-finite gradients) in one launch.
//
// Replaces, for a maintainer who opts in, the tail of the reference training step:
// `scaler.step(opt)` with torch.optim.AdamW(lr, weight_decay) (train.py:154,258-259) and `update_ema(ema,
// model.module)` (train.py:36-45,261: ema = ema * decay + p * (1 - decay) over every parameter, ~2 launches
// per tensor).  HBM-bound: 40 B per parameter per step (read p, g, m, v, ema; write p, m, v, ema), SURVEY §8(f)1.
//
// Arithmetic follows torch.optim.AdamW (non-amsgrad): p *= 1 - lr*wd; m = lerp(m, g, 1-b1);
// v = b2*v + (1-b2) g^2; p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps).  The step count t lives on
// the device so that a step skipped because of inf/nan gradients (GradScaler semantics) does not advance it.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"

namespace osudit {

struct OptSeg {  // mirrors OsuditOptSeg in include/osudit.h
  float* p;
  const float* g;
  float* m;
  float* v;
  float* ema;  // may be null
  long long n;
};

constexpr int kOptChunk = 4096;  // elements per CTA

__global__ void opt_advance_step_kernel(float* step, const float* found_inf) {
  if (found_inf == nullptr || *found_inf == 0.f) *step += 1.f;
}

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, float* ema, float lr_wd,
                                          float b1, float b2, float step_size, float inv_sqrt_bc2, float eps,
                                          float ema_decay) {
  p *= lr_wd;
  m = m + (1.0f - b1) * (g - m);
  v = b2 * v + (1.0f - b2) * g * g;
  const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
  p -= step_size * (m / denom);
  if (ema != nullptr) *ema = *ema * ema_decay + p * (1.0f - ema_decay);
}

__global__ void __launch_bounds__(256)
adamw_ema_kernel(const OptSeg* __restrict__ segs, const int2* __restrict__ chunks, float lr, float b1, float b2,
                 float eps, float wd, float ema_decay, const float* __restrict__ step,
                 const float* __restrict__ grad_scale, const float* __restrict__ found_inf) {
  if (found_inf != nullptr && *found_inf != 0.f) return;  // GradScaler: skip the whole update
  __shared__ float s_coef[2];
  if (threadIdx.x == 0) {
    const double t = static_cast<double>(*step);
    s_coef[0] = static_cast<float>(static_cast<double>(lr) / (1.0 - pow(static_cast<double>(b1), t)));
    s_coef[1] = static_cast<float>(1.0 / sqrt(1.0 - pow(static_cast<double>(b2), t)));
  }
  __syncthreads();
  const float step_size = s_coef[0], inv_sqrt_bc2 = s_coef[1];
  const float inv_scale = grad_scale != nullptr ? 1.0f / *grad_scale : 1.0f;
  const float lr_wd = 1.0f - lr * wd;
  const int2 ch = chunks[blockIdx.x];
  const OptSeg s = segs[ch.x];
  const long long base = static_cast<long long>(ch.y) * kOptChunk;
  const long long end = base + kOptChunk < s.n ? base + kOptChunk : s.n;
  const bool vec = ((reinterpret_cast<uintptr_t>(s.p) | reinterpret_cast<uintptr_t>(s.g) |
                     reinterpret_cast<uintptr_t>(s.m) | reinterpret_cast<uintptr_t>(s.v) |
                     reinterpret_cast<uintptr_t>(s.ema)) & 15) == 0;
  for (long long i = base + threadIdx.x * 4; i < end; i += 1024) {
    if (vec && i + 3 < end) {
      float4 p = *reinterpret_cast<float4*>(s.p + i);
      float4 g = *reinterpret_cast<const float4*>(s.g + i);
      float4 m = *reinterpret_cast<float4*>(s.m + i);
      float4 v = *reinterpret_cast<float4*>(s.v + i);
      float4 e = s.ema != nullptr ? *reinterpret_cast<float4*>(s.ema + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      float* ep = s.ema != nullptr ? &e.x : nullptr;
      adamw_one(p.x, g.x * inv_scale, m.x, v.x, ep, lr_wd, b1, b2, step_size, inv_sqrt_bc2, eps, ema_decay);
      adamw_one(p.y, g.y * inv_scale, m.y, v.y, ep ? ep + 1 : nullptr, lr_wd, b1, b2, step_size, inv_sqrt_bc2, eps,
                ema_decay);
      adamw_one(p.z, g.z * inv_scale, m.z, v.z, ep ? ep + 2 : nullptr, lr_wd, b1, b2, step_size, inv_sqrt_bc2, eps,
                ema_decay);
      adamw_one(p.w, g.w * inv_scale, m.w, v.w, ep ? ep + 3 : nullptr, lr_wd, b1, b2, step_size, inv_sqrt_bc2, eps,
                ema_decay);
      *reinterpret_cast<float4*>(s.p + i) = p;
      *reinterpret_cast<float4*>(s.m + i) = m;
      *reinterpret_cast<float4*>(s.v + i) = v;
      if (s.ema != nullptr) *reinterpret_cast<float4*>(s.ema + i) = e;
    } else {
      for (long long j = i; j < i + 4 && j < end; ++j) {
        float p = s.p[j], m = s.m[j], v = s.v[j];
        adamw_one(p, s.g[j] * inv_scale, m, v, s.ema != nullptr ? s.ema + j : nullptr, lr_wd, b1, b2, step_size,
                  inv_sqrt_bc2, eps, ema_decay);
        s.p[j] = p; s.m[j] = m; s.v[j] = v;
      }
    }
  }
}

}  // namespace osudit

using namespace osudit;

extern "C" int osudit_opt_chunk_elems(void) { return kOptChunk; }

extern "C" int osudit_adamw_ema_step(const void* segs, const int32_t* chunks, int nchunks, float lr, float beta1,
                                     float beta2, float eps, float weight_decay, float ema_decay, float* step,
                                     const float* grad_scale, const float* found_inf, void* stream) {
  if (segs == nullptr || chunks == nullptr || step == nullptr || nchunks <= 0)
    return set_error(-1, "adamw_ema_step: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  opt_advance_step_kernel<<<1, 1, 0, st>>>(step, found_inf);
  OSUDIT_CHECK_LAUNCH();
  adamw_ema_kernel<<<nchunks, 256, 0, st>>>(static_cast<const OptSeg*>(segs),
                                            reinterpret_cast<const int2*>(chunks), lr, beta1, beta2, eps,
                                            weight_decay, ema_decay, step, grad_scale, found_inf);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}
