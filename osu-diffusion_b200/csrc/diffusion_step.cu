// Fused classifier-free-guidance combine + one reverse-diffusion update (and q_sample).
//
// Reference: DiT.forward_with_cfg's combine (models.py:338-343), GaussianDiffusion.p_mean_variance
// for EPSILON + LEARNED_RANGE (gaussian_diffusion.py:312-324,341-358,371-376), q_posterior mean
// (:249-258), p_sample (:454-466), _extract_into_tensor (:951-963).  The reference spends ~25
// elementwise launches and 7 blocking H2D coefficient copies per step here; this is one launch
// reading a device-resident coefficient table (float32(table_f64[i]) exactly as `.float()` there).
//
// Per element (row b, channel ch in {0,1}, datapoint j):
//   eps  = cfg ? u + s (c - u) : out[b, ch]          c = out[b mod n, ch], u = out[b mod n + n, ch]
//   v    = out[b, 2 + ch]                            (variance channel is never guided)
//   lv   = f log_beta + (1 - f) post_logvar,  f = (v + 1) / 2
//   x0   = sqrt_recip x - sqrt_recipm1 eps   [-> denoised_fn on the host if any]  -> clamp(-1, 2)
//   mean = coef1 x0 + coef2 x ;   sample = mean + [t != 0] exp(lv / 2) noise
// Algorithmic traffic 56 B per (row, datapoint) with CFG.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"

namespace osudit {

// coefficient table row layout (6 floats per respaced timestep)
enum { C_LOG_BETA = 0, C_POST_LOGVAR, C_SQRT_RECIP, C_SQRT_RECIPM1, C_COEF1, C_COEF2, C_STRIDE };

__global__ void __launch_bounds__(256)
diffusion_step_kernel(const float* __restrict__ model_out, const float* __restrict__ x,
                      const float* __restrict__ noise, const float* __restrict__ x0_in,
                      const int64_t* __restrict__ t, const float* __restrict__ coef, int B, int T,
                      int cfg_half, float cfg_scale, int clip, int phase,
                      float* __restrict__ sample, float* __restrict__ pred_xstart,
                      float* __restrict__ mean_out, float* __restrict__ logvar_out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(B) * 2 * T;
  if (idx >= total) return;
  const int j = static_cast<int>(idx % T);
  const int ch = static_cast<int>((idx / T) % 2);
  const int b = static_cast<int>(idx / (2 * static_cast<int64_t>(T)));
  const int64_t ti = t[b];
  const float* cf = coef + ti * C_STRIDE;
  const float xv = x[idx];

  float x0;
  if (phase == 2) {
    x0 = x0_in[idx];  // the host applied denoised_fn to the phase-1 output
  } else {
    float eps;
    if (cfg_half > 0) {
      const int bc = b % cfg_half;
      const float c = model_out[(static_cast<int64_t>(bc) * 4 + ch) * T + j];
      const float u = model_out[(static_cast<int64_t>(bc + cfg_half) * 4 + ch) * T + j];
      eps = __fadd_rn(u, __fmul_rn(cfg_scale, __fsub_rn(c, u)));
    } else {
      eps = model_out[(static_cast<int64_t>(b) * 4 + ch) * T + j];
    }
    // x0 is a catastrophic cancellation at high noise levels (both products ~2e4): round each
    // product like the reference does, no FMA contraction.
    x0 = __fsub_rn(__fmul_rn(cf[C_SQRT_RECIP], xv), __fmul_rn(cf[C_SQRT_RECIPM1], eps));
    if (phase == 1) {  // stop before the callback; clamp happens after it (gaussian_diffusion.py:341-346)
      pred_xstart[idx] = x0;
      return;
    }
  }
  if (clip) x0 = fminf(fmaxf(x0, -1.0f), 2.0f);
  const float v = model_out[(static_cast<int64_t>(b) * 4 + 2 + ch) * T + j];
  const float frac = __fmul_rn(__fadd_rn(v, 1.0f), 0.5f);
  const float lv = __fadd_rn(__fmul_rn(frac, cf[C_LOG_BETA]),
                             __fmul_rn(__fsub_rn(1.0f, frac), cf[C_POST_LOGVAR]));
  const float mean = __fadd_rn(__fmul_rn(cf[C_COEF1], x0), __fmul_rn(cf[C_COEF2], xv));
  float out = mean;
  if (ti != 0 && sample) out = __fadd_rn(mean, __fmul_rn(expf(__fmul_rn(0.5f, lv)), noise[idx]));
  if (sample) sample[idx] = out;
  pred_xstart[idx] = x0;
  if (mean_out) mean_out[idx] = mean;
  if (logvar_out) logvar_out[idx] = lv;
}

__global__ void __launch_bounds__(256)
cfg_combine_kernel(const float* __restrict__ model_out, int B, int T, float cfg_scale,
                   float* __restrict__ out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(B) * 4 * T;
  if (idx >= total) return;
  const int j = static_cast<int>(idx % T);
  const int ch = static_cast<int>((idx / T) % 4);
  const int b = static_cast<int>(idx / (4 * static_cast<int64_t>(T)));
  const int half = B / 2;
  if (ch >= 2) {
    out[idx] = model_out[idx];
    return;
  }
  const int bc = b % half;
  const float c = model_out[(static_cast<int64_t>(bc) * 4 + ch) * T + j];
  const float u = model_out[(static_cast<int64_t>(bc + half) * 4 + ch) * T + j];
  out[idx] = __fadd_rn(u, __fmul_rn(cfg_scale, __fsub_rn(c, u)));
}

// x_t = sqrt(acp[t]) x0 + sqrt(1 - acp[t]) noise   (gaussian_diffusion.py:231-247)
__global__ void __launch_bounds__(256)
q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                const int64_t* __restrict__ t, const float* __restrict__ sqrt_acp,
                const float* __restrict__ sqrt_1m_acp, int64_t per_row, int64_t total,
                float* __restrict__ out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t ti = t[idx / per_row];
  out[idx] = __fadd_rn(__fmul_rn(sqrt_acp[ti], x0[idx]), __fmul_rn(sqrt_1m_acp[ti], noise[idx]));
}

}  // namespace osudit

using namespace osudit;

extern "C" int osudit_diffusion_step(const float* model_out, const float* x, const float* noise,
                                     const float* x0_in, const int64_t* t, const float* coef_table,
                                     int B, int T, int cfg_half, float cfg_scale, int clip_denoised,
                                     int phase, float* sample, float* pred_xstart, float* mean,
                                     float* log_variance, void* stream) {
  if (B <= 0 || T <= 0) return set_error(-1, "diffusion_step: bad shape");
  if (phase < 0 || phase > 2) return set_error(-1, "diffusion_step: phase must be 0, 1 or 2");
  if (phase == 2 && x0_in == nullptr) return set_error(-1, "diffusion_step: phase 2 needs x0_in");
  if (phase != 1 && sample != nullptr && noise == nullptr)
    return set_error(-1, "diffusion_step: noise is required when a sample is requested");
  if (cfg_half < 0 || (cfg_half > 0 && 2 * cfg_half != B))
    return set_error(-1, "diffusion_step: cfg_half must be 0 or B/2");
  const int64_t total = static_cast<int64_t>(B) * 2 * T;
  diffusion_step_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                          static_cast<cudaStream_t>(stream)>>>(
      model_out, x, noise, x0_in, t, coef_table, B, T, cfg_half, cfg_scale, clip_denoised, phase,
      sample, pred_xstart, mean, log_variance);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

extern "C" int osudit_cfg_combine(const float* model_out, int B, int T, float cfg_scale, float* out,
                                  void* stream) {
  if (B <= 0 || (B & 1) || T <= 0) return set_error(-1, "cfg_combine: batch must be even");
  const int64_t total = static_cast<int64_t>(B) * 4 * T;
  cfg_combine_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                       static_cast<cudaStream_t>(stream)>>>(model_out, B, T, cfg_scale, out);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

extern "C" int osudit_q_sample(const float* x0, const float* noise, const int64_t* t,
                               const float* sqrt_acp, const float* sqrt_1m_acp, int B,
                               int64_t per_row, float* out, void* stream) {
  if (B <= 0 || per_row <= 0) return set_error(-1, "q_sample: bad shape");
  const int64_t total = static_cast<int64_t>(B) * per_row;
  q_sample_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                    static_cast<cudaStream_t>(stream)>>>(x0, noise, t, sqrt_acp, sqrt_1m_acp, per_row,
                                                         total, out);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}
