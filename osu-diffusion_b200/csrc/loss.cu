// Training loss of the diffusion model, forward and backward in one launch.
//
// Reference: GaussianDiffusion.training_losses for EPSILON + LEARNED_RANGE with MSE or L1
// (gaussian_diffusion.py:785-874), _vb_terms_bpd (:735-783), q_posterior_mean_variance (:249-271),
// p_mean_variance with clip_denoised=False (:312-358), normal_kl / discretized_gaussian_log_likelihood
// / approx_standard_normal_cdf (diffusion_utils.py:9-43,63-89), mean_flat (:15-19).
//
//   main[b] = mean_{c,j} |noise - eps|   (L1)   or   (noise - eps)^2   (MSE)
//   vb[b]   = mean_{c,j} ( t==0 ? NLL : KL ) / ln 2, with the model mean built from DETACHED eps:
//             only the variance channels v receive a gradient from vb (gaussian_diffusion.py:833).
//   d(main + vb)[b] / d model_out[b]  is written next to the values, so the autograd node only has
//   to scale it by the incoming per-sample gradient.
// One CTA per batch element: block reduction, no atomics, deterministic.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"

namespace osudit {

enum { L_LOG_BETA = 0, L_POST_LOGVAR, L_SQRT_RECIP, L_SQRT_RECIPM1, L_COEF1, L_COEF2, L_STRIDE };

__device__ __forceinline__ void approx_cdf(float z, float& cdf, float& dcdf) {
  const float k = 0.7978845608028654f, a = 0.044715f;
  const float th = tanhf(k * (z + a * z * z * z));
  cdf = 0.5f * (1.0f + th);
  dcdf = 0.5f * (1.0f - th * th) * k * (1.0f + 3.0f * a * z * z);
}

__global__ void __launch_bounds__(256)
diffusion_loss_kernel(const float* __restrict__ model_out, const float* __restrict__ x0,
                      const float* __restrict__ x_t, const float* __restrict__ noise,
                      const int64_t* __restrict__ t, const float* __restrict__ coef, int T, int use_l1,
                      float* __restrict__ term_main, float* __restrict__ term_vb,
                      float* __restrict__ dmodel_out) {
  const int b = blockIdx.x;
  const int64_t ti = t[b];
  const float* cf = coef + ti * L_STRIDE;
  const float max_log = cf[L_LOG_BETA], min_log = cf[L_POST_LOGVAR];
  const float sr = cf[L_SQRT_RECIP], srm1 = cf[L_SQRT_RECIPM1], c1 = cf[L_COEF1], c2 = cf[L_COEF2];
  const float inv_n = 1.0f / (2.0f * T);
  const float inv_ln2 = 1.4426950408889634f;
  const float* mo = model_out + static_cast<int64_t>(b) * 4 * T;
  float* dmo = dmodel_out + static_cast<int64_t>(b) * 4 * T;
  float acc_main = 0.f, acc_vb = 0.f;
  for (int i = threadIdx.x; i < 2 * T; i += 256) {
    const int64_t off = static_cast<int64_t>(b) * 2 * T + i;
    const float eps = mo[i], v = mo[2 * T + i];
    const float xs = x0[off], xt = x_t[off], nz = noise[off];
    // ---- main term
    const float diff = nz - eps;
    float deps;
    if (use_l1) {
      acc_main += fabsf(diff);
      deps = diff > 0.f ? -1.0f : (diff < 0.f ? 1.0f : 0.f);
    } else {
      acc_main += diff * diff;
      deps = -2.0f * diff;
    }
    dmo[i] = deps * inv_n;
    // ---- variational-bound term (gradient to v only)
    const float frac = (v + 1.0f) * 0.5f;
    const float lv2 = frac * max_log + (1.0f - frac) * min_log;
    const float pred_x0 = sr * xt - srm1 * eps;
    const float m2 = c1 * pred_x0 + c2 * xt;
    float val, dval_dlv2;
    if (ti != 0) {
      const float m1 = c1 * xs + c2 * xt;
      const float lv1 = min_log;
      const float e1 = expf(lv1 - lv2);
      const float sq = (m1 - m2) * (m1 - m2) * expf(-lv2);
      val = 0.5f * (-1.0f + lv2 - lv1 + e1 + sq);
      dval_dlv2 = 0.5f * (1.0f - e1 - sq);
    } else {
      const float ls = 0.5f * lv2;
      const float inv = expf(-ls);
      const float d = xs - m2;
      const float plus = inv * (d + 1.0f / 255.0f), minus = inv * (d - 1.0f / 255.0f);
      float cp, dcp, cm, dcm;
      approx_cdf(plus, cp, dcp);
      approx_cdf(minus, cm, dcm);
      float lp, dlp_dls;
      if (xs < -0.999f) {
        lp = logf(fmaxf(cp, 1e-12f));
        dlp_dls = cp > 1e-12f ? dcp * (-plus) / cp : 0.f;
      } else if (xs > 0.999f) {
        const float om = 1.0f - cm;
        lp = logf(fmaxf(om, 1e-12f));
        dlp_dls = om > 1e-12f ? dcm * minus / om : 0.f;
      } else {
        const float delta = cp - cm;
        lp = logf(fmaxf(delta, 1e-12f));
        dlp_dls = delta > 1e-12f ? (dcp * (-plus) + dcm * minus) / delta : 0.f;
      }
      val = -lp;
      dval_dlv2 = -0.5f * dlp_dls;
    }
    acc_vb += val;
    dmo[2 * T + i] = dval_dlv2 * 0.5f * (max_log - min_log) * inv_n * inv_ln2;
  }
  __shared__ float s_a[8], s_b[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc_main += __shfl_xor_sync(0xffffffffu, acc_main, o);
    acc_vb += __shfl_xor_sync(0xffffffffu, acc_vb, o);
  }
  if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = acc_main; s_b[threadIdx.x >> 5] = acc_vb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, v = 0.f;
    for (int w = 0; w < 8; ++w) { a += s_a[w]; v += s_b[w]; }
    term_main[b] = a * inv_n;
    term_vb[b] = v * inv_n * inv_ln2;
  }
}

// out[b, :] = in[b, :] * g[b]
__global__ void __launch_bounds__(256)
scale_rows_kernel(const float* __restrict__ in, const float* __restrict__ g, int64_t per_row,
                  int64_t total, float* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < total) out[i] = in[i] * g[i / per_row];
}

}  // namespace osudit

using namespace osudit;

extern "C" int osudit_diffusion_loss(const float* model_out, const float* x0, const float* x_t,
                                     const float* noise, const int64_t* t, const float* coef_table, int B,
                                     int T, int use_l1, float* term_main, float* term_vb,
                                     float* dmodel_out, void* stream) {
  if (B <= 0 || T <= 0) return set_error(-1, "diffusion_loss: bad shape");
  diffusion_loss_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      model_out, x0, x_t, noise, t, coef_table, T, use_l1, term_main, term_vb, dmodel_out);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

extern "C" int osudit_scale_rows(const float* in, const float* g, int B, int64_t per_row, float* out,
                                 void* stream) {
  if (B <= 0 || per_row <= 0) return set_error(-1, "scale_rows: bad shape");
  const int64_t total = static_cast<int64_t>(B) * per_row;
  scale_rows_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                      static_cast<cudaStream_t>(stream)>>>(in, g, per_row, total, out);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}
