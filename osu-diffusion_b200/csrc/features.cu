// Device-side beatmap feature builder: from the raw hit-object sequence to the model's
// conditioning inputs, so that c (B, 144, T) — 151 MB per config-2 batch — is never built on the host or
// moved over PCIe.
//
// Reference: data_loading.py:146-151 (calc_distances: distance to the previous object, the first one measured
// from the playfield centre (256, 192)), :172-187 (x = pos / (512, 384); c = [timestep_embedding(dist, 128)^T ;
// one-hot type rows]), :195-203 and sample.py:64-65 (o = time - time[0] (+ a random shift when training)),
// positional_embedding.py:29-49 ([cos | sin] of dist * freqs, freqs passed in as computed on the host).
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"

namespace osudit {

// seq [B, R, T] (rows: x px, y px, time ms, R-3 type rows) -> x [B, 2, T] (optional), o [B, T], c [B, 128+R-3, T].
__global__ void __launch_bounds__(128)
beatmap_features_kernel(const float* __restrict__ seq, int R, int T, const float* __restrict__ freqs,
                        const float* __restrict__ o_shift, float* __restrict__ x, float* __restrict__ o,
                        float* __restrict__ c) {
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t >= T) return;
  const int b = blockIdx.y;
  const float* s = seq + static_cast<int64_t>(b) * R * T;
  const float px = s[t], py = s[T + t];
  const float qx = t > 0 ? s[t - 1] : 256.0f;
  const float qy = t > 0 ? s[T + t - 1] : 192.0f;
  const float dx = px - qx, dy = py - qy;
  const float d = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
  const int E = 128 + R - 3;
  float* cb = c + static_cast<int64_t>(b) * E * T + t;
#pragma unroll 4
  for (int k = 0; k < 64; ++k) {
    float sn, cs;
    sincosf(__fmul_rn(d, __ldg(freqs + k)), &sn, &cs);
    cb[static_cast<int64_t>(k) * T] = cs;
    cb[static_cast<int64_t>(64 + k) * T] = sn;
  }
  for (int j = 3; j < R; ++j) cb[static_cast<int64_t>(125 + j) * T] = s[static_cast<int64_t>(j) * T + t];
  const float shift = o_shift != nullptr ? o_shift[b] : 0.0f;
  o[static_cast<int64_t>(b) * T + t] = __fadd_rn(__fsub_rn(s[2 * T + t], s[2 * T]), shift);
  if (x != nullptr) {
    x[(static_cast<int64_t>(b) * 2) * T + t] = __fdiv_rn(px, 512.0f);
    x[(static_cast<int64_t>(b) * 2 + 1) * T + t] = __fdiv_rn(py, 384.0f);
  }
}

}  // namespace osudit

using namespace osudit;

extern "C" int osudit_beatmap_features(const float* seq, int B, int R, int T, const float* freqs64,
                                       const float* o_shift, float* x, float* o, float* c, void* stream) {
  if (B <= 0 || T <= 0 || R < 3) return set_error(-1, "beatmap_features: bad shape");
  if (B > 65535) return set_error(-1, "beatmap_features: batch too large for one launch");
  dim3 grid((T + 127) / 128, B);
  beatmap_features_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(seq, R, T, freqs64, o_shift, x, o, c);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}
