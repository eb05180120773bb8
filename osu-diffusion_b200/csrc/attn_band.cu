// Banded multi-head self-attention over packed QKV (flash style: K/V staged in shared memory,
// online softmax in fp32, only key tiles that intersect the band are visited).
//
// Reference: nn.MultiheadAttention inside DiTBlock (models.py:130-135,164-170) with the boolean
// band mask sample.py:81-84 builds (query j may attend key i iff -(W-1) <= i-j <= W; closed form in
// SURVEY.md F3) or no mask (training windows).  The reference computes dense TxT scores and masks
// them; here cost is O(T * band).  A generic (T,T) byte mask is honoured per element as well.
//
// Layout: qkv [B*T, 3*D] bf16 (rows = tokens, [q | k | v], heads are contiguous HD-slices — the
// packed in_proj layout), out [B*T, D] bf16.  One CTA = 64 queries of one (batch, head); 4 warps x
// 16 query rows; mma.sync m16n8k16 bf16 with fp32 accumulation; cp.async double-buffered K/V tiles
// in XOR-swizzled shared memory read through ldmatrix.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/osudit.h"
#include "common.h"
#include "ptx.cuh"

namespace osudit {

constexpr int kBQ = 64;   // queries per CTA
constexpr int kBKV = 64;  // keys per tile

// Shared-memory tile geometry per head_dim.  hd=64: 128-byte rows, XOR swizzle of the 16-byte chunk
// index with (row & 7).  hd=72 (DiT-XL, models.py:410-412): rows padded to 80 columns (the extra
// 16-byte chunk is zero so Q K^T can run 5 k-steps of 16) at a 176-byte pitch, which makes the
// 8-row ldmatrix fetches bank-conflict free without a swizzle.
template <int HD>
struct Geo {
  static constexpr int kChunks = HD / 8;                  // 16-byte chunks of real data per row
  static constexpr int kKSteps = (HD + 15) / 16;          // k-steps of Q K^T
  static constexpr int kDBlocks = 2 * ((HD / 8 + 1) / 2); // 8-wide output blocks (even, incl. padding)
  static constexpr int kPitch = HD == 64 ? 128 : 176;
  static constexpr int kTileBytes = kBKV * kPitch;
  static constexpr bool kPadded = (HD % 16) != 0;
  __device__ static __forceinline__ uint32_t off(int r, int ch) {
    return HD == 64 ? r * 128 + ((ch ^ (r & 7)) << 4) : r * kPitch + (ch << 4);
  }
};

template <int HD>
__device__ __forceinline__ void load_tile_async(uint8_t* s, const __nv_bfloat16* g, int64_t ld,
                                                int row0, int T, int tid) {
  using G = Geo<HD>;
  for (int idx = tid; idx < kBKV * G::kChunks; idx += 128) {
    const int r = idx / G::kChunks;
    const int ch = idx - r * G::kChunks;
    const int row = row0 + r;
    const bool valid = row >= 0 && row < T;
    const __nv_bfloat16* src = g + static_cast<int64_t>(valid ? row : 0) * ld + ch * 8;
    cp_async_16(s + G::off(r, ch), src, valid);
  }
}

template <int HD>
__global__ void __launch_bounds__(128)
attn_band_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int T, int H,
                 int wl, int wr, const uint8_t* __restrict__ mask, float scale_log2,
                 float* __restrict__ lse) {
  using G = Geo<HD>;
  extern __shared__ __align__(128) uint8_t smem_attn[];
  uint8_t* sQ = smem_attn;
  uint8_t* sK[2] = {sQ + G::kTileBytes, sQ + 2 * G::kTileBytes};
  uint8_t* sV[2] = {sQ + 3 * G::kTileBytes, sQ + 4 * G::kTileBytes};

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int q0 = blockIdx.x * kBQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int D = H * HD;
  const int64_t ld = 3 * static_cast<int64_t>(D);
  const __nv_bfloat16* gq = qkv + static_cast<int64_t>(b) * T * ld + h * HD;
  const __nv_bfloat16* gk = gq + D;
  const __nv_bfloat16* gv = gq + 2 * D;

  if (G::kPadded) {  // zero the padding chunk of every row once; cp.async never writes it
    for (int idx = tid; idx < 5 * kBKV; idx += 128)
      *reinterpret_cast<uint4*>(smem_attn + (idx / kBKV) * G::kTileBytes + G::off(idx % kBKV, G::kChunks)) =
          make_uint4(0, 0, 0, 0);
  }

  const int k_first = max(0, q0 - wl);
  const int k_last = min(T - 1, min(q0 + kBQ - 1, T - 1) + wr);
  const int kt_lo = k_first / kBKV;
  const int kt_hi = k_last / kBKV;

  load_tile_async<HD>(sQ, gq, ld, q0, T, tid);
  load_tile_async<HD>(sK[0], gk, ld, kt_lo * kBKV, T, tid);
  load_tile_async<HD>(sV[0], gv, ld, kt_lo * kBKV, T, tid);
  cp_async_commit();

  float o_acc[G::kDBlocks][4];
#pragma unroll
  for (int i = 0; i < G::kDBlocks; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  uint32_t qf[G::kKSteps][4];

  const int qrow[2] = {q0 + warp * 16 + (lane >> 2), q0 + warp * 16 + (lane >> 2) + 8};

  for (int kt = kt_lo; kt <= kt_hi; ++kt) {
    const int buf = (kt - kt_lo) & 1;
    if (kt < kt_hi) {
      load_tile_async<HD>(sK[buf ^ 1], gk, ld, (kt + 1) * kBKV, T, tid);
      load_tile_async<HD>(sV[buf ^ 1], gv, ld, (kt + 1) * kBKV, T, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    if (kt == kt_lo) {
#pragma unroll
      for (int ks = 0; ks < G::kKSteps; ++ks) {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int ch = ks * 2 + (lane >> 4);
        ldmatrix_x4(qf[ks], smem_u32(sQ) + G::off(r, ch));
      }
    }

    // ---- S = Q K^T (16 x 64 per warp)
    float s[kBKV / 8][4];
#pragma unroll
    for (int nb = 0; nb < kBKV / 8; ++nb) s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
    const uint32_t sk = smem_u32(sK[buf]);
#pragma unroll
    for (int ks = 0; ks < G::kKSteps; ++ks) {
#pragma unroll
      for (int nb = 0; nb < kBKV / 8; nb += 2) {
        uint32_t kf[4];
        const int r = nb * 8 + (lane & 7) + ((lane >> 4) & 1) * 8;
        const int ch = ks * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(kf, sk + G::off(r, ch));
        mma_bf16_16816(s[nb], qf[ks], kf[0], kf[1]);
        mma_bf16_16816(s[nb + 1], qf[ks], kf[2], kf[3]);
      }
    }

    // ---- band / tail / generic mask (skipped for tiles entirely inside the band)
    const int k0 = kt * kBKV;
    const bool interior = (k0 - (q0 + kBQ - 1) >= -wl) && (k0 + kBKV - 1 - q0 <= wr) &&
                          (k0 + kBKV - 1 < T) && (mask == nullptr);
    if (!interior) {
#pragma unroll
      for (int nb = 0; nb < kBKV / 8; ++nb) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int q = qrow[e >> 1];
          const int k = k0 + nb * 8 + (lane & 3) * 2 + (e & 1);
          bool ok = (k < T) && (k - q >= -wl) && (k - q <= wr);
          if (ok && mask != nullptr && q < T) ok = mask[static_cast<int64_t>(q) * T + k] == 0;
          if (!ok) s[nb][e] = -INFINITY;
        }
      }
    }

    // ---- online softmax (rows lane/4 and lane/4+8)
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      float mx = -INFINITY;
#pragma unroll
      for (int nb = 0; nb < kBKV / 8; ++nb) mx = fmaxf(mx, fmaxf(s[nb][2 * rr], s[nb][2 * rr + 1]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run[rr], mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float corr = exp2f((m_run[rr] - m_use) * scale_log2);
      const float moff = m_use * scale_log2;
      float sum = 0.f;
#pragma unroll
      for (int nb = 0; nb < kBKV / 8; ++nb) {
        const float p0 = exp2f(s[nb][2 * rr] * scale_log2 - moff);
        const float p1 = exp2f(s[nb][2 * rr + 1] * scale_log2 - moff);
        s[nb][2 * rr] = p0;
        s[nb][2 * rr + 1] = p1;
        sum += p0 + p1;
      }
      l_run[rr] = l_run[rr] * corr + sum;
      m_run[rr] = m_new;
#pragma unroll
      for (int db = 0; db < G::kDBlocks; ++db) {
        o_acc[db][2 * rr] *= corr;
        o_acc[db][2 * rr + 1] *= corr;
      }
    }

    // ---- O += P V
    const uint32_t sv = smem_u32(sV[buf]);
#pragma unroll
    for (int kk = 0; kk < kBKV / 16; ++kk) {
      uint32_t pf[4];
      pf[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      pf[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      pf[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pf[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int db = 0; db < G::kDBlocks; db += 2) {
        uint32_t vf[4];
        const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int ch = db + (lane >> 4);
        ldmatrix_x4_trans(vf, sv + G::off(r, ch));
        mma_bf16_16816(o_acc[db], pf, vf[0], vf[1]);
        mma_bf16_16816(o_acc[db + 1], pf, vf[2], vf[3]);
      }
    }
    __syncthreads();  // everyone done with sK/sV[buf] before it is refilled
  }

  // ---- normalise, stage through sQ (all warps are past their Q-fragment loads), coalesced store
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    float l = l_run[rr];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const float inv = 1.0f / l;  // 0/0 -> NaN for a fully masked row, as the reference's softmax
    const int r = warp * 16 + (lane >> 2) + rr * 8;
    if (lse != nullptr && (lane & 3) == 0 && q0 + r < T)  // log2-domain log-sum-exp, for the backward
      lse[(static_cast<int64_t>(b) * H + h) * T + q0 + r] = m_run[rr] * scale_log2 + log2f(l);
#pragma unroll
    for (int db = 0; db < G::kChunks; ++db) {
      const uint32_t v = pack_bf16(o_acc[db][2 * rr] * inv, o_acc[db][2 * rr + 1] * inv);
      *reinterpret_cast<uint32_t*>(sQ + G::off(r, db) + (lane & 3) * 4) = v;
    }
  }
  __syncthreads();
  __nv_bfloat16* go = out + static_cast<int64_t>(b) * T * D + h * HD;
  for (int idx = tid; idx < kBQ * G::kChunks; idx += 128) {
    const int r = idx / G::kChunks;
    const int ch = idx - r * G::kChunks;
    if (q0 + r < T) {
      const uint4 v = *reinterpret_cast<const uint4*>(sQ + G::off(r, ch));
      *reinterpret_cast<uint4*>(go + static_cast<int64_t>(q0 + r) * D + ch * 8) = v;
    }
  }
}

template <int HD>
static int launch_band(const void* qkv, void* out, int B, int T, int H, int wl, int wr,
                       const uint8_t* mask, float* lse, cudaStream_t stream) {
  constexpr int smem = 5 * Geo<HD>::kTileBytes;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_band_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(-5, cudaGetErrorString(e));
    configured = true;
  }
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  dim3 grid((T + kBQ - 1) / kBQ, H, B);
  attn_band_kernel<HD><<<grid, 128, smem, stream>>>(static_cast<const __nv_bfloat16*>(qkv),
                                                   static_cast<__nv_bfloat16*>(out), T, H, wl, wr, mask,
                                                   scale_log2, lse);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}


// =================================================================================== backward
// Flash-style backward of the same attention (autograd of nn.MultiheadAttention's core in the
// reference, train.py:257).  P is recomputed from Q, K and the forward's log2-domain LSE:
//   P = exp2(S * scale_log2 - lse),  delta = rowsum(dO * O) (computed by the dQ kernel),  dS = P * (dP - delta),
//   dQ = scale * dS K,  dK = scale * dS^T Q,  dV = P^T dO,  with dP = dO V^T.
// Two kernels so that no atomics are needed: one owns 64 queries (dQ), one owns 64 keys (dK, dV).
// head_dim 64 and 72 (DiT-XL) share the code through Geo<HD>: for 72 the contraction over the head
// dimension runs 5 k-steps with a zero padding chunk and the 10th output block is dropped.

// Column sums over the CTA's 64 rows of a [64 x HD] accumulator tile held as mma fragments (rows lane/4 and
// lane/4 + 8 of each warp's 16, columns db*8 + (lane&3)*2 + {0,1}), added to dst[0..HD).  Must be called
// by all 128 threads after the main loop's final __syncthreads (s_col aliases the dead tiles).
template <int HD>
__device__ __forceinline__ void tile_colsum(const float (&acc)[Geo<HD>::kDBlocks][4], float mul, float* s_col,
                                            float* __restrict__ dst, int warp, int lane, int tid) {
  using G = Geo<HD>;
  if (tid < HD) s_col[tid] = 0.f;
  __syncthreads();
#pragma unroll
  for (int db = 0; db < G::kChunks; ++db) {
    float c0 = acc[db][0] + acc[db][2], c1 = acc[db][1] + acc[db][3];
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      c0 += __shfl_xor_sync(0xffffffffu, c0, o);
      c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    }
    if (lane < 4) {
      atomicAdd(&s_col[db * 8 + lane * 2], c0);
      atomicAdd(&s_col[db * 8 + lane * 2 + 1], c1);
    }
  }
  __syncthreads();
  if (tid < HD) atomicAdd(dst + tid, s_col[tid] * mul);
}

template <int HD>
__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out_fwd,
                   const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse, float* __restrict__ delta,
                   __nv_bfloat16* __restrict__ dqkv, int T, int H, int wl, int wr, float scale_log2,
                   float scale, float* __restrict__ dbias) {
  using G = Geo<HD>;
  extern __shared__ __align__(128) uint8_t smem_attn[];
  uint8_t* sQ = smem_attn;
  uint8_t* sdO = sQ + G::kTileBytes;
  uint8_t* sK[2] = {sQ + 2 * G::kTileBytes, sQ + 3 * G::kTileBytes};
  uint8_t* sV[2] = {sQ + 4 * G::kTileBytes, sQ + 5 * G::kTileBytes};
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kBQ, h = blockIdx.y, b = blockIdx.z;
  const int D = H * HD;
  const int64_t ld = 3 * static_cast<int64_t>(D);
  const __nv_bfloat16* gq = qkv + static_cast<int64_t>(b) * T * ld + h * HD;
  const __nv_bfloat16* gk = gq + D;
  const __nv_bfloat16* gv = gq + 2 * D;
  const __nv_bfloat16* gdo = dout + static_cast<int64_t>(b) * T * D + h * HD;

  if (G::kPadded) {  // zero the padding chunk of every row once; cp.async never writes it
    for (int idx = tid; idx < 6 * kBKV; idx += 128)
      *reinterpret_cast<uint4*>(smem_attn + (idx / kBKV) * G::kTileBytes + G::off(idx % kBKV, G::kChunks)) =
          make_uint4(0, 0, 0, 0);
  }
  const int kt_lo = max(0, q0 - wl) / kBKV;
  const int kt_hi = min(T - 1, min(q0 + kBQ - 1, T - 1) + wr) / kBKV;
  load_tile_async<HD>(sQ, gq, ld, q0, T, tid);
  load_tile_async<HD>(sdO, gdo, D, q0, T, tid);
  load_tile_async<HD>(sK[0], gk, ld, kt_lo * kBKV, T, tid);
  load_tile_async<HD>(sV[0], gv, ld, kt_lo * kBKV, T, tid);
  cp_async_commit();

  float dq[G::kDBlocks][4];
#pragma unroll
  for (int i = 0; i < G::kDBlocks; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
  uint32_t qf[G::kKSteps][4], dof[G::kKSteps][4];
  const int qrow[2] = {q0 + warp * 16 + (lane >> 2), q0 + warp * 16 + (lane >> 2) + 8};
  float row_lse[2], row_delta[2];
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const bool ok = qrow[rr] < T;
    const int64_t idx = (static_cast<int64_t>(b) * H + h) * T + (ok ? qrow[rr] : 0);
    row_lse[rr] = ok ? lse[idx] : 0.f;
  }
  {  // delta = rowsum(dO * O) of this warp's 16 query rows (lane -> row lane/2, half lane&1), published for the
     // dK/dV kernel that runs next and handed to the two fragment rows this lane owns
    const int r = q0 + warp * 16 + (lane >> 1);
    float acc = 0.f;
    if (r < T) {
      const __nv_bfloat162* po = reinterpret_cast<const __nv_bfloat162*>(
          out_fwd + (static_cast<int64_t>(b) * T + r) * D + h * HD) + (lane & 1) * (HD / 4);
      const __nv_bfloat162* pd = reinterpret_cast<const __nv_bfloat162*>(gdo + static_cast<int64_t>(r) * D) +
                                 (lane & 1) * (HD / 4);
#pragma unroll
      for (int i = 0; i < HD / 4; ++i) {
        const float2 a = __bfloat1622float2(po[i]);
        const float2 d = __bfloat1622float2(pd[i]);
        acc = fmaf(a.x, d.x, fmaf(a.y, d.y, acc));
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if ((lane & 1) == 0 && r < T) delta[(static_cast<int64_t>(b) * H + h) * T + r] = acc;
    row_delta[0] = __shfl_sync(0xffffffffu, acc, 2 * (lane >> 2));
    row_delta[1] = __shfl_sync(0xffffffffu, acc, 2 * (lane >> 2) + 16);
  }

  for (int kt = kt_lo; kt <= kt_hi; ++kt) {
    const int buf = (kt - kt_lo) & 1;
    if (kt < kt_hi) {
      load_tile_async<HD>(sK[buf ^ 1], gk, ld, (kt + 1) * kBKV, T, tid);
      load_tile_async<HD>(sV[buf ^ 1], gv, ld, (kt + 1) * kBKV, T, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kt == kt_lo) {
#pragma unroll
      for (int ks = 0; ks < G::kKSteps; ++ks) {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int ch = ks * 2 + (lane >> 4);
        ldmatrix_x4(qf[ks], smem_u32(sQ) + G::off(r, ch));
        ldmatrix_x4(dof[ks], smem_u32(sdO) + G::off(r, ch));
      }
    }
    float s[kBKV / 8][4], dp[kBKV / 8][4];
#pragma unroll
    for (int nb = 0; nb < kBKV / 8; ++nb) {
      s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
      dp[nb][0] = dp[nb][1] = dp[nb][2] = dp[nb][3] = 0.f;
    }
    const uint32_t sk = smem_u32(sK[buf]), sv = smem_u32(sV[buf]);
#pragma unroll
    for (int ks = 0; ks < G::kKSteps; ++ks) {
#pragma unroll
      for (int nb = 0; nb < kBKV / 8; nb += 2) {
        uint32_t kf[4], vf[4];
        const int r = nb * 8 + (lane & 7) + ((lane >> 4) & 1) * 8;
        const int ch = ks * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(kf, sk + G::off(r, ch));
        ldmatrix_x4(vf, sv + G::off(r, ch));
        mma_bf16_16816(s[nb], qf[ks], kf[0], kf[1]);
        mma_bf16_16816(s[nb + 1], qf[ks], kf[2], kf[3]);
        mma_bf16_16816(dp[nb], dof[ks], vf[0], vf[1]);
        mma_bf16_16816(dp[nb + 1], dof[ks], vf[2], vf[3]);
      }
    }
    const int k0 = kt * kBKV;
#pragma unroll
    for (int nb = 0; nb < kBKV / 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int q = qrow[e >> 1];
        const int k = k0 + nb * 8 + (lane & 3) * 2 + (e & 1);
        const bool ok = (k < T) && (q < T) && (k - q >= -wl) && (k - q <= wr);
        const float p = ok ? exp2f(s[nb][e] * scale_log2 - row_lse[e >> 1]) : 0.f;
        s[nb][e] = p * (dp[nb][e] - row_delta[e >> 1]);  // dS
      }
    }
#pragma unroll
    for (int kk = 0; kk < kBKV / 16; ++kk) {
      uint32_t pf[4];
      pf[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      pf[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      pf[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pf[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int db = 0; db < G::kDBlocks; db += 2) {
        uint32_t kf[4];
        const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int ch = db + (lane >> 4);
        ldmatrix_x4_trans(kf, sk + G::off(r, ch));
        mma_bf16_16816(dq[db], pf, kf[0], kf[1]);
        mma_bf16_16816(dq[db + 1], pf, kf[2], kf[3]);
      }
    }
    __syncthreads();
  }
  __nv_bfloat16* gdq = dqkv + static_cast<int64_t>(b) * T * ld + h * HD;
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    if (qrow[rr] >= T) continue;
#pragma unroll
    for (int db = 0; db < G::kChunks; ++db)
      *reinterpret_cast<uint32_t*>(gdq + static_cast<int64_t>(qrow[rr]) * ld + db * 8 + (lane & 3) * 2) =
          pack_bf16(dq[db][2 * rr] * scale, dq[db][2 * rr + 1] * scale);
  }
  if (dbias != nullptr)  // in_proj_bias gradient, q part (rows >= T contribute exact zeros)
    tile_colsum<HD>(dq, scale, reinterpret_cast<float*>(smem_attn), dbias + h * HD, warp, lane, tid);
}

template <int HD>
__global__ void __launch_bounds__(128)
attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                    const float* __restrict__ lse, const float* __restrict__ delta,
                    __nv_bfloat16* __restrict__ dqkv, int T, int H, int wl, int wr, float scale_log2,
                    float scale, float* __restrict__ dbias) {
  using G = Geo<HD>;
  extern __shared__ __align__(128) uint8_t smem_attn[];
  uint8_t* sK = smem_attn;
  uint8_t* sV = sK + G::kTileBytes;
  uint8_t* sQ[2] = {sK + 2 * G::kTileBytes, sK + 3 * G::kTileBytes};
  uint8_t* sdO[2] = {sK + 4 * G::kTileBytes, sK + 5 * G::kTileBytes};
  float* s_lse = reinterpret_cast<float*>(sK + 6 * G::kTileBytes);  // [2][64]
  float* s_delta = s_lse + 2 * kBQ;                                 // [2][64]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k0 = blockIdx.x * kBKV, h = blockIdx.y, b = blockIdx.z;
  const int D = H * HD;
  const int64_t ld = 3 * static_cast<int64_t>(D);
  const __nv_bfloat16* gq = qkv + static_cast<int64_t>(b) * T * ld + h * HD;
  const __nv_bfloat16* gk = gq + D;
  const __nv_bfloat16* gv = gq + 2 * D;
  const __nv_bfloat16* gdo = dout + static_cast<int64_t>(b) * T * D + h * HD;
  const float* g_lse = lse + (static_cast<int64_t>(b) * H + h) * T;
  const float* g_delta = delta + (static_cast<int64_t>(b) * H + h) * T;

  // queries that may see any of these keys: key - query in [-wl, wr]
  const int qt_lo = max(0, k0 - wr) / kBQ;
  const int qt_hi = min(T - 1, min(k0 + kBKV - 1, T - 1) + wl) / kBQ;
  auto load_stats = [&](int buf, int qbase) {
    if (tid < kBQ) {
      const int q = qbase + tid;
      s_lse[buf * kBQ + tid] = q < T ? g_lse[q] : 0.f;
      s_delta[buf * kBQ + tid] = q < T ? g_delta[q] : 0.f;
    }
  };
  if (G::kPadded) {  // zero the padding chunk of every row once; cp.async never writes it
    for (int idx = tid; idx < 6 * kBKV; idx += 128)
      *reinterpret_cast<uint4*>(smem_attn + (idx / kBKV) * G::kTileBytes + G::off(idx % kBKV, G::kChunks)) =
          make_uint4(0, 0, 0, 0);
  }
  load_tile_async<HD>(sK, gk, ld, k0, T, tid);
  load_tile_async<HD>(sV, gv, ld, k0, T, tid);
  load_tile_async<HD>(sQ[0], gq, ld, qt_lo * kBQ, T, tid);
  load_tile_async<HD>(sdO[0], gdo, D, qt_lo * kBQ, T, tid);
  cp_async_commit();
  load_stats(0, qt_lo * kBQ);

  float dk[G::kDBlocks][4], dv[G::kDBlocks][4];
#pragma unroll
  for (int i = 0; i < G::kDBlocks; ++i) {
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
  }
  uint32_t kfr[G::kKSteps][4], vfr[G::kKSteps][4];
  const int krow[2] = {k0 + warp * 16 + (lane >> 2), k0 + warp * 16 + (lane >> 2) + 8};

  for (int qt = qt_lo; qt <= qt_hi; ++qt) {
    const int buf = (qt - qt_lo) & 1;
    if (qt < qt_hi) {
      load_tile_async<HD>(sQ[buf ^ 1], gq, ld, (qt + 1) * kBQ, T, tid);
      load_tile_async<HD>(sdO[buf ^ 1], gdo, D, (qt + 1) * kBQ, T, tid);
      cp_async_commit();
      load_stats(buf ^ 1, (qt + 1) * kBQ);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (qt == qt_lo) {
#pragma unroll
      for (int ks = 0; ks < G::kKSteps; ++ks) {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int ch = ks * 2 + (lane >> 4);
        ldmatrix_x4(kfr[ks], smem_u32(sK) + G::off(r, ch));
        ldmatrix_x4(vfr[ks], smem_u32(sV) + G::off(r, ch));
      }
    }
    // S^T[key, query] and dP^T[key, query]
    float st[kBQ / 8][4], dpt[kBQ / 8][4];
#pragma unroll
    for (int nb = 0; nb < kBQ / 8; ++nb) {
      st[nb][0] = st[nb][1] = st[nb][2] = st[nb][3] = 0.f;
      dpt[nb][0] = dpt[nb][1] = dpt[nb][2] = dpt[nb][3] = 0.f;
    }
    const uint32_t sq = smem_u32(sQ[buf]), sdo = smem_u32(sdO[buf]);
#pragma unroll
    for (int ks = 0; ks < G::kKSteps; ++ks) {
#pragma unroll
      for (int nb = 0; nb < kBQ / 8; nb += 2) {
        uint32_t qfr[4], dofr[4];
        const int r = nb * 8 + (lane & 7) + ((lane >> 4) & 1) * 8;
        const int ch = ks * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(qfr, sq + G::off(r, ch));
        ldmatrix_x4(dofr, sdo + G::off(r, ch));
        mma_bf16_16816(st[nb], kfr[ks], qfr[0], qfr[1]);
        mma_bf16_16816(st[nb + 1], kfr[ks], qfr[2], qfr[3]);
        mma_bf16_16816(dpt[nb], vfr[ks], dofr[0], dofr[1]);
        mma_bf16_16816(dpt[nb + 1], vfr[ks], dofr[2], dofr[3]);
      }
    }
    const int qb = qt * kBQ;
    uint32_t pf[kBQ / 16][4], dsf[kBQ / 16][4];
#pragma unroll
    for (int nb = 0; nb < kBQ / 8; ++nb) {
      float pv[4], dsv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = krow[e >> 1];
        const int qc = nb * 8 + (lane & 3) * 2 + (e & 1);
        const int q = qb + qc;
        const bool ok = (k < T) && (q < T) && (k - q >= -wl) && (k - q <= wr);
        const float p = ok ? exp2f(st[nb][e] * scale_log2 - s_lse[buf * kBQ + qc]) : 0.f;
        pv[e] = p;
        dsv[e] = p * (dpt[nb][e] - s_delta[buf * kBQ + qc]);
      }
      // accumulator blocks (2kk, 2kk+1) form the A fragment of k-step kk
      pf[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16(pv[0], pv[1]);
      pf[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16(pv[2], pv[3]);
      dsf[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16(dsv[0], dsv[1]);
      dsf[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16(dsv[2], dsv[3]);
    }
#pragma unroll
    for (int kk = 0; kk < kBQ / 16; ++kk) {
#pragma unroll
      for (int db = 0; db < G::kDBlocks; db += 2) {
        uint32_t dofr[4], qfr[4];
        const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int ch = db + (lane >> 4);
        ldmatrix_x4_trans(dofr, sdo + G::off(r, ch));
        ldmatrix_x4_trans(qfr, sq + G::off(r, ch));
        mma_bf16_16816(dv[db], pf[kk], dofr[0], dofr[1]);
        mma_bf16_16816(dv[db + 1], pf[kk], dofr[2], dofr[3]);
        mma_bf16_16816(dk[db], dsf[kk], qfr[0], qfr[1]);
        mma_bf16_16816(dk[db + 1], dsf[kk], qfr[2], qfr[3]);
      }
    }
    __syncthreads();
  }
  __nv_bfloat16* gdk = dqkv + static_cast<int64_t>(b) * T * ld + D + h * HD;
  __nv_bfloat16* gdv = gdk + D;
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    if (krow[rr] >= T) continue;
#pragma unroll
    for (int db = 0; db < G::kChunks; ++db) {
      const int64_t off = static_cast<int64_t>(krow[rr]) * ld + db * 8 + (lane & 3) * 2;
      *reinterpret_cast<uint32_t*>(gdk + off) = pack_bf16(dk[db][2 * rr] * scale, dk[db][2 * rr + 1] * scale);
      *reinterpret_cast<uint32_t*>(gdv + off) = pack_bf16(dv[db][2 * rr], dv[db][2 * rr + 1]);
    }
  }
  if (dbias != nullptr) {  // in_proj_bias gradient, k and v parts
    float* s_col = reinterpret_cast<float*>(smem_attn);
    tile_colsum<HD>(dk, scale, s_col, dbias + D + h * HD, warp, lane, tid);
    tile_colsum<HD>(dv, 1.0f, s_col + 128, dbias + 2 * D + h * HD, warp, lane, tid);
  }
}

bool attn_window_applicable(int T, int head_dim, int w_left, int w_right, const uint8_t* mask);
int attn_window_launch(const void* qkv, void* out, int B, int T, int H, int w_left, int w_right,
                       cudaStream_t stream);
bool attn_bwd_tc_applicable(int T, int head_dim);
int attn_bwd_tc_launch(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B, int T,
                       int H, int w_left, int w_right, float* dbias, cudaStream_t stream);
bool attn_fa_applicable(int head_dim, const uint8_t* mask);
int attn_fa_launch(const void* qkv, void* out, int B, int T, int H, int w_left, int w_right, float* lse,
                   cudaStream_t stream);
bool attn_stream_applicable(int head_dim, const uint8_t* mask);
int attn_stream_launch(const void* qkv, void* out, int B, int T, int H, int w_left, int w_right, float* lse,
                       cudaStream_t stream);

}  // namespace osudit

using namespace osudit;

extern "C" int osudit_colsum(const void* in, int in_is_f32, int64_t rows, int N, float* out, void* stream);

extern "C" int osudit_attn_band(const void* qkv, void* out, int B, int T, int H, int head_dim,
                                int w_left, int w_right, const uint8_t* mask, int algo, float* lse,
                                void* stream) {
  if (head_dim != 64 && head_dim != 72)
    return set_error(-1, "attn_band: head_dim must be 64 (DiT-S/B/L) or 72 (DiT-XL)");
  if (B <= 0 || T <= 0 || H <= 0) return set_error(-1, "attn_band: bad shape");
  if (B > 65535 || H > 65535) return set_error(-1, "attn_band: batch/heads exceed the launch grid");
  if (w_left < 0 || w_left > T) w_left = T;
  if (w_right < 0 || w_right > T) w_right = T;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool window_ok = lse == nullptr && attn_window_applicable(T, head_dim, w_left, w_right, mask);
  if (algo == OSUDIT_ATTN_TCGEN05 && !window_ok)
    return set_error(-1, "attn_band: tcgen05 window kernel needs head_dim 64, no generic mask, and "
                         "a band within +-128 or T <= 256");
  const bool fa_ok = attn_fa_applicable(head_dim, mask);
  if (algo == OSUDIT_ATTN_FA && !fa_ok)
    return set_error(-1, "attn_band: the streaming tcgen05 kernel needs head_dim 64 and no generic mask");
  // AUTO, head_dim 64 without a generic mask: the double-buffered streaming kernel (attn_stream.cu) for every shape —
  // 0.60 ms per DiT-B layer on the sampling band (window kernel 0.73, two-slot kernel 0.73), 0.047 ms on 256 training
  // windows of 128 (two-slot 0.056), 0.307 ms at 128 x 512 x 16 heads (0.310).  The older tcgen05 kernels stay
  // selectable: OSUDIT_ATTN_STREAM=0 restores the round-2a choice (two-slot kernel, window kernel for the long band),
  // additionally OSUDIT_ATTN_FA=0 the round-1 choice (window kernel / mma.sync).
  static const bool prefer_stream = [] {
    const char* e = getenv("OSUDIT_ATTN_STREAM");
    return !(e && e[0] == '0');
  }();
  static const bool prefer_fa = [] {
    const char* e = getenv("OSUDIT_ATTN_FA");
    return !(e && e[0] == '0');
  }();
  static const bool band_window = [] {  // OSUDIT_ATTN_BAND_WINDOW=0: the two-slot kernel for the sampling band too
    const char* e = getenv("OSUDIT_ATTN_BAND_WINDOW");
    return !(e && e[0] == '0');
  }();
  if (algo == OSUDIT_ATTN_STREAM && !fa_ok)
    return set_error(-1, "attn_band: the streaming tcgen05 kernel needs head_dim 64 and no generic mask");
  if (algo == OSUDIT_ATTN_STREAM || (algo == OSUDIT_ATTN_AUTO && fa_ok && prefer_stream))
    return attn_stream_launch(qkv, out, B, T, H, w_left, w_right, lse, st);
  const bool long_band = band_window && window_ok && T > 256;
  if (algo == OSUDIT_ATTN_FA || (algo == OSUDIT_ATTN_AUTO && fa_ok && prefer_fa && !long_band))
    return attn_fa_launch(qkv, out, B, T, H, w_left, w_right, lse, st);
  if (algo != OSUDIT_ATTN_MMA_SYNC && window_ok)
    return attn_window_launch(qkv, out, B, T, H, w_left, w_right, st);
  if (head_dim == 64) return launch_band<64>(qkv, out, B, T, H, w_left, w_right, mask, lse, st);
  return launch_band<72>(qkv, out, B, T, H, w_left, w_right, mask, lse, st);
}

// ----------------------------------------------------------------- backward, whole sequence in one CTA
// Training sequences are short (seq-len 128 in BASELINE configs 3): per (sample, head) the whole problem —
// Q, K, V, dO (4 x 16 KB) and the recomputed P and dS (2 x 32 KB, bf16) — fits in shared memory, so one CTA of
// 8 warps produces dQ, dK and dV together: S and dP are computed once (the two-kernel path computes them twice),
// every operand is loaded once, and nothing but the results touches HBM.
//   phase 1 (warp = 16 query rows): S = Q K^T, P = exp2(S*scale_log2 - lse), dP = dO V^T, dS = P*(dP - delta)
//            -> P, dS to shared memory;  dQ = scale * dS K from the register fragments.
//   phase 2 (warp = 16 key rows):  dV = P^T dO,  dK = scale * dS^T Q  (A operands = P / dS read through
//            ldmatrix.trans, i.e. transposed on the fly).
// head_dim 64, T <= 128; band semantics as everywhere else.
// Measured (B200, 256 x 12 heads x 128): 334 us per layer vs 460 us for the two-kernel path.  ncu: HMMA pipe 18 %,
// 8 warps per SM; a 16-warp variant (rows x key halves in phase 1, keys x head-dim halves in phase 2, <= 128
// registers) was correct but slower (376 us): the legacy mma.sync path is the limit here, not latency hiding.
constexpr int kSmallT = 128;

__device__ __forceinline__ uint32_t off_p(int r, int ch) {  // [128][128] bf16, 256-byte rows, XOR-swizzled chunks
  return r * 256 + ((ch ^ (r & 7)) << 4);
}

__global__ void __launch_bounds__(256, 1)
attn_bwd_small_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out_fwd,
                      const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                      __nv_bfloat16* __restrict__ dqkv, int T, int H, int wl, int wr, float scale_log2, float scale,
                      float* __restrict__ dbias) {
  constexpr int HD = 64;
  using G = Geo<HD>;
  extern __shared__ __align__(128) uint8_t smem_attn[];
  uint8_t* sQ = smem_attn;
  uint8_t* sK = sQ + 2 * G::kTileBytes;
  uint8_t* sV = sK + 2 * G::kTileBytes;
  uint8_t* sdO = sV + 2 * G::kTileBytes;
  uint8_t* sP = sdO + 2 * G::kTileBytes;
  uint8_t* sdS = sP + kSmallT * 256;
  float* s_col = reinterpret_cast<float*>(sdS + kSmallT * 256);  // [3][64] bias-gradient partial sums
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int D = H * HD;
  const int64_t ld = 3 * static_cast<int64_t>(D);
  const __nv_bfloat16* gq = qkv + static_cast<int64_t>(b) * T * ld + h * HD;
  const __nv_bfloat16* gdo = dout + static_cast<int64_t>(b) * T * D + h * HD;
  const __nv_bfloat16* go = out_fwd + static_cast<int64_t>(b) * T * D + h * HD;

  for (int idx = tid; idx < kSmallT * G::kChunks; idx += 256) {
    const int r = idx >> 3, ch = idx & 7;
    const bool valid = r < T;
    const int64_t rr = valid ? r : 0;
    cp_async_16(sQ + G::off(r, ch), gq + rr * ld + ch * 8, valid);
    cp_async_16(sK + G::off(r, ch), gq + rr * ld + D + ch * 8, valid);
    cp_async_16(sV + G::off(r, ch), gq + rr * ld + 2 * D + ch * 8, valid);
    cp_async_16(sdO + G::off(r, ch), gdo + rr * D + ch * 8, valid);
  }
  cp_async_commit();
  if (tid < 192) s_col[tid] = 0.f;

  // ---- per-row statistics of this warp's 16 queries (delta as in attn_bwd_dq_kernel)
  const int q_base = warp * 16;
  const int qrow[2] = {q_base + (lane >> 2), q_base + (lane >> 2) + 8};
  float row_lse[2], row_delta[2];
#pragma unroll
  for (int rr = 0; rr < 2; ++rr)
    row_lse[rr] = qrow[rr] < T ? lse[(static_cast<int64_t>(b) * H + h) * T + qrow[rr]] : 0.f;
  {
    const int r = q_base + (lane >> 1);
    float acc = 0.f;
    if (r < T) {
      const __nv_bfloat162* po = reinterpret_cast<const __nv_bfloat162*>(go + static_cast<int64_t>(r) * D) + (lane & 1) * 16;
      const __nv_bfloat162* pd = reinterpret_cast<const __nv_bfloat162*>(gdo + static_cast<int64_t>(r) * D) + (lane & 1) * 16;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 a = __bfloat1622float2(po[i]);
        const float2 d = __bfloat1622float2(pd[i]);
        acc = fmaf(a.x, d.x, fmaf(a.y, d.y, acc));
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    row_delta[0] = __shfl_sync(0xffffffffu, acc, 2 * (lane >> 2));
    row_delta[1] = __shfl_sync(0xffffffffu, acc, 2 * (lane >> 2) + 16);
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---------------------------------------------------------------- phase 1
  {
    uint32_t qf[4][4], dof[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int r = q_base + (lane & 7) + ((lane >> 3) & 1) * 8;
      const int ch = ks * 2 + (lane >> 4);
      ldmatrix_x4(qf[ks], smem_u32(sQ) + G::off(r, ch));
      ldmatrix_x4(dof[ks], smem_u32(sdO) + G::off(r, ch));
    }
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    const uint32_t sk = smem_u32(sK), sv = smem_u32(sV);
#pragma unroll 1
    for (int kh = 0; kh < 2; ++kh) {  // two halves of 64 keys keep the accumulators at 64 registers
      const int k0 = kh * 64;
      if (k0 >= T) break;
      float s[8][4], dp[8][4];
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
        dp[nb][0] = dp[nb][1] = dp[nb][2] = dp[nb][3] = 0.f;
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int nb = 0; nb < 8; nb += 2) {
          uint32_t kf[4], vf[4];
          const int r = k0 + nb * 8 + (lane & 7) + ((lane >> 4) & 1) * 8;
          const int ch = ks * 2 + ((lane >> 3) & 1);
          ldmatrix_x4(kf, sk + G::off(r, ch));
          ldmatrix_x4(vf, sv + G::off(r, ch));
          mma_bf16_16816(s[nb], qf[ks], kf[0], kf[1]);
          mma_bf16_16816(s[nb + 1], qf[ks], kf[2], kf[3]);
          mma_bf16_16816(dp[nb], dof[ks], vf[0], vf[1]);
          mma_bf16_16816(dp[nb + 1], dof[ks], vf[2], vf[3]);
        }
      }
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        float pv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int q = qrow[e >> 1];
          const int k = k0 + nb * 8 + (lane & 3) * 2 + (e & 1);
          const bool ok = (k < T) && (q < T) && (k - q >= -wl) && (k - q <= wr);
          const float p = ok ? exp2f(s[nb][e] * scale_log2 - row_lse[e >> 1]) : 0.f;
          pv[e] = p;
          s[nb][e] = p * (dp[nb][e] - row_delta[e >> 1]);  // dS
        }
        const int ch = (k0 >> 3) + nb;
        const int cb = (lane & 3) * 4;
        *reinterpret_cast<uint32_t*>(sP + off_p(qrow[0], ch) + cb) = pack_bf16(pv[0], pv[1]);
        *reinterpret_cast<uint32_t*>(sP + off_p(qrow[1], ch) + cb) = pack_bf16(pv[2], pv[3]);
        *reinterpret_cast<uint32_t*>(sdS + off_p(qrow[0], ch) + cb) = pack_bf16(s[nb][0], s[nb][1]);
        *reinterpret_cast<uint32_t*>(sdS + off_p(qrow[1], ch) + cb) = pack_bf16(s[nb][2], s[nb][3]);
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {  // dQ += dS[:, 16 keys] K[16 keys, :]
        uint32_t pf[4];
        pf[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
        pf[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
        pf[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pf[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int db = 0; db < 8; db += 2) {
          uint32_t kf[4];
          const int r = k0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          const int ch = db + (lane >> 4);
          ldmatrix_x4_trans(kf, sk + G::off(r, ch));
          mma_bf16_16816(dq[db], pf, kf[0], kf[1]);
          mma_bf16_16816(dq[db + 1], pf, kf[2], kf[3]);
        }
      }
    }
    if (T <= 64) {  // the second half of P / dS is read by phase 2 of short sequences: zero it
      for (int nb = 8; nb < 16; ++nb) {
        const int cb = (lane & 3) * 4;
        *reinterpret_cast<uint32_t*>(sP + off_p(qrow[0], nb) + cb) = 0u;
        *reinterpret_cast<uint32_t*>(sP + off_p(qrow[1], nb) + cb) = 0u;
        *reinterpret_cast<uint32_t*>(sdS + off_p(qrow[0], nb) + cb) = 0u;
        *reinterpret_cast<uint32_t*>(sdS + off_p(qrow[1], nb) + cb) = 0u;
      }
    }
    __nv_bfloat16* gdq = dqkv + static_cast<int64_t>(b) * T * ld + h * HD;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      if (qrow[rr] >= T) continue;
#pragma unroll
      for (int db = 0; db < 8; ++db)
        *reinterpret_cast<uint32_t*>(gdq + static_cast<int64_t>(qrow[rr]) * ld + db * 8 + (lane & 3) * 2) =
            pack_bf16(dq[db][2 * rr] * scale, dq[db][2 * rr + 1] * scale);
    }
    if (dbias != nullptr) {
#pragma unroll
      for (int db = 0; db < 8; ++db) {
        float c0 = dq[db][0] + dq[db][2], c1 = dq[db][1] + dq[db][3];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          c0 += __shfl_xor_sync(0xffffffffu, c0, o);
          c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        }
        if (lane < 4) {
          atomicAdd(&s_col[db * 8 + lane * 2], c0 * scale);
          atomicAdd(&s_col[db * 8 + lane * 2 + 1], c1 * scale);
        }
      }
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase 2
  {
    const int key_base = warp * 16;
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
      dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    if (key_base < T) {
      const uint32_t sq = smem_u32(sQ), sdo = smem_u32(sdO), sp = smem_u32(sP), sds = smem_u32(sdS);
      const int n_q16 = (T + 15) >> 4;
#pragma unroll 1
      for (int kk = 0; kk < n_q16; ++kk) {  // contraction over 16 queries at a time
        uint32_t pa[4], dsa[4];
        // A = P^T / dS^T [16 keys x 16 queries]: 8x8 blocks of the stored [query][key] matrices, transposed on load
        const int qr = kk * 16 + (lane & 7) + (lane >> 4) * 8;
        const int kch = (key_base >> 3) + ((lane >> 3) & 1);
        ldmatrix_x4_trans(pa, sp + off_p(qr, kch));
        ldmatrix_x4_trans(dsa, sds + off_p(qr, kch));
#pragma unroll
        for (int db = 0; db < 8; db += 2) {
          uint32_t dofr[4], qfr[4];
          const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          const int ch = db + (lane >> 4);
          ldmatrix_x4_trans(dofr, sdo + G::off(r, ch));
          ldmatrix_x4_trans(qfr, sq + G::off(r, ch));
          mma_bf16_16816(dv[db], pa, dofr[0], dofr[1]);
          mma_bf16_16816(dv[db + 1], pa, dofr[2], dofr[3]);
          mma_bf16_16816(dk[db], dsa, qfr[0], qfr[1]);
          mma_bf16_16816(dk[db + 1], dsa, qfr[2], qfr[3]);
        }
      }
    }
    const int krow[2] = {key_base + (lane >> 2), key_base + (lane >> 2) + 8};
    __nv_bfloat16* gdk = dqkv + static_cast<int64_t>(b) * T * ld + D + h * HD;
    __nv_bfloat16* gdv = gdk + D;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      if (krow[rr] >= T) continue;
#pragma unroll
      for (int db = 0; db < 8; ++db) {
        const int64_t off = static_cast<int64_t>(krow[rr]) * ld + db * 8 + (lane & 3) * 2;
        *reinterpret_cast<uint32_t*>(gdk + off) = pack_bf16(dk[db][2 * rr] * scale, dk[db][2 * rr + 1] * scale);
        *reinterpret_cast<uint32_t*>(gdv + off) = pack_bf16(dv[db][2 * rr], dv[db][2 * rr + 1]);
      }
    }
    if (dbias != nullptr) {
#pragma unroll
      for (int db = 0; db < 8; ++db) {
        float c0 = dk[db][0] + dk[db][2], c1 = dk[db][1] + dk[db][3];
        float e0 = dv[db][0] + dv[db][2], e1 = dv[db][1] + dv[db][3];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          c0 += __shfl_xor_sync(0xffffffffu, c0, o);
          c1 += __shfl_xor_sync(0xffffffffu, c1, o);
          e0 += __shfl_xor_sync(0xffffffffu, e0, o);
          e1 += __shfl_xor_sync(0xffffffffu, e1, o);
        }
        if (lane < 4) {
          atomicAdd(&s_col[64 + db * 8 + lane * 2], c0 * scale);
          atomicAdd(&s_col[64 + db * 8 + lane * 2 + 1], c1 * scale);
          atomicAdd(&s_col[128 + db * 8 + lane * 2], e0);
          atomicAdd(&s_col[128 + db * 8 + lane * 2 + 1], e1);
        }
      }
      __syncthreads();
      if (tid < 192) atomicAdd(dbias + (tid >> 6) * D + h * HD + (tid & 63), s_col[tid]);
    }
  }
}

static int launch_bwd_small(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B,
                            int T, int H, int w_left, int w_right, float* dbias, cudaStream_t st) {
  constexpr int smem = 8 * Geo<64>::kTileBytes + 2 * kSmallT * 256 + 192 * static_cast<int>(sizeof(float));
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(-5, cudaGetErrorString(e));
    configured = true;
  }
  const float scale = 0.125f;
  const float scale_log2 = 1.4426950408889634f * scale;
  attn_bwd_small_kernel<<<dim3(H, B), 256, smem, st>>>(
      static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(out),
      static_cast<const __nv_bfloat16*>(dout), lse, static_cast<__nv_bfloat16*>(dqkv), T, H, w_left, w_right,
      scale_log2, scale, dbias);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

template <int HD>
static int launch_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv, int B,
                      int T, int H, int w_left, int w_right, float* dbias, cudaStream_t st) {
  constexpr int smem_dq = 6 * Geo<HD>::kTileBytes;
  constexpr int smem_dkv = 6 * Geo<HD>::kTileBytes + 4 * kBQ * sizeof(float);
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(attn_bwd_dq_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dq);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_dkv_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dkv);
    if (e != cudaSuccess) return set_error(-5, cudaGetErrorString(e));
    configured = true;
  }
  const float scale = 1.0f / sqrtf(static_cast<float>(HD));
  const float scale_log2 = 1.4426950408889634f * scale;
  dim3 grid((T + kBQ - 1) / kBQ, H, B);
  attn_bwd_dq_kernel<HD><<<grid, 128, smem_dq, st>>>(
      static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(out),
      static_cast<const __nv_bfloat16*>(dout), lse, delta,
      static_cast<__nv_bfloat16*>(dqkv), T, H, w_left, w_right, scale_log2, scale, dbias);
  OSUDIT_CHECK_LAUNCH();
  attn_bwd_dkv_kernel<HD><<<grid, 128, smem_dkv, st>>>(
      static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(dout), lse, delta,
      static_cast<__nv_bfloat16*>(dqkv), T, H, w_left, w_right, scale_log2, scale, dbias);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

extern "C" int osudit_attn_band_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                                    float* delta, void* dqkv, int B, int T, int H, int head_dim,
                                    int w_left, int w_right, float* dbias_qkv, void* stream) {
  if (head_dim != 64 && head_dim != 72)
    return set_error(-1, "attn_band_bwd: head_dim must be 64 (DiT-S/B/L) or 72 (DiT-XL)");
  if (B <= 0 || T <= 0 || H <= 0 || B > 65535 || H > 65535) return set_error(-1, "attn_band_bwd: bad shape");
  if (w_left < 0 || w_left > T) w_left = T;
  if (w_right < 0 || w_right > T) w_right = T;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static const bool use_tc = [] {  // OSUDIT_ATTN_BWD_TC=0: the round-1 mma.sync kernels
    const char* e = getenv("OSUDIT_ATTN_BWD_TC");
    return !(e && e[0] == '0');
  }();
  if (use_tc && attn_bwd_tc_applicable(T, head_dim)) {  // tcgen05: S, dP, dV, dK, dQ in TMEM (attn_bwd_tc.cu)
    return attn_bwd_tc_launch(qkv, out, dout, lse, dqkv, B, T, H, w_left, w_right, dbias_qkv, st);
  }
  static const bool use_small = [] {
    const char* e = getenv("OSUDIT_ATTN_BWD_SMALL");
    return !(e && e[0] == '0');
  }();
  if (head_dim == 64 && T <= kSmallT && use_small)  // the whole (sample, head) problem in one CTA
    return launch_bwd_small(qkv, out, dout, lse, dqkv, B, T, H, w_left, w_right, dbias_qkv, st);
  if (head_dim == 64)
    return launch_bwd<64>(qkv, out, dout, lse, delta, dqkv, B, T, H, w_left, w_right, dbias_qkv, st);
  return launch_bwd<72>(qkv, out, dout, lse, delta, dqkv, B, T, H, w_left, w_right, dbias_qkv, st);
}
