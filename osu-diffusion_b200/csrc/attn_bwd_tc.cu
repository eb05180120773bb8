// Attention backward on tcgen05 for sequences of up to 128 datapoints (the training windows of train.py:325):
// one (sample, head) problem = five 128-row tile GEMMs with every operand in shared memory and every product in TMEM.
//
// Reference: autograd of nn.MultiheadAttention's core (models.py:130-135,164-170) under loss.backward() (train.py:257):
//   P = softmax(q k^T / sqrt(hd) + mask),  dV = P^T dO,  dP = dO V^T,  dS = P o (dP - rowsum(dO o O)),
//   dQ = dS K / sqrt(hd),  dK = dS^T Q / sqrt(hd).
// P is recomputed from the forward's log2-domain log-sum-exp (osudit_attn_band); head_dim 64.
//
// Per problem:   S = Q K^T and dP = dO V^T (tcgen05, 2 x 128 TMEM columns)
//             -> 8 softmax warps, two threads per query row: P = exp2(S c - lse), dS = P (dP - delta), both as bf16
//                into ONE 128-byte-swizzled [query][key] tile each
//             -> dV = P^T dO, dK = dS^T Q (that tile read as an MN-major A operand: no transposes), dQ = dS K
//                (the same dS tile as a K-major A operand), 3 x 64 TMEM columns
//             -> 4 epilogue warps: TMEM -> bf16 -> dqkv.
// The problems of a CTA are software-pipelined: S / dP of problem i+1 are issued in front of the three gradient
// GEMMs of problem i, the epilogue of i runs beside the softmax of i+1, Q / K / V / dO tiles are double-buffered.
//   warp 0       TMA producer: Q, K, V (3-D map over packed qkv [B, T, 3D]) and dO (3-D map over [B, T, D])
//   warp 1       tcgen05.mma issuer + TMEM allocation
//   warps 2-3    idle (keep the softmax / epilogue warps aligned to TMEM lane quadrants = warp % 4)
//   warps 4-11   softmax: quadrant = warp % 4, column half = (warp - 4) / 4
//   warps 12-15  epilogue
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "common.h"
#include "ptx.cuh"

namespace osudit {

int make_tensor_map_3d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2,
                       uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1);

namespace attn_bwd_tc {

constexpr int kT = 128;                      // rows of every tile (queries / keys)
constexpr int kHD = 64;
constexpr int kTile = kT * kHD * 2;          // 16 KB: one [128][64] bf16 box
constexpr int kStage = 5 * kTile;            // Q, K, V, dO, O (the first three double as dQ / dK / dV staging)
constexpr int kOffP = 2 * kStage;            // P [128 q][128 k] as two 64-key column blocks
constexpr int kOffDS = kOffP + 2 * kTile;
constexpr int kOffDelta = kOffDS + 2 * kTile;  // [2 halves][128 rows] partial rowsum(dO o O)
constexpr int kSmemBar = kOffDelta + 2 * kT * 4;  // 225 KB
constexpr int kSmemBytes = kSmemBar + 128 + 1024;
constexpr int kThreads = 16 * 32;
// TMEM columns
constexpr uint32_t kColS = 0, kColDP = 128, kColDV = 256, kColDK = 320, kColDQ = 384;

struct Params {
  CUtensorMap tma_qkv;   // [3D, T, B], box [64, 128, 1]
  CUtensorMap tma_dout;  // [D, T, B], box [64, 128, 1]
  CUtensorMap tma_out;   // forward output, same geometry
  CUtensorMap tma_dqkv;  // [3D, T, B], box [64, 128, 1] (stores; rows >= T are clipped)
  const float* lse;      // [B, H, T], log2 domain
  float* dbias;          // [3D] or nullptr: += column sums of dqkv (the in_proj_bias gradient)
  int B, T, H, D, problems;
  int w_left, w_right;
  float scale, scale_log2;
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar,
                                            int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// SWIZZLE_128B descriptor, 8-row groups 1024 B apart.  `lbo_bytes` (MN-major operands wider than one 64-element
// atom): distance between the atoms along M/N.
__device__ __forceinline__ void tma_store_3d(const void* tmap, const void* smem_src, int32_t c0, int32_t c1,
                                             int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes = 16) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// kind::f16, D fp32, A/B bf16, M = 128; bit 15 / 16: A / B is MN-major.
__device__ __forceinline__ constexpr uint32_t idesc(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}

// Optional timeline instrumentation (-DOSUDIT_ATTN_TRACE): CTA 0, problems 4..19; roles 0 MMA, 1 softmax, 2 epilogue.
#ifdef OSUDIT_ATTN_TRACE
__device__ long long g_bwd_trace[3 * 16 * 8];
#define BWD_TRACE(role, i, ev)                                                                  \
  do {                                                                                          \
    if (blockIdx.x == 0 && (i) >= 4 && (i) < 20) g_bwd_trace[((role) * 16 + (i) - 4) * 8 + (ev)] = clock64(); \
  } while (0)
#else
#define BWD_TRACE(role, i, ev) do {} while (0)
#endif

__global__ void __launch_bounds__(kThreads, 1) attn_bwd_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);
  uint64_t* in_full = bars + 0;   // [2] TMA: Q, K, V, dO of a problem landed
  uint64_t* in_free = bars + 2;   // [2] the gradient GEMMs that read the stage are complete
  uint64_t* sdp_full = bars + 4;  // MMA: S and dP complete
  uint64_t* p_full = bars + 5;    // softmax: P and dS written, S and dP read (8 warp arrivals)
  uint64_t* g_full = bars + 6;    // MMA: dV, dK, dQ complete (P / dS and the input stage are free again)
  uint64_t* g_free = bars + 7;    // epilogue: dV, dK, dQ read out of TMEM (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int G = static_cast<int>(gridDim.x);
  const int n_my = (p.problems - static_cast<int>(blockIdx.x) + G - 1) / G;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_qkv);
    tma_prefetch_desc(&p.tma_dout);
    tma_prefetch_desc(&p.tma_out);
    tma_prefetch_desc(&p.tma_dqkv);
    mbar_init(&in_full[0], 1);
    mbar_init(&in_full[1], 1);
    mbar_init(&in_free[0], 1);
    mbar_init(&in_free[1], 1);
    mbar_init(sdp_full, 1);
    mbar_init(p_full, 8);
    mbar_init(g_full, 1);
    mbar_init(g_free, 4);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int i = 0; i < n_my; ++i) {
        const int prob = blockIdx.x + i * G;
        const int h = prob % p.H, b = prob / p.H;
        const int st = i & 1;
        uint8_t* base = smem + st * kStage;
        mbar_wait_sleep(&in_free[st], ((i >> 1) & 1) ^ 1, 200);
        mbar_expect_tx(&in_full[st], kStage);
        tma_load_3d(base + 0 * kTile, &p.tma_qkv, &in_full[st], h * kHD, 0, b);
        tma_load_3d(base + 1 * kTile, &p.tma_qkv, &in_full[st], p.D + h * kHD, 0, b);
        tma_load_3d(base + 2 * kTile, &p.tma_qkv, &in_full[st], 2 * p.D + h * kHD, 0, b);
        tma_load_3d(base + 3 * kTile, &p.tma_dout, &in_full[st], h * kHD, 0, b);
        tma_load_3d(base + 4 * kTile, &p.tma_out, &in_full[st], h * kHD, 0, b);
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t id_s = idesc(128, false, false);   // S, dP: both operands K-major
      constexpr uint32_t id_t = idesc(kHD, true, true);     // dV, dK: A = P^T / dS^T and B both MN-major
      constexpr uint32_t id_q = idesc(kHD, false, true);    // dQ: A = dS K-major, B = K MN-major
      const uint32_t sP = smem_u32(smem + kOffP), sDS = smem_u32(smem + kOffDS);
      auto issue_sdp = [&](int i) {
        const int st = i & 1;
        const uint32_t base = smem_u32(smem + st * kStage);
        mbar_wait(&in_full[st], (i >> 1) & 1);
        tc_fence_after();
        const uint64_t dq = desc_sw128(base), dk = desc_sw128(base + kTile);
        const uint64_t dv = desc_sw128(base + 2 * kTile), ddo = desc_sw128(base + 3 * kTile);
#pragma unroll
        for (int k = 0; k < kHD / 16; ++k) umma_bf16(tmem_base + kColS, dq + 2 * k, dk + 2 * k, id_s, k != 0);
#pragma unroll
        for (int k = 0; k < kHD / 16; ++k) umma_bf16(tmem_base + kColDP, ddo + 2 * k, dv + 2 * k, id_s, k != 0);
        umma_commit(sdp_full);
      };
      if (n_my > 0) issue_sdp(0);
      for (int i = 0; i < n_my; ++i) {
        BWD_TRACE(0, i, 0);
        mbar_wait_sleep(p_full, i & 1, 32);  // S / dP of problem i read, P / dS in shared memory
        BWD_TRACE(0, i, 1);
        if (i + 1 < n_my) issue_sdp(i + 1);
        BWD_TRACE(0, i, 2);
        mbar_wait_sleep(g_free, (i & 1) ^ 1, 32);  // the epilogue has read problem i-1's gradients out of TMEM
        BWD_TRACE(0, i, 3);
        tc_fence_after();
        const uint32_t base = smem_u32(smem + (i & 1) * kStage);
        // MN-major A: keys 0-63 / 64-127 are the two column blocks of the [query][key] tile, 16 KB apart;
        // one K = 16 step = 16 query rows = 2 KB
        const uint64_t a_p = desc_sw128(sP, kTile), a_ds = desc_sw128(sDS, kTile);
        const uint64_t b_q = desc_sw128(base), b_do = desc_sw128(base + 3 * kTile);
#pragma unroll
        for (int k = 0; k < kT / 16; ++k)  // dV[key][d] = sum_q P[q][key] dO[q][d]
          umma_bf16(tmem_base + kColDV, a_p + 128 * k, b_do + 128 * k, id_t, k != 0);
#pragma unroll
        for (int k = 0; k < kT / 16; ++k)  // dK[key][d] = sum_q dS[q][key] Q[q][d]
          umma_bf16(tmem_base + kColDK, a_ds + 128 * k, b_q + 128 * k, id_t, k != 0);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {   // dQ[q][d] = sum_key dS[q][key] K[key][d]
          const uint64_t a = desc_sw128(sDS + kb * kTile);
          const uint64_t bk = desc_sw128(base + kTile + kb * (64 * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + kColDQ, a + 2 * k, bk + 128 * k, id_q, (kb | k) != 0);
        }
        umma_commit(g_full);
        BWD_TRACE(0, i, 4);
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ---------------------------------------------------- softmax recomputation: two threads per query row
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = quad * 32 + lane;  // query
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t p_row = smem_u32(smem + kOffP) + half * kTile + row * 128;   // this thread's 64-key column block
    const uint32_t ds_row = smem_u32(smem + kOffDS) + half * kTile + row * 128;
    const int swz = row & 7;
    const bool valid_q = row < p.T;
    const int c_lo = max(row - p.w_left, 0), c_hi = min(row + p.w_right, p.T - 1);  // allowed keys of this query

    float* s_delta = reinterpret_cast<float*>(smem + kOffDelta);
    for (int i = 0; i < n_my; ++i) {
      const int prob = blockIdx.x + i * G;
      const int h = prob % p.H, b = prob / p.H;
      const bool tracer = quad == 0 && lane == 0 && half == 0;
      if (tracer) BWD_TRACE(1, i, 0);
      // delta = rowsum(dO o O) from the TMA-staged tiles (a thread reading its own 128-byte row from global memory
      // touches 32 cache lines per load instruction): each half sums 32 of the 64 dims, halves meet in shared memory
      const float lse = valid_q ? __ldg(p.lse + (static_cast<int64_t>(b) * p.H + h) * p.T + row) : 0.f;
      warp_mbar_wait(&in_full[i & 1], (i >> 1) & 1);
      {
        const uint32_t t_do = smem_u32(smem + (i & 1) * kStage + 3 * kTile) + row * 128;
        const uint32_t t_o = t_do + kTile;
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = static_cast<uint32_t>(((4 * half + j) ^ swz) << 4);
          const uint4 a = lds128(t_o + off), c = lds128(t_do + off);
          const uint32_t* ua = reinterpret_cast<const uint32_t*>(&a);
          const uint32_t* uc = reinterpret_cast<const uint32_t*>(&c);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ua[k]));
            const float2 fc = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&uc[k]));
            d0 = fmaf(fa.x, fc.x, d0);
            d1 = fmaf(fa.y, fc.y, d1);
          }
        }
        s_delta[half * kT + row] = d0 + d1;
      }
      named_bar_sync(1, 256);
      const float delta = s_delta[row] + s_delta[kT + row];
      warp_mbar_wait_sleep(sdp_full, i & 1, 20);
      if (tracer) BWD_TRACE(1, i, 1);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c = 2 * half + cc;  // 32-key chunk of the row
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(t_lane + kColS + c * 32, sv);
        tmem_ld_32x32(t_lane + kColDP + c * 32, dv);
        tmem_ld_wait();
        uint32_t pp[16], pd[16];
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          float pv[2], dsv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = c * 32 + k + e;
            const bool ok = valid_q && col >= c_lo && col <= c_hi;
            const float pe = ok ? fast_exp2(fmaf(__uint_as_float(sv[k + e]), p.scale_log2, -lse)) : 0.f;
            pv[e] = pe;
            dsv[e] = pe * (__uint_as_float(dv[k + e]) - delta);
          }
          pp[k >> 1] = pack_bf16(pv[0], pv[1]);
          pd[k >> 1] = pack_bf16(dsv[0], dsv[1]);
        }
        if (tracer) BWD_TRACE(1, i, 2 + 2 * cc);
        if (cc == 0 && i > 0) warp_mbar_wait(g_full, (i - 1) & 1);  // problem i-1's GEMMs still read P / dS
        if (tracer) BWD_TRACE(1, i, 3 + 2 * cc);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {  // 32 keys = 64 B = four 16-byte chunks of the 128-byte row
          const int chunk = cc * 4 + jj;
          sts128(p_row + ((chunk ^ swz) << 4), pp[4 * jj], pp[4 * jj + 1], pp[4 * jj + 2], pp[4 * jj + 3]);
          sts128(ds_row + ((chunk ^ swz) << 4), pd[4 * jj], pd[4 * jj + 1], pd[4 * jj + 2], pd[4 * jj + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      warp_mbar_arrive(p_full);
      if (tracer) BWD_TRACE(1, i, 6);
      named_bar_sync(1, 256);  // every partial delta of this problem has been read before the next one is written
    }
  } else if (warp >= 12) {
    // ------------------------------------------------------------------ epilogue: dQ, dK, dV -> bf16 -> dqkv
    // The rows go through the problem's own Q / K / V tiles (free once g_full has fired) and leave with three TMA
    // stores: per-thread 128-byte row stores to global memory cost 32 cache lines per instruction.
    const int quad = warp & 3;
    const int row = quad * 32 + lane;  // query row of dQ, key row of dK / dV
    const int swz = row & 7;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    for (int i = 0; i < n_my; ++i) {
      const int prob = blockIdx.x + i * G;
      const int h = prob % p.H, b = prob / p.H;
      const bool tracer = warp == 12 && lane == 0;
      if (tracer) BWD_TRACE(2, i, 0);
      warp_mbar_wait_sleep(g_full, i & 1, 40);
      if (tracer) BWD_TRACE(2, i, 1);
      tc_fence_after();
      uint8_t* stage = smem + (i & 1) * kStage;
#pragma unroll 1
      for (int m = 0; m < 3; ++m) {  // 0: dQ (x scale) -> Q tile, 1: dK (x scale) -> K tile, 2: dV -> V tile
        const uint32_t col = m == 0 ? kColDQ : (m == 1 ? kColDK : kColDV);
        const float f = m == 2 ? 1.0f : p.scale;
        const uint32_t dst = smem_u32(stage + m * kTile) + row * 128;
        uint32_t lo[32], hi[32];
        tmem_ld_32x32(t_lane + col, lo);
        tmem_ld_32x32(t_lane + col + 32, hi);
        tmem_ld_wait();
        if (m == 2) {  // everything read: the next problem's gradient GEMMs may overwrite TMEM
          tc_fence_before();
          warp_mbar_arrive(g_free);
          if (tracer) BWD_TRACE(2, i, 2);
        }
        auto a = [&](const uint32_t(&r)[32], int k) { return __uint_as_float(r[k]) * f; };
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          sts128(dst + ((jj ^ swz) << 4), pack_bf16(a(lo, 8 * jj + 0), a(lo, 8 * jj + 1)), pack_bf16(a(lo, 8 * jj + 2), a(lo, 8 * jj + 3)),
                 pack_bf16(a(lo, 8 * jj + 4), a(lo, 8 * jj + 5)), pack_bf16(a(lo, 8 * jj + 6), a(lo, 8 * jj + 7)));
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          sts128(dst + (((4 + jj) ^ swz) << 4), pack_bf16(a(hi, 8 * jj + 0), a(hi, 8 * jj + 1)), pack_bf16(a(hi, 8 * jj + 2), a(hi, 8 * jj + 3)),
                 pack_bf16(a(hi, 8 * jj + 4), a(hi, 8 * jj + 5)), pack_bf16(a(hi, 8 * jj + 6), a(hi, 8 * jj + 7)));
      }
      fence_proxy_async_smem();
      named_bar_sync(2, 128);
      if (warp == 12 && lane == 0) {  // three TMA stores (rows >= T are clipped by the tensor map)
        tma_store_3d(&p.tma_dqkv, stage + 0 * kTile, h * kHD, 0, b);
        tma_store_3d(&p.tma_dqkv, stage + 1 * kTile, p.D + h * kHD, 0, b);
        tma_store_3d(&p.tma_dqkv, stage + 2 * kTile, 2 * p.D + h * kHD, 0, b);
        tma_store_commit();
      }
      // while the stores drain: the in_proj_bias gradient = column sums of the staged (bf16) rows; 8 lanes per row,
      // 8 rows per thread and tile, then lanes l, l+8, l+16, l+24 (same columns) are folded and 8 lanes add to global
      if (p.dbias != nullptr) {
        const int et = threadIdx.x - 12 * 32;
        const int piece = et & 7, row0 = et >> 3;
#pragma unroll 1
        for (int m = 0; m < 3; ++m) {
          const uint32_t src = smem_u32(stage + m * kTile);
          float cs[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) cs[k] = 0.f;
#pragma unroll
          for (int r8 = 0; r8 < 8; ++r8) {
            const int r = row0 + 16 * r8;
            if (r < p.T) {
              const uint4 v = lds128(src + r * 128 + ((piece ^ (r & 7)) << 4));
              const uint32_t* u = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u[k]));
                cs[2 * k] += f.x;
                cs[2 * k + 1] += f.y;
              }
            }
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 8);
            cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 16);
          }
          if (lane < 8) {
            float* db = p.dbias + m * p.D + h * kHD + piece * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k) atomicAdd(db + k, cs[k]);
          }
        }
      }
      named_bar_sync(2, 128);  // every column sum has been read out of the tiles
      if (warp == 12 && lane == 0) {
        tma_store_wait_read<0>();  // ... and so have the stores: the stage may be refilled
        mbar_arrive(&in_free[i & 1]);
        if (tracer) BWD_TRACE(2, i, 3);
      }
    }
    if (warp == 12 && lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace attn_bwd_tc

#ifdef OSUDIT_ATTN_TRACE
extern "C" int osudit_debug_bwd_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, attn_bwd_tc::g_bwd_trace, sizeof(attn_bwd_tc::g_bwd_trace)) == cudaSuccess ? 0 : -1;
}
#endif

bool attn_bwd_tc_applicable(int T, int head_dim) { return head_dim == 64 && T <= attn_bwd_tc::kT; }

int attn_bwd_tc_launch(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B, int T,
                       int H, int w_left, int w_right, float* dbias, cudaStream_t stream) {
  using namespace attn_bwd_tc;
  Params p;
  const int D = H * kHD;
  int rc = make_tensor_map_3d(&p.tma_qkv, qkv, 3ull * D, T, B, 3ull * D * 2, 3ull * D * 2 * T, kHD, kT);
  if (rc) return rc;
  rc = make_tensor_map_3d(&p.tma_dout, dout, 1ull * D, T, B, 1ull * D * 2, 1ull * D * 2 * T, kHD, kT);
  if (rc) return rc;
  rc = make_tensor_map_3d(&p.tma_out, out, 1ull * D, T, B, 1ull * D * 2, 1ull * D * 2 * T, kHD, kT);
  if (rc) return rc;
  rc = make_tensor_map_3d(&p.tma_dqkv, dqkv, 3ull * D, T, B, 3ull * D * 2, 3ull * D * 2 * T, kHD, kT);
  if (rc) return rc;
  p.lse = lse;
  p.dbias = dbias;
  p.B = B; p.T = T; p.H = H; p.D = D;
  p.problems = B * H;
  p.w_left = w_left;
  p.w_right = w_right;
  p.scale = 1.0f / sqrtf(static_cast<float>(kHD));
  p.scale_log2 = 1.4426950408889634f * p.scale;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return set_error(-5, cudaGetErrorString(e));
    configured = true;
  }
  const int grid = p.problems < num_sms() ? p.problems : num_sms();
  attn_bwd_tc_kernel<<<grid, kThreads, kSmemBytes, stream>>>(p);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

}  // namespace osudit
