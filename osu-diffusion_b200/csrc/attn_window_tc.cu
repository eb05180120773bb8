// Windowed self-attention on tcgen05: one CTA step = 128 queries of one (batch, head) against the
// 384-key window [q0-128, q0+256) that contains every key the band (or a short full sequence) allows.
//
// Reference: nn.MultiheadAttention core inside DiTBlock (models.py:164-170) under the band mask of
// sample.py:81-84 (query j sees key i iff -(W-1) <= i-j <= W, W=128) or no mask for T <= 256
// (training windows, train.py:325).  Because the whole allowed key range of a 128-query tile fits in
// one 384-key window, softmax is exact in a single pass structure (row max, then exp/sum): no
// online rescaling of the output accumulator.
//
//   warp 0   TMA producer: Q tile + 3 K tiles + 3 V tiles as 128-byte-swizzled [128][64] bf16 boxes
//            cut from packed qkv [B, T, 3D] with a 3-D tensor map (rows outside [0,T) are zero-filled
//            per batch by TMA; they are masked to -inf before the softmax anyway)
//   warp 1   tcgen05.mma issuer: S[128x384] = Q K^T into TMEM (3 x N=128, K=64), later
//            O[128x64] = P V (24 x K=16; V is consumed as an MN-major B operand straight from the
//            [key][dim] tile, P as a K-major A operand from shared memory)
//   warps 2-5 one thread per query row: tcgen05.ld the S row, band/tail mask, max, exp2, sum — all
//            thread-local, no shuffles — write P as bf16 into the swizzled A-operand layout, then
//            normalise O and store it
// TMEM: 384 columns of S + 64 of O.  Shared memory: Q 16K + K 48K + V 48K + P 96K.
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

namespace attn_tc {

constexpr int kQ = 128;        // queries per tile
constexpr int kWin = 384;      // keys per window
constexpr int kHD = 64;
constexpr int kTile = kQ * kHD * 2;            // 16 KB: one [128][64] bf16 box
constexpr int kSmemQ = 0;
constexpr int kSmemK = kSmemQ + kTile;         // 3 boxes
constexpr int kSmemV = kSmemK + 3 * kTile;     // 3 boxes
constexpr int kSmemP = kSmemV + 3 * kTile;     // 6 boxes: P[128][384] as 6 K-blocks of 64 keys
constexpr int kSmemBar = kSmemP + 6 * kTile;
constexpr int kSmemBytes = kSmemBar + 256 + 1024;
constexpr int kThreads = 192;
constexpr uint32_t kColS = 0, kColO = kWin;    // TMEM columns

struct Params {
  CUtensorMap tma_qkv;  // 3-D: [3D cols, T, B], box [64, 128, 1]
  __nv_bfloat16* out;   // [B*T, D]
  int B, T, H, D;
  int q_tiles, total_tiles;
  int lo, hi;           // allowed iff lo <= (col - row) <= hi with col = key - (q0 - 128), row = q - q0
  float scale_log2;
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar,
                                            int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// SWIZZLE_128B smem descriptor, 8-row groups 1024 B apart.  Serves K-major operands (rows = M/N
// index, 128 B of K per row) and the MN-major V operand (rows = K index, 128 B of N per row).
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// kind::f16, D fp32, A/B bf16, M=128; b_mn_major selects an MN-major B operand.
__device__ __forceinline__ constexpr uint32_t idesc(int n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}

__global__ void __launch_bounds__(kThreads, 1) attn_window_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);
  uint64_t* qk_full = bars + 0;   // TMA: Q + 3 K landed
  uint64_t* v_full = bars + 1;    // TMA: 3 V landed
  uint64_t* s_done = bars + 2;    // MMA: S complete (Q/K smem reusable)
  uint64_t* o_done = bars + 3;    // MMA: O complete (V/P smem reusable)
  uint64_t* s_free = bars + 4;    // softmax: S fully read (128 arrivals)
  uint64_t* p_full = bars + 5;    // softmax: P written (128 arrivals)
  uint64_t* o_free = bars + 6;    // epilogue: O fully read (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_qkv);
    mbar_init(qk_full, 1);
    mbar_init(v_full, 1);
    mbar_init(s_done, 1);
    mbar_init(o_done, 1);
    mbar_init(s_free, 128);
    mbar_init(p_full, 128);
    mbar_init(o_free, 128);
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
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int qt = tile % p.q_tiles;
        const int bh = tile / p.q_tiles;
        const int h = bh % p.H;
        const int b = bh / p.H;
        const int q0 = qt * kQ;
        const int kw0 = q0 - kQ;
        const uint32_t par = (it & 1) ^ 1;  // wait for the previous tile's completion
        mbar_wait(s_done, par);
        mbar_expect_tx(qk_full, 4 * kTile);
        tma_load_3d(smem + kSmemQ, &p.tma_qkv, qk_full, h * kHD, q0, b);
#pragma unroll
        for (int j = 0; j < 3; ++j)
          tma_load_3d(smem + kSmemK + j * kTile, &p.tma_qkv, qk_full, p.D + h * kHD, kw0 + j * kQ, b);
        mbar_wait(o_done, par);
        mbar_expect_tx(v_full, 3 * kTile);
#pragma unroll
        for (int j = 0; j < 3; ++j)
          tma_load_3d(smem + kSmemV + j * kTile, &p.tma_qkv, v_full, 2 * p.D + h * kHD, kw0 + j * kQ, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = idesc(128, false);
      constexpr uint32_t idesc_o = idesc(kHD, true);
      const uint32_t sq = smem_u32(smem + kSmemQ);
      const uint32_t sk = smem_u32(smem + kSmemK);
      const uint32_t sv = smem_u32(smem + kSmemV);
      const uint32_t sp = smem_u32(smem + kSmemP);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const uint32_t par = it & 1;
        // ---- S = Q K^T
        mbar_wait(qk_full, par);
        mbar_wait(s_free, par ^ 1);
        tc_fence_after();
        const uint64_t dq = desc_sw128(sq);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const uint64_t dk = desc_sw128(sk + j * kTile);
#pragma unroll
          for (int k = 0; k < kHD / 16; ++k)
            umma_bf16(tmem_base + kColS + j * 128, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        }
        umma_commit(s_done);
        // ---- O = P V
        mbar_wait(p_full, par);
        mbar_wait(v_full, par);
        mbar_wait(o_free, par ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 6; ++kb) {       // 64-key blocks of P; V box j = kb / 2
          const uint64_t dp = desc_sw128(sp + kb * kTile);
          const uint64_t dv = desc_sw128(sv + (kb >> 1) * kTile + (kb & 1) * (64 * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k)          // 16 keys: +32 B along P's row, +16 rows (2048 B) in V
            umma_bf16(tmem_base + kColO, dp + 2 * k, dv + 128 * k, idesc_o, (kb | k) != 0);
        }
        umma_commit(o_done);
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    uint8_t* p_row = smem + kSmemP + row * 128;
    const int swz = row & 7;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int qt = tile % p.q_tiles;
      const int bh = tile / p.q_tiles;
      const int h = bh % p.H;
      const int b = bh / p.H;
      const int q0 = qt * kQ;
      const int kw0 = q0 - kQ;
      const uint32_t par = it & 1;
      // per-row allowed column range [c_lo, c_hi] inside the window
      int c_lo = max(row + p.lo, -kw0);
      int c_hi = min(row + p.hi, p.T - 1 - kw0);
      c_lo = max(c_lo, 0);
      c_hi = min(c_hi, kWin - 1);
      // warp-uniform chunk range that contains any allowed column of any row of this warp
      const int w_lo = max(max(quad * 32 + p.lo, -kw0), 0);
      const int w_hi = min(min(quad * 32 + 31 + p.hi, p.T - 1 - kw0), kWin - 1);
      const int ch_lo = w_lo >> 5, ch_hi = w_hi >> 5;

      mbar_wait(s_done, par);
      tc_fence_after();
      // ---- pass 1: row max
      float mx = -INFINITY;
      for (int c = ch_lo; c <= ch_hi; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_lane + kColS + c * 32, r);
        tmem_ld_wait();
        const int base = c * 32;
        if (base >= c_lo && base + 31 <= c_hi) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (base + i >= c_lo && base + i <= c_hi) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      const float moff = (mx == -INFINITY ? 0.f : mx) * p.scale_log2;
      // ---- pass 2: p = exp2(s*scale - max*scale), row sum, P -> smem (bf16, A-operand layout)
      float sum = 0.f;
      for (int c = 0; c < kWin / 32; ++c) {
        uint32_t packed[16];
        if (c < ch_lo || c > ch_hi) {
#pragma unroll
          for (int i = 0; i < 16; ++i) packed[i] = 0u;
        } else {
          uint32_t r[32];
          tmem_ld_32x32(t_lane + kColS + c * 32, r);
          tmem_ld_wait();
          const int base = c * 32;
          const bool inner = base >= c_lo && base + 31 <= c_hi;
          float pv[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float e = fast_exp2(fmaf(__uint_as_float(r[i]), p.scale_log2, -moff));
            if (!inner && (base + i < c_lo || base + i > c_hi)) e = 0.f;
            pv[i] = e;
            sum += e;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) packed[i] = pack_bf16(pv[2 * i], pv[2 * i + 1]);
        }
        // 32 keys = 64 B = four 16-byte chunks of the 128-byte row in K-block c/2
        uint8_t* blk = p_row + (c >> 1) * kTile;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int chunk = (c & 1) * 4 + j;
          *reinterpret_cast<uint4*>(blk + ((chunk ^ swz) << 4)) =
              make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(s_free);
      fence_proxy_async_smem();
      mbar_arrive(p_full);

      // ---- epilogue: O / sum -> bf16 -> global
      mbar_wait(o_done, par);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32(t_lane + kColO, o0);
      tmem_ld_32x32(t_lane + kColO + 32, o1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(o_free);
      const int q = q0 + row;
      if (q < p.T) {
        const float inv = 1.0f / sum;
        uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<int64_t>(b) * p.T + q) * p.D + h * kHD);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = make_uint4(
              pack_bf16(__uint_as_float(o0[8 * j + 0]) * inv, __uint_as_float(o0[8 * j + 1]) * inv),
              pack_bf16(__uint_as_float(o0[8 * j + 2]) * inv, __uint_as_float(o0[8 * j + 3]) * inv),
              pack_bf16(__uint_as_float(o0[8 * j + 4]) * inv, __uint_as_float(o0[8 * j + 5]) * inv),
              pack_bf16(__uint_as_float(o0[8 * j + 6]) * inv, __uint_as_float(o0[8 * j + 7]) * inv));
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[4 + j] = make_uint4(
              pack_bf16(__uint_as_float(o1[8 * j + 0]) * inv, __uint_as_float(o1[8 * j + 1]) * inv),
              pack_bf16(__uint_as_float(o1[8 * j + 2]) * inv, __uint_as_float(o1[8 * j + 3]) * inv),
              pack_bf16(__uint_as_float(o1[8 * j + 4]) * inv, __uint_as_float(o1[8 * j + 5]) * inv),
              pack_bf16(__uint_as_float(o1[8 * j + 6]) * inv, __uint_as_float(o1[8 * j + 7]) * inv));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace attn_tc

// True when every key a 128-query tile may attend lies inside its [q0-128, q0+256) window.
bool attn_window_applicable(int T, int head_dim, int w_left, int w_right, const uint8_t* mask) {
  if (head_dim != 64 || mask != nullptr) return false;
  if (w_left <= 128 && w_right <= 128) return true;  // band
  return T <= 256;                                   // short full sequence: window covers [0, T)
}

int attn_window_launch(const void* qkv, void* out, int B, int T, int H, int w_left, int w_right,
                       cudaStream_t stream) {
  using namespace attn_tc;
  Params p;
  const int D = H * kHD;
  int rc = make_tensor_map_3d(&p.tma_qkv, qkv, 3ull * D, T, B, 3ull * D * 2, 3ull * D * 2 * T, kHD, kQ);
  if (rc) return rc;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.B = B; p.T = T; p.H = H; p.D = D;
  p.q_tiles = (T + kQ - 1) / kQ;
  p.total_tiles = p.q_tiles * H * B;
  // key - query in [-w_left, w_right]  <=>  col - row in [128 - w_left, 128 + w_right]
  p.lo = kQ - w_left;
  p.hi = kQ + w_right;
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(kHD));
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemBytes);
    if (e != cudaSuccess) return set_error(-5, cudaGetErrorString(e));
    configured = true;
  }
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  attn_window_kernel<<<grid, kThreads, kSmemBytes, stream>>>(p);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

}  // namespace osudit
