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
//   warps 2-5 one thread per query row: tcgen05.ld the S row ONCE (pipelined 32-column chunks),
//            band/tail mask, exp2 against a lazily updated power-of-two reference, sum — all
//            thread-local, no shuffles — write P as bf16 into the swizzled A-operand layout, then
//            normalise O and store it.  (TMEM->register bandwidth, ~64 B/clk/SM, and MUFU exp2
//            bound this kernel for head_dim 64, so S must not be read twice.)
// Software pipeline: the MMA warp issues S(i+1) slab by slab while tile i is exponentiated (each
// 128-column slab of S is released as soon as it has been read), then PV(i).
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
constexpr int kSmemX = kSmemBar + 256;         // (reference, sum) exchange between the two column halves
constexpr int kSmemBytes = kSmemX + 2 * 128 * 8 + 1024;
constexpr int kSoftmaxThreads = 256;          // 8 warps: two per TMEM lane quadrant, 192 columns each
constexpr int kThreads = 64 + kSoftmaxThreads;
constexpr uint32_t kColS = 0, kColO = kWin;    // TMEM columns: S 0-383, O0 384-447, O1 448-511

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

struct TileInfo {
  int b, h, q0, kw0;
};

__device__ __forceinline__ TileInfo decode_tile(const Params& p, int tile) {
  TileInfo t;
  const int qt = tile % p.q_tiles;
  const int bh = tile / p.q_tiles;
  t.h = bh % p.H;
  t.b = bh / p.H;
  t.q0 = qt * kQ;
  t.kw0 = t.q0 - kQ;
  return t;
}

// The running softmax reference may lag the true row maximum by up to 2^kJump before the row's
// already-written probabilities are rescaled (by an exact power of two): P then stays <= 2^kJump,
// far inside bf16/fp32 range, and the final O / sum is independent of the reference.
constexpr float kJump = 24.0f;

// Optional timeline instrumentation (build with -DOSUDIT_ATTN_TRACE): CTA 0 records clock64() at the hand-off
// points of tiles 8..23 for the MMA warp (role 0), softmax half 0 (role 1) and half 1 (role 2);
// osudit_debug_attn_trace() copies the table out.  Used to find the critical path, never in the shipped build.
#ifdef OSUDIT_ATTN_TRACE
__device__ long long g_trace[3 * 16 * 8];
__device__ long long g_trace_chunks[16 * 16];  // half 0, quadrant 0: per chunk (after the TMEM wait, after emit)
#define ATTN_TRACE_CHUNK(i, ev)                                                                      \
  do {                                                                                               \
    if (blockIdx.x == 0 && (i) >= 8 && (i) < 24) g_trace_chunks[((i) - 8) * 16 + (ev)] = clock64(); \
  } while (0)
#define ATTN_TRACE(role, i, ev)                                                     \
  do {                                                                              \
    if (blockIdx.x == 0 && (i) >= 8 && (i) < 24) g_trace[((role) * 16 + (i) - 8) * 8 + (ev)] = clock64(); \
  } while (0)
#else
#define ATTN_TRACE(role, i, ev) do {} while (0)
#define ATTN_TRACE_CHUNK(i, ev) do {} while (0)
#endif

__global__ void __launch_bounds__(kThreads, 1) attn_window_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);
  uint64_t* qk_full = bars + 0;   // TMA: Q + 3 K landed
  uint64_t* v_full = bars + 1;    // TMA: 3 V landed
  uint64_t* s01_done = bars + 2;  // MMA: slabs 0,1 of S complete (what half 0 reads)
  uint64_t* s_done = bars + 3;    // MMA: all of S complete (half 1 may start; Q/K smem reusable)
  uint64_t* o_done = bars + 4;    // [2] MMA: O_h complete (P_h smem reusable; after [1]: V reusable)
  uint64_t* o_free = bars + 6;    // epilogue: O0 and O1 fully read (128 arrivals, half 1)
  uint64_t* s_free = bars + 7;    // [3] 128-column slab j of S fully read
  uint64_t* p_full = bars + 10;   // [2] P_h written (128 arrivals: the owning half)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  float2* xchg = reinterpret_cast<float2*>(smem + kSmemX);  // [2 parities][128 rows] half 0's (ref, sum)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_qkv);
    mbar_init(qk_full, 1);
    mbar_init(v_full, 1);
    mbar_init(s01_done, 1);
    mbar_init(s_done, 1);
    mbar_init(&o_done[0], 1);
    mbar_init(&o_done[1], 1);
    mbar_init(o_free, 128);
    mbar_init(&s_free[0], 128);  // columns   0-127: half 0 only
    mbar_init(&s_free[1], 256);  // columns 128-255: both halves
    mbar_init(&s_free[2], 128);  // columns 256-383: half 1 only
    mbar_init(&p_full[0], 128);
    mbar_init(&p_full[1], 128);
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
  const int n_my = (p.total_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                   static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int i = 0; i < n_my; ++i) {
        const TileInfo t = decode_tile(p, blockIdx.x + i * gridDim.x);
        const uint32_t prev = (i & 1) ^ 1;  // parity of tile i-1's completion (passes at i == 0)
        mbar_wait(s_done, prev);            // S(i-1) done: Q/K smem free
        mbar_expect_tx(qk_full, 4 * kTile);
        tma_load_3d(smem + kSmemQ, &p.tma_qkv, qk_full, t.h * kHD, t.q0, t.b);
#pragma unroll
        for (int j = 0; j < 3; ++j)
          tma_load_3d(smem + kSmemK + j * kTile, &p.tma_qkv, qk_full, p.D + t.h * kHD,
                      t.kw0 + j * kQ, t.b);
        mbar_wait(&o_done[1], prev);        // PV_1(i-1), the last reader of V(i-1), is done
        mbar_expect_tx(v_full, 3 * kTile);
#pragma unroll
        for (int j = 0; j < 3; ++j)
          tma_load_3d(smem + kSmemV + j * kTile, &p.tma_qkv, v_full, 2 * p.D + t.h * kHD,
                      t.kw0 + j * kQ, t.b);
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    // The two column halves of the softmax run staggered (half 0 starts as soon as slabs 0,1 of S
    // exist, half 1 after slab 2), so the tensor pipe interleaves, per tile i:
    //   S(i+1) slab 0 | PV_0(i) | S(i+1) slab 1 | S(i+1) slab 2 | PV_1(i)
    // and each half's MMA latency is hidden behind the other half's exponentials.
    if (lane == 0) {
      constexpr uint32_t idesc_s = idesc(128, false);
      constexpr uint32_t idesc_o = idesc(kHD, true);
      const uint32_t sq = smem_u32(smem + kSmemQ);
      const uint32_t sk = smem_u32(smem + kSmemK);
      const uint32_t sv = smem_u32(smem + kSmemV);
      const uint32_t sp = smem_u32(smem + kSmemP);
      const uint64_t dq = desc_sw128(sq);
      auto issue_s_slab = [&](int i, int j) {  // S(i)[:, 128j : 128j+128] = Q K_j^T
        mbar_wait(&s_free[j], (i & 1) ^ 1);
        tc_fence_after();
        const uint64_t dk = desc_sw128(sk + j * kTile);
#pragma unroll
        for (int k = 0; k < kHD / 16; ++k)
          umma_bf16(tmem_base + kColS + j * 128, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
      };
      auto issue_pv_half = [&](int i, int h) {  // O_h = P[:, 192h : 192h+192] V[192h : 192h+192]
        mbar_wait(&p_full[h], i & 1);
        tc_fence_after();
#pragma unroll
        for (int b3 = 0; b3 < 3; ++b3) {
          const int kb = 3 * h + b3;           // 64-key block of P; V box = kb / 2
          const uint64_t dp = desc_sw128(sp + kb * kTile);
          const uint64_t dv = desc_sw128(sv + (kb >> 1) * kTile + (kb & 1) * (64 * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k)          // 16 keys: +32 B along P's row, +16 rows (2048 B) in V
            umma_bf16(tmem_base + kColO + h * kHD, dp + 2 * k, dv + 128 * k, idesc_o, (b3 | k) != 0);
        }
        umma_commit(&o_done[h]);
      };
      if (n_my > 0) {
        mbar_wait(qk_full, 0);
        issue_s_slab(0, 0);
        issue_s_slab(0, 1);
        umma_commit(s01_done);
        issue_s_slab(0, 2);
        umma_commit(s_done);
      }
      for (int i = 0; i < n_my; ++i) {
        const bool has_next = i + 1 < n_my;
        ATTN_TRACE(0, i, 0);
        if (has_next) {
          mbar_wait(qk_full, (i + 1) & 1);
          ATTN_TRACE(0, i, 1);
          issue_s_slab(i + 1, 0);
        }
        ATTN_TRACE(0, i, 2);
        mbar_wait(v_full, i & 1);
        mbar_wait(o_free, (i & 1) ^ 1);
        ATTN_TRACE(0, i, 3);
        issue_pv_half(i, 0);
        ATTN_TRACE(0, i, 4);
        if (has_next) {
          issue_s_slab(i + 1, 1);
          umma_commit(s01_done);
          ATTN_TRACE(0, i, 5);
          issue_s_slab(i + 1, 2);
          umma_commit(s_done);
        }
        ATTN_TRACE(0, i, 6);
        issue_pv_half(i, 1);
        ATTN_TRACE(0, i, 7);
      }
    }
  } else {
    // ---------------------- softmax + epilogue: two threads per query row, 192 window columns each
    // S is read from TMEM exactly once; each 32-column chunk is exponentiated against a running
    // integer reference (log2 domain) private to the thread.  The two halves of a row never
    // reconcile their references in P: they accumulate into separate O accumulators (O0, O1) and
    // the epilogue (done by half 1, which finishes last) combines them as
    // (2^r0 O0 + 2^r1 O1) / (2^r0 sum0 + 2^r1 sum1).
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t p_row = smem_u32(smem + kSmemP) + row * 128;  // shared-space address of this row of P
    const int swz = row & 7;
    const int cbeg = half * 6;  // this thread's chunks: [cbeg, cbeg + 6)

    for (int i = 0; i < n_my; ++i) {
      const TileInfo t = decode_tile(p, blockIdx.x + i * gridDim.x);
      const uint32_t par = i & 1;
      // allowed columns of this row, and the chunk range any row of this warp needs
      const int c_lo = max(max(row + p.lo, -t.kw0), 0);
      const int c_hi = min(min(row + p.hi, p.T - 1 - t.kw0), kWin - 1);
      const int ch_lo = max(max(quad * 32 + p.lo, -t.kw0), 0) >> 5;
      const int ch_hi = min(min(quad * 32 + 31 + p.hi, p.T - 1 - t.kw0), kWin - 1) >> 5;
      auto in_range = [&](int c) { return c >= ch_lo && c <= ch_hi; };

      float ref = -INFINITY;  // running reference (integer-valued, log2 domain)
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;

      auto store_chunk = [&](int c, const uint32_t(&packed)[16]) {
        // 32 keys = 64 B = four 16-byte chunks of the 128-byte row in K-block c/2
        const uint32_t blk = p_row + (c >> 1) * kTile;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int chunk = (c & 1) * 4 + j;
          sts128(blk + ((chunk ^ swz) << 4), packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        }
      };
      // rare: the reference moved up by more than kJump: rescale what this thread already wrote
      // (nothing of it has been published to the tensor pipe yet: p_full[half] fires after the
      // thread's last chunk)
      auto rescale_written = [&](int c_end, float factor) {
        for (int cc = cbeg; cc < c_end; ++cc) {
          const uint32_t blk = p_row + (cc >> 1) * kTile;
          for (int j = 0; j < 4; ++j) {
            const int chunk = (cc & 1) * 4 + j;
            const uint32_t addr = blk + ((chunk ^ swz) << 4);
            uint4 w = lds128(addr);
            uint32_t* e = reinterpret_cast<uint32_t*>(&w);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&e[k]));
              e[k] = pack_bf16(f.x * factor, f.y * factor);
            }
            sts128(addr, w.x, w.y, w.z, w.w);
          }
        }
        sum0 *= factor; sum1 *= factor; sum2 *= factor; sum3 *= factor;
      };
      auto emit = [&](uint32_t(&v)[32], int c) {
        uint32_t packed[16];
        if (!in_range(c)) {
#pragma unroll
          for (int k = 0; k < 16; ++k) packed[k] = 0u;
          store_chunk(c, packed);
          return;
        }
        const int base = c * 32;
        if (!(base >= c_lo && base + 31 <= c_hi)) {  // boundary chunk: masked scores become -inf
          const int klo = c_lo - base, khi = c_hi - base;
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (k < klo || k > khi) v[k] = 0xff800000u;
        }
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          m0 = fmaxf(m0, __uint_as_float(v[k]));
          m1 = fmaxf(m1, __uint_as_float(v[k + 1]));
        }
        const float cm = fmaxf(m0, m1) * p.scale_log2;
        if (cm > ref + kJump) {  // also taken on the thread's first allowed chunk (ref = -inf)
          const float new_ref = ceilf(cm);
          if (ref != -INFINITY) rescale_written(c, fast_exp2(ref - new_ref));
          ref = new_ref;
        }
        const float off = (ref == -INFINITY) ? 0.f : ref;
#pragma unroll
        for (int g = 0; g < 2; ++g) {  // two groups of 16 columns keep the live register set small
          float pv[16];
#pragma unroll
          for (int k = 0; k < 16; ++k)
            pv[k] = fast_exp2(fmaf(__uint_as_float(v[16 * g + k]), p.scale_log2, -off));
#pragma unroll
          for (int k = 0; k < 16; k += 4) {
            sum0 += pv[k];
            sum1 += pv[k + 1];
            sum2 += pv[k + 2];
            sum3 += pv[k + 3];
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) packed[8 * g + k] = pack_bf16(pv[2 * k], pv[2 * k + 1]);
        }
        store_chunk(c, packed);
      };

      // S slabs this half reads exist; PV_half(i-1) has finished reading this half's P blocks
      const bool tracer = (quad == 0 && lane == 0);
      if (tracer) ATTN_TRACE(1 + half, i, 0);
      mbar_wait(half == 0 ? s01_done : s_done, par);
      if (tracer) ATTN_TRACE(1 + half, i, 1);
      if (i > 0) mbar_wait(&o_done[half], par ^ 1);
      tc_fence_after();
      if (tracer) ATTN_TRACE(1 + half, i, 2);

      uint32_t ra[32], rb[32];
      if (in_range(cbeg)) tmem_ld_32x32(t_lane + kColS + cbeg * 32, ra);
#pragma unroll 1  // keep the body (2 x emit) resident in the instruction cache
      for (int u = 0; u < 6; u += 2) {
        const int c = cbeg + u;
        tmem_ld_wait();
        if (tracer && half == 0) ATTN_TRACE_CHUNK(i, 2 * u);
        if (in_range(c + 1)) tmem_ld_32x32(t_lane + kColS + (c + 1) * 32, rb);
        emit(ra, c);
        if (tracer && half == 0) ATTN_TRACE_CHUNK(i, 2 * u + 1);
        tmem_ld_wait();
        if (tracer && half == 0) ATTN_TRACE_CHUNK(i, 2 * u + 2);
        if (u + 2 < 6 && in_range(c + 2)) tmem_ld_32x32(t_lane + kColS + (c + 2) * 32, ra);
        emit(rb, c + 1);
        if (tracer && half == 0) ATTN_TRACE_CHUNK(i, 2 * u + 3);
        // release slabs of S: half 0 owns chunks 0-5 (slab 0, first half of slab 1), half 1 6-11
        if ((half == 0 && u == 2) || (half == 1 && u == 0) || u == 4) {
          const int slab = (half == 0) ? (u == 2 ? 0 : 1) : (u == 0 ? 1 : 2);
          tc_fence_before();
          mbar_arrive(&s_free[slab]);
        }
      }
      const float sum = (sum0 + sum1) + (sum2 + sum3);
      if (tracer) ATTN_TRACE(1 + half, i, 3);
      float2* xc = xchg + par * 128;
      if (half == 0) {
        xc[row] = make_float2(ref, sum);
        fence_proxy_async_smem();
        mbar_arrive(&p_full[0]);
        // publishes xc to half 1 (which bar.syncs); ids alternate per tile because half 0 may
        // already be one tile ahead of half 1 (never two: s_free[1] ties them together)
        asm volatile("bar.arrive %0, 256;" ::"r"(1 + static_cast<int>(par)) : "memory");
        continue;
      }
      fence_proxy_async_smem();
      mbar_arrive(&p_full[1]);
      named_bar_sync(1 + par, kSoftmaxThreads);
      const float2 other = xc[row];

      // ---- epilogue (half 1): (w0 O0 + w1 O1) / (w0 sum0 + w1 sum1) -> bf16 -> global
      if (tracer) ATTN_TRACE(2, i, 4);
      mbar_wait(&o_done[0], par);
      mbar_wait(&o_done[1], par);
      tc_fence_after();
      if (tracer) ATTN_TRACE(2, i, 5);
      const float rmax = fmaxf(other.x, ref);
      const float w_a = (other.x == -INFINITY) ? 0.f : fast_exp2(other.x - rmax);
      const float w_b = (ref == -INFINITY) ? 0.f : fast_exp2(ref - rmax);
      const float inv = 1.0f / (w_a * other.y + w_b * sum);
      const float ka = w_a * inv, kb2 = w_b * inv;
      const int q = t.q0 + row;
      uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<int64_t>(t.b) * p.T + q) * p.D + t.h * kHD);
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {  // two groups of 32 head dims
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(t_lane + kColO + hh * 32, o0);
        tmem_ld_32x32(t_lane + kColO + kHD + hh * 32, o1);
        tmem_ld_wait();
        if (hh == 1) {
          tc_fence_before();
          mbar_arrive(o_free);
          if (tracer) ATTN_TRACE(2, i, 6);
        }
        if (q < p.T) {
          auto mix = [&](int k) {
            return fmaf(__uint_as_float(o0[k]), ka, __uint_as_float(o1[k]) * kb2);
          };
#pragma unroll
          for (int j = 0; j < 4; ++j)
            dst[hh * 4 + j] =
                make_uint4(pack_bf16(mix(8 * j + 0), mix(8 * j + 1)), pack_bf16(mix(8 * j + 2), mix(8 * j + 3)),
                           pack_bf16(mix(8 * j + 4), mix(8 * j + 5)), pack_bf16(mix(8 * j + 6), mix(8 * j + 7)));
        }
      }
      if (tracer) ATTN_TRACE(2, i, 7);
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

#ifdef OSUDIT_ATTN_TRACE
extern "C" int osudit_debug_attn_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, attn_tc::g_trace, sizeof(attn_tc::g_trace)) == cudaSuccess ? 0 : -1;
}
extern "C" int osudit_debug_attn_trace_chunks(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, attn_tc::g_trace_chunks, sizeof(attn_tc::g_trace_chunks)) == cudaSuccess ? 0 : -1;
}
#endif

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
