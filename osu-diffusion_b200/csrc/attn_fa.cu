// Streaming ("flash") self-attention on tcgen05 with two query tiles in flight per CTA.
//
// Reference: nn.MultiheadAttention's core inside DiTBlock (models.py:164-170): per head
// softmax(q k^T / sqrt(hd) + mask) v, under the band mask of sample.py:81-84 (query j sees key i iff
// -w_left <= i - j <= w_right) or no mask (training windows, train.py:249-255).  head_dim 64.
//
// One CTA works on TWO 128-query tiles at a time ("slots"); each slot streams the 128-key slabs its tile may attend
// (3 for the W = 128 band, T/128 for full attention, 1 for a 128-datapoint training window) through
//     S = Q K_j^T (tcgen05, 128 TMEM columns)  ->  softmax warps: P_j = exp2(S * c - ref), bf16, to shared memory
//     O_h += P_j[:, half h] V_j[half h] (tcgen05, 2 x 64 TMEM columns; V consumed as an MN-major B operand)
// Two threads share a query row (64 of the slab's 128 keys each: 8 softmax warps per slot, 4 per scheduler with both
// slots — a single warp per scheduler is latency-bound at ~900 cycles per 32-column chunk, tools/attn_fa_trace.py).
// Each half keeps its OWN running reference and its OWN output accumulator, so the halves never synchronise inside a
// tile; the epilogue combines them as (2^r0 O0 + 2^r1 O1) / (2^r0 s0 + 2^r1 s1).  The reference of a half is a lazily
// updated power of two: O_h and the running sum are rescaled only when the scores outgrow it by more than 2^24 (an
// exact power-of-two factor, applied to O_h in TMEM by the row's own thread), so the common path never touches O
// between slabs.  While one slot's softmax warps exponentiate, the other slot's MMAs and TMEM hand-offs run.
//
//   warp 0 / 1    TMA producer of slot 0 / 1: Q tile, then K_j / V_j through two-stage rings (3-D tensor map over the
//                 packed qkv [B, T, 3D]; rows outside [0, T) are zero-filled per batch)
//   warp 2 / 3    tcgen05.mma issuer of slot 0 / 1 (warp 2 also owns the TMEM allocation)
//   warps 4-11    softmax of slot 0: TMEM lane quadrant = warp % 4, key half = (warp - 4) / 4; half 1 runs the epilogue
//   warps 12-19   same for slot 1
// Optionally writes the log2-domain log-sum-exp of every row (the backward's input).
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

namespace attn_fa {

constexpr int kQ = 128;                      // queries per tile
constexpr int kS = 128;                      // keys per slab
constexpr int kHD = 64;
constexpr int kTile = kQ * kHD * 2;          // 16 KB: one [128][64] bf16 box
constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + kTile;         // 2 stages
constexpr int kOffV = kOffK + 2 * kTile;     // 2 stages
constexpr int kOffP = kOffV + 2 * kTile;     // P[128][128] as 2 K-blocks of 64 keys
constexpr int kSlot = kOffP + 2 * kTile;     // 112 KB per slot
constexpr int kSmemX = 2 * kSlot;              // [slot][row] (reference, sum) of half 0, handed to half 1's epilogue
constexpr int kSmemBar = kSmemX + 2 * kQ * 8;
constexpr int kBarsPerSlot = 13;
constexpr int kSmemBytes = kSmemBar + 2 * kBarsPerSlot * 8 + 8 + 16;  // 231.9 KB: no slack, the base is declared 1024-aligned
constexpr int kThreads = 20 * 32;
constexpr float kJump = 24.0f;               // see attn_window_tc.cu: P <= 2^24 before a re-reference

struct Params {
  CUtensorMap tma_qkv;  // 3-D: [3D cols, T, B], box [64, 128, 1]
  __nv_bfloat16* out;   // [B*T, D]
  float* lse;           // [B, H, T] or nullptr
  int B, T, H, D;
  int q_tiles, total_tiles;
  int w_left, w_right;  // allowed iff -w_left <= key - query <= w_right
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

// SWIZZLE_128B shared-memory descriptor, 8-row groups 1024 B apart (K-major operands with 128 B of K per row, and
// the MN-major V operand with 128 B of N per row).
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// kind::f16, D fp32, A/B bf16, M = 128; b_mn_major selects an MN-major B operand.
__device__ __forceinline__ constexpr uint32_t idesc(int n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}

struct Tile {
  int b, h, q0, slab_lo, n_slabs;
};

__device__ __forceinline__ Tile decode_tile(const Params& p, int tile) {
  Tile t;
  const int qt = tile % p.q_tiles;
  const int bh = tile / p.q_tiles;
  t.h = bh % p.H;
  t.b = bh / p.H;
  t.q0 = qt * kQ;
  const int kmin = max(t.q0 - p.w_left, 0);
  const int kmax = min(t.q0 + kQ - 1 + p.w_right, p.T - 1);
  t.slab_lo = kmin / kS;
  t.n_slabs = kmax / kS - t.slab_lo + 1;
  return t;
}

struct Bars {
  uint64_t* q_full;
  uint64_t* q_free;
  uint64_t* k_full;  // [2]
  uint64_t* k_free;  // [2]
  uint64_t* v_full;  // [2]
  uint64_t* v_free;  // [2]
  uint64_t* s_full;
  uint64_t* p_full;
  uint64_t* pv_done;
};

#if defined(OSUDIT_FA_ABL) && OSUDIT_FA_ABL == 3  // ablation: no TMEM reads of S
#define FA_LD(addr, regs) do { for (int k_ = 0; k_ < 32; ++k_) regs[k_] = __float_as_uint(0.01f * (k_ + threadIdx.x)); } while (0)
#else
#define FA_LD(addr, regs) tmem_ld_32x32(addr, regs)
#endif

__device__ __forceinline__ Bars slot_bars(uint64_t* base, int s) {
  uint64_t* b = base + s * kBarsPerSlot;
  return Bars{b + 0, b + 1, b + 2, b + 4, b + 6, b + 8, b + 10, b + 11, b + 12};
}

// Optional timeline instrumentation (build with -DOSUDIT_ATTN_TRACE): CTA 0 records clock64() at the hand-off points
// of slab steps 8..39 of each slot, for the MMA issuer (role 0) and softmax quadrant 0 (role 1);
// osudit_debug_fa_trace() copies the table out.  Never in the shipped build.
#ifdef OSUDIT_ATTN_TRACE
__device__ long long g_fa_trace[2 * 2 * 32 * 8];
#define FA_TRACE(slot, role, n, ev)                                                       \
  do {                                                                                    \
    if (blockIdx.x == 0 && (n) >= 8 && (n) < 40)                                          \
      g_fa_trace[((((slot) * 2 + (role)) * 32) + (n) - 8) * 8 + (ev)] = clock64();        \
  } while (0)
#else
#define FA_TRACE(slot, role, n, ev) do {} while (0)
#endif

__global__ void __launch_bounds__(kThreads, 1) attn_fa_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar_base = reinterpret_cast<uint64_t*>(smem + kSmemBar);
  uint64_t* stagger = bar_base + 2 * kBarsPerSlot;  // slot 0 finished its first slab: slot 1 may start
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_base + 2 * kBarsPerSlot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_qkv);
    for (int s = 0; s < 2; ++s) {
      const Bars b = slot_bars(bar_base, s);
      mbar_init(b.q_full, 1);
      mbar_init(b.q_free, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&b.k_full[i], 1);
        mbar_init(&b.k_free[i], 1);
        mbar_init(&b.v_full[i], 1);
        mbar_init(&b.v_free[i], 1);
      }
      mbar_init(b.s_full, 1);
      mbar_init(b.p_full, 8);  // one arrival per softmax warp
      mbar_init(b.pv_done, 1);
    }
    mbar_init(stagger, 8);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int G = static_cast<int>(gridDim.x);
  // slot s of this CTA works on tiles 2 * (blockIdx.x + i * G) + s, i = 0, 1, ...

  if (warp < 2) {
    // ------------------------------------------------------------------ TMA producer of slot `warp`
    if (lane == 0) {
      const int s = warp;
      const Bars bar = slot_bars(bar_base, s);
      uint8_t* base = smem + s * kSlot;
      uint32_t kn = 0, vn = 0, tn = 0;
      for (int tile = 2 * static_cast<int>(blockIdx.x) + s; tile < p.total_tiles; tile += 2 * G, ++tn) {
        const Tile t = decode_tile(p, tile);
        mbar_wait_sleep(bar.q_free, (tn & 1) ^ 1, 200);  // every S MMA of this slot's previous tile has read Q
        mbar_expect_tx(bar.q_full, kTile);
        tma_load_3d(base + kOffQ, &p.tma_qkv, bar.q_full, t.h * kHD, t.q0, t.b);
        for (int j = 0; j < t.n_slabs; ++j) {
          const int k0 = (t.slab_lo + j) * kS;
          const uint32_t ks = kn & 1, vs = vn & 1;
          mbar_wait_sleep(&bar.k_free[ks], ((kn >> 1) & 1) ^ 1, 200);
          mbar_expect_tx(&bar.k_full[ks], kTile);
          tma_load_3d(base + kOffK + ks * kTile, &p.tma_qkv, &bar.k_full[ks], p.D + t.h * kHD, k0, t.b);
          ++kn;
          mbar_wait_sleep(&bar.v_free[vs], ((vn >> 1) & 1) ^ 1, 200);
          mbar_expect_tx(&bar.v_full[vs], kTile);
          tma_load_3d(base + kOffV + vs * kTile, &p.tma_qkv, &bar.v_full[vs], 2 * p.D + t.h * kHD, k0, t.b);
          ++vn;
        }
      }
    }
  } else if (warp < 4) {
    // -------------------------------------------------------------------- MMA issuer of slot `warp - 2`
    // S(0) | wait P(0): S(1), PV(0) | wait P(1): S(2), PV(1) | ... ; S(j+1) goes in FRONT of PV(j) so the slot's
    // softmax warps get their next scores while the tensor pipe is still busy with this slab's PV.  One issuing
    // thread per slot (blocking waits only: a thread polling both slots would take issue cycles from the softmax
    // warps that share its scheduler); the two threads' MMAs target different accumulators and interleave freely.
    if (lane == 0) {
      constexpr uint32_t idesc_s = idesc(kS, false);
      constexpr uint32_t idesc_o = idesc(kHD, true);
      const int s = warp - 2;
      const Bars bar = slot_bars(bar_base, s);
      const uint32_t base = smem_u32(smem + s * kSlot);
      const uint64_t dq = desc_sw128(base + kOffQ);
      uint32_t kn = 0, vn = 0, sn = 0, tn = 0;
      auto issue_s = [&]() {  // S = Q K^T for the slot's next slab
        const uint32_t ks = kn & 1;
        mbar_wait(&bar.k_full[ks], (kn >> 1) & 1);
        tc_fence_after();
        const uint64_t dk = desc_sw128(base + kOffK + ks * kTile);
#pragma unroll
        for (int k = 0; k < kHD / 16; ++k)
          umma_bf16(tmem_base + s * kS, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        umma_commit(bar.s_full);  // the ONLY commit of this batch: a tcgen05.commit holds the issuing thread ~170 cycles,
        ++kn;                     // so K / Q are released by the softmax warps once they have seen s_full
      };
      auto issue_pv = [&](bool first) {  // O (+)= P V for the slab whose P was just published
        const uint32_t vs = vn & 1;
        mbar_wait(&bar.v_full[vs], (vn >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {  // key half kb of the slab accumulates into its own O (own softmax reference)
          const uint64_t dp = desc_sw128(base + kOffP + kb * kTile);
          const uint64_t dv = desc_sw128(base + kOffV + vs * kTile + kb * (64 * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 16 keys: +32 B along P's row, +16 rows (2048 B) in V
            umma_bf16(tmem_base + 2 * kS + (2 * s + kb) * kHD, dp + 2 * k, dv + 128 * k, idesc_o,
                      (first && k == 0) ? 0u : 1u);
        }
        umma_commit(bar.pv_done);  // V is released by the softmax warps once they have seen pv_done
        ++vn;
      };
      for (int tile = 2 * static_cast<int>(blockIdx.x) + s; tile < p.total_tiles; tile += 2 * G, ++tn) {
        const int n = decode_tile(p, tile).n_slabs;
        mbar_wait_sleep(bar.q_full, tn & 1, 64);
        issue_s();
        for (int j = 0; j < n; ++j, ++sn) {
          FA_TRACE(s, 0, sn, 0);
          mbar_wait_sleep(bar.p_full, sn & 1, 32);  // S(j) read, P(j) in shared memory
          FA_TRACE(s, 0, sn, 1);
          if (j + 1 < n) issue_s();
          FA_TRACE(s, 0, sn, 2);
          issue_pv(j == 0);
          FA_TRACE(s, 0, sn, 3);
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------- softmax + epilogue: two threads per query row (64 keys each)
    const int s = (warp - 4) >> 3;
    const int half = ((warp - 4) >> 2) & 1;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const bool lead = warp == 4 + 8 * s && lane == 0;  // releases the slot's K / V / Q buffers
    const Bars bar = slot_bars(bar_base, s);
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t col_s = s * kS + half * 64, col_o = 2 * kS + (2 * s + half) * kHD;
    const uint32_t p_row = smem_u32(smem + s * kSlot + kOffP) + half * kTile + row * 128;  // this thread's 64 keys of P
    const int swz = row & 7;
    float2* xchg = reinterpret_cast<float2*>(smem + kSmemX) + s * kQ;
    const uint32_t bar_a = 1 + 2 * s, bar_b = 2 + 2 * s;  // named barriers: half 0's (ref, sum) published / consumed
    uint32_t sn = 0;  // slab steps of this slot so far (parity of s_full / p_full / pv_done)
    uint32_t tiles_done = 0;
    // The two slots must not run in lockstep (both exponentiating, then both waiting for the tensor pipe): slot 1
    // starts its first slab when slot 0 has finished its own, after which the slots alternate.
    if (s == 1 && 2 * static_cast<int>(blockIdx.x) + 1 < p.total_tiles) warp_mbar_wait(stagger, 0);

    for (int tile = 2 * static_cast<int>(blockIdx.x) + s; tile < p.total_tiles; tile += 2 * G, ++tiles_done) {
      const Tile t = decode_tile(p, tile);
      const int q = t.q0 + row;
      float ref = -INFINITY;  // this half's running reference (integer-valued, log2 domain)
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;

      for (int j = 0; j < t.n_slabs; ++j) {
        const int k0 = (t.slab_lo + j) * kS + half * 64;  // first key of this thread's 64 columns
        // allowed columns of this row inside its 64, and the 32-column chunks any row of this warp needs
        const int c_lo = max(max(q - p.w_left, 0) - k0, 0);
        const int c_hi = min(min(q + p.w_right, p.T - 1) - k0, 63);
        const int w_lo = max(t.q0 + quad * 32 - p.w_left, 0) - k0;
        const int w_hi = min(t.q0 + quad * 32 + 31 + p.w_right, p.T - 1) - k0;
        const int ch_lo = (w_hi >= 0 && w_lo < 64) ? (max(w_lo, 0) >> 5) : 2;
        const int ch_hi = (w_hi >= 0 && w_lo < 64) ? (min(w_hi, 63) >> 5) : -1;
        float o_factor = 1.0f;  // what O_h (slabs 0..j-1) must be multiplied by before PV(j) accumulates onto it
        bool pv_pending = sn > 0;  // PV(sn-1) still reads P and V: awaited at the first store

        const bool tracer = quad == 0 && lane == 0 && half == 0;
        if (tracer) FA_TRACE(s, 1, sn, 0);
        warp_mbar_wait_sleep(bar.s_full, sn & 1, 20);
        if (tracer) FA_TRACE(s, 1, sn, 1);
        tc_fence_after();
        if (lead) {  // S(j) is complete: its K stage (and, after the last slab, Q) is free
          mbar_arrive(&bar.k_free[sn & 1]);
          if (j + 1 == t.n_slabs) mbar_arrive(bar.q_free);
        }
#pragma unroll 1  // one copy of the chunk body: the loop stays resident in the instruction cache
        for (int c = 0; c < 2; ++c) {
          uint32_t packed[16];
          const bool live = c >= ch_lo && c <= ch_hi;
          if (live) {
            uint32_t v[32];
            FA_LD(t_lane + col_s + c * 32, v);
            tmem_ld_wait();
            const int base = c * 32;
            if (!(base >= c_lo && base + 31 <= c_hi)) {  // boundary chunk: masked scores become -inf
              const int klo = c_lo - base, khi = c_hi - base;
#pragma unroll
              for (int k = 0; k < 32; ++k)
                if (k < klo || k > khi) v[k] = 0xff800000u;
            }
            if (ref == -INFINITY) {  // no allowed key met yet: the first chunk maximum is the reference
              float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
              for (int k = 0; k < 32; k += 2) {
                m0 = fmaxf(m0, __uint_as_float(v[k]));
                m1 = fmaxf(m1, __uint_as_float(v[k + 1]));
              }
              const float cm = fmaxf(m0, m1) * p.scale_log2;
              if (cm > -INFINITY) ref = ceilf(cm);
            }
            // Common path: exponentiate against the current reference straight away and track the largest exponent
            // on the side (off the MUFU stream's critical path); only if it exceeds kJump is the chunk redone.
            float c0, c1, c2, c3;
            auto exps = [&](float off) {
              float a0 = -INFINITY, a1 = -INFINITY;
              c0 = c1 = c2 = c3 = 0.f;
#pragma unroll
              for (int g = 0; g < 2; ++g) {  // two groups of 16 columns keep the live register set small
                float pv[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                  const float a = fmaf(__uint_as_float(v[16 * g + k]), p.scale_log2, -off);
#if defined(OSUDIT_FA_ABL) && OSUDIT_FA_ABL == 1  // ablation builds: no MUFU
                  pv[k] = a * 0.001f;
#else
                  pv[k] = fast_exp2(a);
#endif
                  if (k & 1) a1 = fmaxf(a1, a); else a0 = fmaxf(a0, a);
                }
#pragma unroll
                for (int k = 0; k < 16; k += 4) {
                  c0 += pv[k];
                  c1 += pv[k + 1];
                  c2 += pv[k + 2];
                  c3 += pv[k + 3];
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) packed[8 * g + k] = pack_bf16(pv[2 * k], pv[2 * k + 1]);
              }
              return fmaxf(a0, a1);
            };
            const float off = (ref == -INFINITY) ? 0.f : ref;
            const float amax = exps(off);
            if (amax > kJump) {  // rare: the scores outgrew the reference by more than 2^24: re-reference, redo
              const float new_ref = ceilf(amax + off);
              const float factor = fast_exp2(ref - new_ref);
              if (c == 1) {  // rescale chunk 0 of this slab, written but not yet published
                for (int jj = 0; jj < 4; ++jj) {
                  const uint32_t addr = p_row + ((jj ^ swz) << 4);
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
              o_factor *= factor;
              ref = new_ref;
              exps(ref);
            }
            sum0 += c0; sum1 += c1; sum2 += c2; sum3 += c3;
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) packed[k] = 0u;
          }
          if (pv_pending) {  // first store of the slab (warp-uniform)
            warp_mbar_wait(bar.pv_done, (sn - 1) & 1);
            if (lead) mbar_arrive(&bar.v_free[(sn - 1) & 1]);
            pv_pending = false;
          }
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)  // 32 keys = 64 B = four 16-byte pieces of the 128-byte row
            sts128(p_row + (((c * 4 + jj) ^ swz) << 4), packed[4 * jj], packed[4 * jj + 1], packed[4 * jj + 2], packed[4 * jj + 3]);
        }
        if (j > 0 && __any_sync(0xffffffffu, o_factor != 1.0f)) {  // rare: bring O_h down to the new reference
          tc_fence_after();
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t o[32];
            tmem_ld_32x32(t_lane + col_o + hh * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * o_factor);
            tmem_st_32x32(t_lane + col_o + hh * 32, o);
          }
          tmem_st_wait();
        }
        fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core's async proxy
        tc_fence_before();
        warp_mbar_arrive(bar.p_full);
        if (s == 0 && sn == 0) warp_mbar_arrive(stagger);
        if (tracer) FA_TRACE(s, 1, sn, 7);
        ++sn;
      }

      const float sum = (sum0 + sum1) + (sum2 + sum3);
      if (half == 0) {
        // hand (reference, sum) to the row's other thread and move on to the next tile
        if (tiles_done > 0) named_bar_sync(bar_b, 256);  // half 1 has consumed the previous tile's values
        xchg[row] = make_float2(ref, sum);
        __threadfence_block();
        asm volatile("bar.arrive %0, 256;" ::"r"(bar_a) : "memory");
        continue;
      }
      // ---- epilogue (half 1): (w0 O0 + w1 O1) / (w0 s0 + w1 s1) -> bf16 -> global; log-sum-exp for the backward
      named_bar_sync(bar_a, 256);
      const float2 other = xchg[row];
      asm volatile("bar.arrive %0, 256;" ::"r"(bar_b) : "memory");
      warp_mbar_wait(bar.pv_done, (sn - 1) & 1);
      tc_fence_after();
      const float rmax = fmaxf(other.x, ref);
      const float w_a = (other.x == -INFINITY) ? 0.f : fast_exp2(other.x - rmax);
      const float w_b = (ref == -INFINITY) ? 0.f : fast_exp2(ref - rmax);
      const float total = w_a * other.y + w_b * sum;
      const float inv = 1.0f / total;
      const float ka = w_a * inv, kb = w_b * inv;
      uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<int64_t>(t.b) * p.T + q) * p.D + t.h * kHD);
      const uint32_t col_o0 = 2 * kS + 2 * s * kHD;
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {  // two groups of 32 head dims
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(t_lane + col_o0 + hh * 32, o0);
        tmem_ld_32x32(t_lane + col_o0 + kHD + hh * 32, o1);
        tmem_ld_wait();
        if (q < p.T) {
          auto f = [&](int k) { return fmaf(__uint_as_float(o0[k]), ka, __uint_as_float(o1[k]) * kb); };
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            dst[hh * 4 + jj] =
                make_uint4(pack_bf16(f(8 * jj + 0), f(8 * jj + 1)), pack_bf16(f(8 * jj + 2), f(8 * jj + 3)),
                           pack_bf16(f(8 * jj + 4), f(8 * jj + 5)), pack_bf16(f(8 * jj + 6), f(8 * jj + 7)));
        }
      }
      if (p.lse != nullptr && q < p.T)
        p.lse[(static_cast<int64_t>(t.b) * p.H + t.h) * p.T + q] = rmax + log2f(total);
      tc_fence_before();  // the O reads above are ordered before this thread's next p_full arrival
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace attn_fa

#ifdef OSUDIT_ATTN_TRACE
extern "C" int osudit_debug_fa_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, attn_fa::g_fa_trace, sizeof(attn_fa::g_fa_trace)) == cudaSuccess ? 0 : -1;
}
#endif

bool attn_fa_applicable(int head_dim, const uint8_t* mask) { return head_dim == 64 && mask == nullptr; }

int attn_fa_launch(const void* qkv, void* out, int B, int T, int H, int w_left, int w_right, float* lse,
                   cudaStream_t stream) {
  using namespace attn_fa;
  Params p;
  const int D = H * kHD;
  int rc = make_tensor_map_3d(&p.tma_qkv, qkv, 3ull * D, T, B, 3ull * D * 2, 3ull * D * 2 * T, kHD, kQ);
  if (rc) return rc;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.B = B; p.T = T; p.H = H; p.D = D;
  p.q_tiles = (T + kQ - 1) / kQ;
  p.total_tiles = p.q_tiles * H * B;
  p.w_left = w_left;
  p.w_right = w_right;
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(kHD));
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return set_error(-5, cudaGetErrorString(e));
    configured = true;
  }
  const int pairs = (p.total_tiles + 1) / 2;
  const int grid = pairs < num_sms() ? pairs : num_sms();
  attn_fa_kernel<<<grid, kThreads, kSmemBytes, stream>>>(p);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

}  // namespace osudit
