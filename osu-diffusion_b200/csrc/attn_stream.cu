// Streaming self-attention on tcgen05, third schedule: ONE query tile per CTA at a time, every hand-off buffered two or
// three deep, softmax warps that do not wait for the tensor pipe in steady state, an instruction-lean exponentiation loop.
//
// Reference: nn.MultiheadAttention's core inside DiTBlock (models.py:164-170): per head
// softmax(q k^T / sqrt(hd) + mask) v, under the band mask of sample.py:81-84 (query j sees key i iff
// -w_left <= i - j <= w_right) or no mask (training windows, train.py:249-255).  head_dim 64.
//
// Why a third kernel (DESIGN.md section 4.2 [r2b], profiles/r02b_*): on the sampling band the window kernel's
// softmax warps idle 47 % of the time (S is single-buffered: they wait for S(i+1) and for PV(i)), the two-slot kernel
// (attn_fa.cu) hides those waits but issues 13.6 k warp instructions per tile, 9 per score, against the 8-cycle MUFU
// cadence that is the real floor (16 exp2 per clock per SM, measured; TMEM reads run at 225 B/clk and are not a bound).
// Here:
//   * the CTA's slabs (128 keys each) form ONE sequence n = 0, 1, 2, ... across its tiles; S(n) lands in TMEM buffer
//     n % 3 and is issued three slabs ahead of PV (together with PV(n-3), whose P it overwrites: the tensor pipe runs in
//     issue order), i.e. two slabs ahead of the softmax; Q / K / V travel through 3 / 4 / 5 stage rings;
//   * P never touches shared memory: each softmax thread writes its bf16 probabilities back into the TMEM columns
//     its scores came from (tcgen05.st) and PV takes its A operand from tensor memory.  With P in shared memory the
//     kernel moved 150 KB per slab through the 128 B/clk shared-memory port (Q + K and P + V operand reads, the P
//     stores, the TMA writes): 1 170 cycles per slab before any conflict, and the tensor pipe, starved, back-pressured
//     the issuing thread (750 cycles to issue one PV batch, tools/attn_stream_trace.py) — the bound all three earlier
//     attention kernels shared.  Now 85 KB per slab;
//   * two threads share a query row (64 keys of the slab each), each with its own lazily updated power-of-two
//     reference and its own output accumulator, so they never synchronise (as attn_fa.cu); four extra warps run the
//     epilogue ((2^r0 O0 + 2^r1 O1) / (2^r0 s0 + 2^r1 s1), bf16, log-sum-exp) so the softmax warps go straight on;
//   * per score: half an FFMA2 (scale and reference), one MUFU.EX2, half an FADD2 (row sum), half an F2FP — the row
//     maximum is NOT tracked per element: a chunk whose probabilities sum to more than 2^24 (or overflow) is redone
//     against a fresh reference, which is the only case the maximum matters for;
//   * the band mask costs two instructions per score on the boundary chunks only (a per-thread bit mask).
//
//   warp 0        TMA producer: Q tiles and K slabs        warp 1   TMA producer: V slabs
//   warp 2        tcgen05.mma issuer (owns the TMEM allocation)
//   warps 4-11    softmax: TMEM lane quadrant = warp % 4, key half = (warp - 4) / 4
//   warps 12-15   epilogue: one thread per query row
// (15 warps at 136 registers do not fit: registers are allocated to warps in groups of four.)
// TMEM: S/P buffers 0-127, 128-255, 256-383 (slab n uses n % 3), O[half] 384 + 64 * half.
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

namespace attn_st {

// Ablation builds (tools/build_variants.sh; wrong results, timing only): 1 = no MUFU, 2 = half the exponentials,
// 3 = no softmax arithmetic at all (the floor of the pipeline around it).
#ifndef ST_ABL
#define ST_ABL 0
#endif

constexpr int kQ = 128;                      // queries per tile
constexpr int kS = 128;                      // keys per slab
constexpr int kHD = 64;
constexpr int kTile = kQ * kHD * 2;          // 16 KB: one [128][64] bf16 box
constexpr int kQStages = 3, kKStages = 4, kVStages = 5;
constexpr int kRing = kQStages + kKStages + kVStages;
constexpr int kOffOut = kRing * kTile;       // output tile [128][64] bf16, 128-byte swizzled, for the TMA store
constexpr int kOffX = kOffOut + kTile;       // (reference, sum)[tile parity][half][row]
constexpr int kOffBar = kOffX + 2 * 2 * kQ * 8;
constexpr int kSBufs = 3;                     // S/P buffers in TMEM: S runs kSBufs - 1 slabs ahead of the softmax
constexpr int kNumBars = 2 * kRing + 2 * kSBufs + 2 + 2 + 2;
constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 16;
constexpr int kThreads = 16 * 32;
constexpr int kSoftWarp0 = 4, kEpiWarp0 = 12;

struct Params {
  CUtensorMap tma_qkv;  // 3-D: [3D cols, T, B], box [64, 128, 1]
  CUtensorMap tma_out;  // 3-D: [D cols, T, B], box [64, 128, 1]: out bf16 [B*T, D]
  float* lse;           // [B, H, T] or nullptr
  int B, T, H, D;
  int q_tiles, total_tiles;
  int step_q, step_r;   // gridDim.x / q_tiles, gridDim.x % q_tiles
  int w_left, w_right;  // allowed iff -w_left <= key - query <= w_right
  float scale_log2;
};

__device__ __forceinline__ void tma_store_3d(const void* tmap, const void* smem_src, int32_t c0, int32_t c1,
                                             int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar,
                                            int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// SWIZZLE_128B shared-memory descriptor, 8-row groups 1024 B apart (see attn_fa.cu).
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

// The tiles of a CTA: blockIdx.x, + gridDim.x, ...  One division at the start, then (query tile, batch * head) move
// by the precomputed quotient / remainder of gridDim.x by q_tiles: the per-tile decode sat on the MMA warp's path.
struct Walk {
  int tile, qt, bh;
  __device__ __forceinline__ void init(const Params& p, int t0) {
    tile = t0;
    qt = t0 % p.q_tiles;
    bh = t0 / p.q_tiles;
  }
  __device__ __forceinline__ void next(const Params& p, int G) {
    tile += G;
    qt += p.step_r;
    bh += p.step_q;
    if (qt >= p.q_tiles) {
      qt -= p.q_tiles;
      ++bh;
    }
  }
  __device__ __forceinline__ Tile get(const Params& p, bool need_bh) const {
    Tile t;
    t.h = need_bh ? bh % p.H : 0;
    t.b = need_bh ? bh / p.H : 0;
    t.q0 = qt * kQ;
    const int kmin = max(t.q0 - p.w_left, 0);
    const int kmax = min(t.q0 + kQ - 1 + p.w_right, p.T - 1);
    t.slab_lo = kmin / kS;
    t.n_slabs = kmax / kS - t.slab_lo + 1;
    return t;
  }
};

// A position in a ring of `stages` buffers; `phase` is the parity of the number of completed laps.
struct Ring {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int stages) {
    if (++stage == stages) {
      stage = 0;
      phase ^= 1;
    }
  }
};

// d = a * b + c on two lanes at once (FFMA2: one issue slot for two scores); b and c are broadcast.
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmov.b64 rc, {%5, %5};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1) {  // d += a
  asm("{\n\t.reg .b64 ra, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rd, {%0, %1};\n\t"
      "add.rn.f32x2 rd, rd, ra;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1));
}

#ifdef OSUDIT_ATTN_TRACE
__device__ long long g_st_trace[3 * 48 * 6];
#define ST_TRACE(role, n, ev)                                                                       \
  do {                                                                                              \
    if (blockIdx.x == 0 && (n) >= 12 && (n) < 60) g_st_trace[(((role) * 48) + (n) - 12) * 6 + (ev)] = clock64(); \
  } while (0)
#else
#define ST_TRACE(role, n, ev) do {} while (0)
#endif

__global__ void __launch_bounds__(kThreads, 1) attn_stream_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars;                         // [kQStages]
  uint64_t* q_free = q_full + kQStages;
  uint64_t* k_full = q_free + kQStages;            // [kKStages]
  uint64_t* k_free = k_full + kKStages;
  uint64_t* v_full = k_free + kKStages;            // [kVStages]
  uint64_t* v_free = v_full + kVStages;
  uint64_t* s_full = v_free + kVStages;            // [kSBufs] S(n) complete in TMEM buffer n % kSBufs
  uint64_t* p_full = s_full + kSBufs;              // [kSBufs] P(n) written back into that buffer (8 warp arrivals)
  uint64_t* o_full = p_full + kSBufs;              // every PV of a tile complete: O final
  uint64_t* o_free = o_full + 1;                   // the epilogue has read O (4 warp arrivals)
  uint64_t* x_full = o_free + 1;                   // [2] (reference, sum) of a tile published (8 warp arrivals)
  uint64_t* x_free = x_full + 2;                   // [2] ... and consumed by the epilogue (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int G = static_cast<int>(gridDim.x);
  uint8_t* ring_q = smem;
  uint8_t* ring_k = smem + kQStages * kTile;
  uint8_t* ring_v = smem + (kQStages + kKStages) * kTile;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_qkv);
    tma_prefetch_desc(&p.tma_out);
    for (int i = 0; i < 2 * kRing + kSBufs; ++i) mbar_init(&bars[i], 1);  // rings, s_full
    for (int i = 0; i < kSBufs; ++i) mbar_init(&p_full[i], 8);
    mbar_init(o_full, 1);
    mbar_init(o_free, 4);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 8);
      mbar_init(&x_free[i], 4);
    }
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

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: Q tiles and K slabs
    if (lane == 0) {
      Ring rq, rk;
      Walk w;
      for (w.init(p, blockIdx.x); w.tile < p.total_tiles; w.next(p, G)) {
        const Tile t = w.get(p, true);
        mbar_wait_sleep(&q_free[rq.stage], rq.phase ^ 1, 200);
        mbar_expect_tx(&q_full[rq.stage], kTile);
        tma_load_3d(ring_q + rq.stage * kTile, &p.tma_qkv, &q_full[rq.stage], t.h * kHD, t.q0, t.b);
        rq.advance(kQStages);
        for (int j = 0; j < t.n_slabs; ++j) {
          mbar_wait_sleep(&k_free[rk.stage], rk.phase ^ 1, 200);
          mbar_expect_tx(&k_full[rk.stage], kTile);
          tma_load_3d(ring_k + rk.stage * kTile, &p.tma_qkv, &k_full[rk.stage], p.D + t.h * kHD,
                      (t.slab_lo + j) * kS, t.b);
          rk.advance(kKStages);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ TMA producer: V slabs
    if (lane == 0) {
      Ring rv;
      Walk w;
      for (w.init(p, blockIdx.x); w.tile < p.total_tiles; w.next(p, G)) {
        const Tile t = w.get(p, true);
        for (int j = 0; j < t.n_slabs; ++j) {
          mbar_wait_sleep(&v_free[rv.stage], rv.phase ^ 1, 200);
          mbar_expect_tx(&v_full[rv.stage], kTile);
          tma_load_3d(ring_v + rv.stage * kTile, &p.tma_qkv, &v_full[rv.stage], 2 * p.D + t.h * kHD,
                      (t.slab_lo + j) * kS, t.b);
          rv.advance(kVStages);
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ MMA issuer
    // S(0) S(1) S(2) | wait P(0): PV(0) S(3) | wait P(1): PV(1) S(4) | ...   S runs two slabs ahead of the softmax
    // (three buffers), so the scores of a slab are complete long before the softmax warps ask for them; PV(n) reads
    // P(n) from the TMEM buffer S(n) came in, and S(n+3), issued behind it, overwrites that buffer (the tensor pipe runs
    // in issue order).  With two buffers S(n+1) could only be issued after P(n-1) was published, i.e. when the softmax
    // warps were already asking for it: they waited 44 % of the time (ncu, profiles/r02_summary.md).
    // The WHOLE warp runs this loop converged and one elected lane issues: every operand of tcgen05.mma must sit in
    // a uniform register, and inside an `if (lane == 0)` region the compiler cannot prove that, so it wraps each MMA
    // in an ELECT / R2UR / BRA.U.ANY loop.
    {
      constexpr uint32_t idesc_s = idesc(kS, false);
      constexpr uint32_t idesc_o = idesc(kHD, true);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);  // provably warp-uniform
      const uint32_t sbase = smem_u32(smem);
      // descriptors of stage 0 of each ring; stage s adds s * (kTile >> 4) to the 14-bit address field (the whole
      // dynamic shared memory lies below 256 KB, so the field cannot carry)
      const uint64_t dq0 = desc_sw128(sbase);
      const uint64_t dk0 = desc_sw128(sbase + kQStages * kTile);
      const uint64_t dv0 = desc_sw128(sbase + (kQStages + kKStages) * kTile);
      constexpr uint32_t kStageStep = kTile >> 4;
      // A barrier probe is a round trip through the memory-instruction queue (~200 cycles even on success, more
      // while the softmax warps of this scheduler keep that queue full of MUFU work), and so is a tcgen05.commit; the
      // warp's service time per slab, not the tensor pipe, bounded the kernel (tools/attn_stream_trace.py).  So: the
      // barriers that are normally complete (V, K, Q, O free) are probed BEFORE the wait for P, their answers are
      // used after it; and there is ONE commit per slab: PV(n) and S(n+3) are issued together, S(n+3)'s completion
      // (the tensor pipe runs in issue order) doubles as "PV(n) done" for the V ring and the softmax warps.
      auto probe = [&](uint64_t* bar, uint32_t parity) { return lane != 0 || mbar_try_wait(bar, parity); };
      auto ensure = [&](bool ok, uint64_t* bar, uint32_t parity) {
        if (!ok) mbar_wait_sleep(bar, parity, 20);  // lane 0 only
      };
      // S cursor: the slab whose scores are issued next (three ahead of the PV cursor)
      Walk ws;
      ws.init(p, blockIdx.x);
      int j_s = 0, ns_s = ws.get(p, false).n_slabs;
      Ring rq, rk, rs_s;
      auto s_mmas = [&]() {  // elected lane
        const uint64_t dq = dq0 + static_cast<uint32_t>(rq.stage) * kStageStep;
        const uint64_t dk = dk0 + static_cast<uint32_t>(rk.stage) * kStageStep;
#pragma unroll
        for (int k = 0; k < kHD / 16; ++k)
          umma_bf16(tb + rs_s.stage * kS, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
      };
      auto s_advance = [&]() {
        rk.advance(kKStages);
        if (++j_s == ns_s) {
          j_s = 0;
          ws.next(p, G);
          rq.advance(kQStages);
          ns_s = ws.get(p, false).n_slabs;
        }
      };
      for (int i = 0; i < kSBufs; ++i) {
        if (ws.tile < p.total_tiles) {
          if (lane == 0) {
            if (j_s == 0) mbar_wait_sleep(&q_full[rq.stage], rq.phase, 20);
            mbar_wait_sleep(&k_full[rk.stage], rk.phase, 20);
          }
          __syncwarp();
          tc_fence_after();
          if (elect_one()) {
            s_mmas();
            umma_commit(&s_full[rs_s.stage]);
          }
          __syncwarp();
          s_advance();
          rs_s.advance(kSBufs);
        }
      }
      uint32_t n = 0, tn = 0;
      Ring rv, rs;
      Walk w;
      for (w.init(p, blockIdx.x); w.tile < p.total_tiles; w.next(p, G), ++tn) {
        const int ns = w.get(p, false).n_slabs;
        constexpr uint32_t col_o = kSBufs * kS;
        for (int j = 0; j < ns; ++j, ++n) {
          const bool s_valid = ws.tile < p.total_tiles;
          const bool ok_v = probe(&v_full[rv.stage], rv.phase);
          const bool ok_k = !s_valid || probe(&k_full[rk.stage], rk.phase);
          const bool ok_q = !(s_valid && j_s == 0) || probe(&q_full[rq.stage], rq.phase);
          const bool ok_o = !(j == 0 && tn >= 1) || probe(o_free, (tn - 1) & 1);  // the epilogue has read the previous O
          if (lane == 0) {
            ST_TRACE(0, n, 0);
            mbar_wait_sleep(&p_full[rs.stage], rs.phase, 20);
            ST_TRACE(0, n, 1);
            ensure(ok_v, &v_full[rv.stage], rv.phase);
            ensure(ok_o, o_free, (tn - 1) & 1);
            if (s_valid) {
              ensure(ok_k, &k_full[rk.stage], rk.phase);
              ensure(ok_q, &q_full[rq.stage], rq.phase);
            }
            ST_TRACE(0, n, 3);
          }
          __syncwarp();
          tc_fence_after();
          const uint64_t dv = dv0 + static_cast<uint32_t>(rv.stage) * kStageStep;
          const uint32_t pa = tb + rs.stage * kS;  // P(n): per key half 32 columns of packed bf16 pairs
          if (elect_one()) {
            ST_TRACE(0, n, 4);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {  // key half kb accumulates into its own O (own softmax reference)
#pragma unroll
              for (int k = 0; k < 4; ++k)  // 16 keys: 8 TMEM columns of P, 16 rows (2048 B) of V
                umma_bf16_ts(tb + col_o + kb * kHD, pa + kb * 64 + 8 * k, dv + kb * (64 * 128 >> 4) + 128 * k, idesc_o,
                             (j == 0 && k == 0) ? 0u : 1u);
            }
            if (j + 1 == ns) umma_commit(o_full);
            if (s_valid) s_mmas();
            ST_TRACE(0, n, 5);
            umma_commit(&s_full[rs_s.stage]);  // S(n+3) complete; without further slabs it still marks "PV(n) done"
          }
          __syncwarp();
          if (s_valid) s_advance();
          rs_s.advance(kSBufs);
          rv.advance(kVStages);
          rs.advance(kSBufs);
          if (lane == 0) ST_TRACE(0, n, 2);
        }
      }
    }
  } else if (warp >= kSoftWarp0 && warp < kEpiWarp0) {
    // ------------------------------------------------- softmax: two threads per query row (64 keys of the slab each)
    const int half = (warp - kSoftWarp0) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const bool lead = warp == kSoftWarp0 && lane == 0;  // releases K / Q / V buffers
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    float2* xchg = reinterpret_cast<float2*>(smem + kOffX);
    uint32_t n = 0, tn = 0;
    Ring rk, rq, rv_rel, rs;  // K / Q stages to release; V stage of slab n - 2; S buffer of slab n
    const float scale = p.scale_log2;

    // The scores of slab n+1 are fetched while slab n is still being exponentiated (its buffer is the other one and
    // S runs two slabs ahead), so the barrier probe and the TMEM read latency sit under the MUFU stream.
    uint32_t va[32], vb[32];
    warp_mbar_wait_sleep(&s_full[0], 0, 20);
    tc_fence_after();
    constexpr uint32_t col_o_base = kSBufs * kS;
    tmem_ld_32x32(t_lane + half * 64, va);
    tmem_ld_32x32(t_lane + half * 64 + 32, vb);

    Walk w;
    for (w.init(p, blockIdx.x); w.tile < p.total_tiles; w.next(p, G), ++tn) {
      const Tile t = w.get(p, false);
      const int tile = w.tile;
      const int q = t.q0 + row;
      const uint32_t col_o = col_o_base + half * kHD;
      float ref = -INFINITY;  // this half's running reference (integer-valued, log2 domain)
      float sum = 0.f;

      for (int j = 0; j < t.n_slabs; ++j, ++n) {
        const int k0 = (t.slab_lo + j) * kS + half * 64;  // first key of this thread's 64 columns
        // allowed columns of this row inside its 64, and the 32-column chunks any row of this warp needs
        const int c_lo = max(max(q - p.w_left, 0) - k0, 0);
        const int c_hi = min(min(q + p.w_right, p.T - 1) - k0, 63);
        const int w_lo = max(t.q0 + quad * 32 - p.w_left, 0) - k0;
        const int w_hi = min(t.q0 + quad * 32 + 31 + p.w_right, p.T - 1) - k0;
        const bool any_live = w_hi >= 0 && w_lo < 64;
        const bool live0 = any_live && w_lo < 32;
        const bool live1 = any_live && w_hi >= 32;
        const bool has_next = j + 1 < t.n_slabs || tile + G < p.total_tiles;
        // this thread's 64 score columns; its probabilities go back into the first 32 of them, two keys per column
        Ring rs_next = rs;
        rs_next.advance(kSBufs);
        const uint32_t t_s = t_lane + rs.stage * kS + half * 64;
        const uint32_t t_next = t_lane + rs_next.stage * kS + half * 64;
        float o_factor = 1.0f;  // what O_h (slabs 0..j-1 of the tile) must be multiplied by before PV(n) accumulates

        const bool tracer = quad == 0 && lane == 0;
        if (tracer) ST_TRACE(1 + half, n, 0);
        // probe S(n+1) now, use the answer after the first chunk: a barrier probe costs ~240 cycles even on success
        bool next_ready = true;
        if (has_next && lane == 0) next_ready = mbar_try_wait(&s_full[rs_next.stage], rs_next.phase);
        tmem_ld_wait();  // all 64 scores of slab n are in registers: their columns may be overwritten from here on
        if (tracer) ST_TRACE(1 + half, n, 2);
        if (lead) {  // S(n) is complete: its K stage (and, after the tile's last slab, Q) is free
          mbar_arrive(&k_free[rk.stage]);
          if (j + 1 == t.n_slabs) mbar_arrive(&q_free[rq.stage]);
        }
        rk.advance(kKStages);
        if (j + 1 == t.n_slabs) rq.advance(kQStages);

        auto chunk = [&](uint32_t(&v)[32], int c, bool live) {
          uint32_t packed[16];
          float csum = 0.f;
          auto row_max = [&]() {
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
              m0 = fmaxf(m0, fmaxf(__uint_as_float(v[k]), __uint_as_float(v[k + 1])));
              m1 = fmaxf(m1, fmaxf(__uint_as_float(v[k + 2]), __uint_as_float(v[k + 3])));
            }
            return fmaxf(m0, m1) * scale;
          };
          auto exps = [&](float noff) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
              float e0, e1, e2, e3;
              ffma2(e0, e1, __uint_as_float(v[k]), __uint_as_float(v[k + 1]), scale, noff);
              ffma2(e2, e3, __uint_as_float(v[k + 2]), __uint_as_float(v[k + 3]), scale, noff);
#if ST_ABL == 1  // ablation: no MUFU at all (wrong results; timing only)
              e0 *= 0.001f; e1 *= 0.001f; e2 *= 0.001f; e3 *= 0.001f;
#elif ST_ABL == 2  // ablation: half the exponentials
              e0 = fast_exp2(e0);
              e1 *= 0.001f;
              e2 = fast_exp2(e2);
              e3 *= 0.001f;
#else
              e0 = fast_exp2(e0);
              e1 = fast_exp2(e1);
              e2 = fast_exp2(e2);
              e3 = fast_exp2(e3);
#endif
              fadd2(a0, a1, e0, e1);
              fadd2(a2, a3, e2, e3);
              packed[k >> 1] = pack_bf16(e0, e1);
              packed[(k >> 1) + 1] = pack_bf16(e2, e3);
            }
            csum = (a0 + a1) + (a2 + a3);
          };
#if ST_ABL == 3  // ablation: no softmax arithmetic at all (wrong results; the floor of the pipeline around it)
          if (live) {
#pragma unroll
            for (int k = 0; k < 16; ++k) packed[k] = v[2 * k] & 0x3f803f80u;
            sum += 1.0f;
          } else
#endif
          if (live) {  // warp-uniform
            const int base = c * 32;
            if (!(base >= c_lo && base + 31 <= c_hi)) {  // boundary chunk: masked scores become -inf
              const int klo = c_lo - base, khi = c_hi - base;  // allowed: klo <= k <= khi
              uint32_t m = 0u;
              if (khi >= 0 && klo <= 31) {
                const uint32_t hi_mask = khi >= 31 ? 0xffffffffu : ((2u << khi) - 1u);
                const uint32_t lo_mask = klo <= 0 ? 0xffffffffu : ~((1u << klo) - 1u);
                m = hi_mask & lo_mask;
              }
#pragma unroll
              for (int k = 0; k < 32; ++k)
                if (!(m & (1u << k))) v[k] = 0xff800000u;
            }
            if (ref == -INFINITY) {  // no allowed key met yet: the first chunk maximum is the reference
              const float cm = row_max();
              if (cm > -INFINITY) ref = ceilf(cm);
            }
            exps((ref == -INFINITY) ? 0.f : -ref);
            const bool redo = !(csum <= 16777216.0f);  // the scores outgrew the reference (or overflowed)
            if (__any_sync(0xffffffffu, redo)) {  // rare; tcgen05.ld / st are warp-collective, so the whole warp comes
              float factor = 1.0f;
              if (redo) {
                const float new_ref = ceilf(row_max());
                factor = fast_exp2(ref - new_ref);
                ref = new_ref;
                sum *= factor;
                o_factor *= factor;
              }
              if (c == 1) {  // rescale chunk 0 of this slab, written but not yet published
                uint32_t w[16];
                tmem_st_wait();
                tmem_ld_32x16(t_s, w);
                tmem_ld_wait();  // (also completes the prefetch of the next slab's chunk 0: harmless)
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                  const float2 f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w[k]));
                  w[k] = pack_bf16(f.x * factor, f.y * factor);
                }
                tmem_st_32x16(t_s, w);
              }
              if (redo) exps(-ref);
            }
            sum += csum;
          } else {  // no row of this warp may attend these 32 keys: P = 0 (PV covers the whole slab)
            uint32_t zero[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) zero[k] = 0u;
            tmem_st_32x16(t_s + c * 16, zero);
            return;
          }
          tmem_st_32x16(t_s + c * 16, packed);
        };
        chunk(va, 0, live0);
        if (tracer) ST_TRACE(1 + half, n, 3);
        if (has_next) {  // slab n+1's scores, first chunk: va is free
          if (!next_ready) mbar_wait_sleep(&s_full[rs_next.stage], rs_next.phase, 20);  // lane 0 only
          if (lead && n >= 2) mbar_arrive(&v_free[rv_rel.stage]);  // S(n+1) complete => PV(n-2), issued before it, too
          __syncwarp();
          tc_fence_after();
          tmem_ld_32x32(t_next, va);
        }
        if (tracer) ST_TRACE(1 + half, n, 1);
        chunk(vb, 1, live1);
        if (tracer) ST_TRACE(1 + half, n, 4);
        if (has_next) tmem_ld_32x32(t_next + 32, vb);

        if (j > 0 && __any_sync(0xffffffffu, o_factor != 1.0f)) {  // rare: bring O_h down to the new reference
          Ring rs_nn = rs_next;
          rs_nn.advance(kSBufs);
          warp_mbar_wait(&s_full[rs_nn.stage], rs_nn.phase);  // S(n+2)'s commit follows PV(n-1): it has landed in O_h
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
        }
        tmem_st_wait();  // P(n) (and a rescaled O_h) are in tensor memory
        tc_fence_before();
        warp_mbar_arrive(&p_full[rs.stage]);
        rs = rs_next;
        if (n >= 2) rv_rel.advance(kVStages);
        if (tracer) ST_TRACE(1 + half, n, 5);
      }

      // hand (reference, sum) to the epilogue warps
      if (tn >= 2) warp_mbar_wait_sleep(&x_free[tn & 1], ((tn >> 1) & 1) ^ 1, 20);
      xchg[((tn & 1) * 2 + half) * kQ + row] = make_float2(ref, sum);
      warp_mbar_arrive(&x_full[tn & 1]);
    }
  } else if (warp >= kEpiWarp0) {
    // ---- epilogue: (w0 O0 + w1 O1) / (w0 s0 + w1 s1) -> bf16 -> global; log-sum-exp for the backward
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float2* xchg = reinterpret_cast<const float2*>(smem + kOffX);
    uint32_t tn = 0;
    Walk w;
    for (w.init(p, blockIdx.x); w.tile < p.total_tiles; w.next(p, G), ++tn) {
      const Tile t = w.get(p, true);
      const int q = t.q0 + row;
      const uint32_t par = (tn >> 1) & 1;
      // the previous tile's TMA store has read the staging tile (per-thread rows would touch 32 cache lines per
      // store instruction and hold the LSU for ~1 000 cycles per tile, which every mbarrier operation of the CTA
      // queues behind: tools/attn_stream_trace.py showed all roles stalling at tile boundaries)
      warp_mbar_wait_sleep(&x_full[tn & 1], par, 100);
      const float2 ha = xchg[((tn & 1) * 2 + 0) * kQ + row];
      const float2 hb = xchg[((tn & 1) * 2 + 1) * kQ + row];
      warp_mbar_arrive(&x_free[tn & 1]);
      const float rmax = fmaxf(ha.x, hb.x);
      const float w_a = (ha.x == -INFINITY) ? 0.f : fast_exp2(ha.x - rmax);
      const float w_b = (hb.x == -INFINITY) ? 0.f : fast_exp2(hb.x - rmax);
      const float total = w_a * ha.y + w_b * hb.y;
      const float inv = 1.0f / total;
      const float ka = w_a * inv, kb = w_b * inv;
      constexpr uint32_t col_o = kSBufs * kS;
      warp_mbar_wait_sleep(o_full, tn & 1, 40);
      tc_fence_after();
      if (warp == kEpiWarp0 && lane == 0) tma_store_wait_read<0>();
      named_bar_sync(1, 128);
      const uint32_t st_row = smem_u32(smem + kOffOut) + row * 128;
      const int swz = row & 7;
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {  // two groups of 32 head dims
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(t_lane + col_o + hh * 32, o0);
        tmem_ld_32x32(t_lane + col_o + kHD + hh * 32, o1);
        tmem_ld_wait();
        if (hh == 1) {  // O is in registers: the next tile may overwrite it
          tc_fence_before();
          warp_mbar_arrive(o_free);
        }
        auto f = [&](int k) { return fmaf(__uint_as_float(o0[k]), ka, __uint_as_float(o1[k]) * kb); };
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          sts128(st_row + (((hh * 4 + jj) ^ swz) << 4), pack_bf16(f(8 * jj + 0), f(8 * jj + 1)),
                 pack_bf16(f(8 * jj + 2), f(8 * jj + 3)), pack_bf16(f(8 * jj + 4), f(8 * jj + 5)),
                 pack_bf16(f(8 * jj + 6), f(8 * jj + 7)));
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (warp == kEpiWarp0 && lane == 0) {  // rows >= T are clipped by the tensor map
        tma_store_3d(&p.tma_out, smem + kOffOut, t.h * kHD, t.q0, t.b);
        tma_store_commit();
      }
      if (p.lse != nullptr && q < p.T)
        p.lse[(static_cast<int64_t>(t.b) * p.H + t.h) * p.T + q] = rmax + log2f(total);
    }
    if (warp == kEpiWarp0 && lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace attn_st

#ifdef OSUDIT_ATTN_TRACE
extern "C" int osudit_debug_stream_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, attn_st::g_st_trace, sizeof(attn_st::g_st_trace)) == cudaSuccess ? 0 : -1;
}
#endif

bool attn_stream_applicable(int head_dim, const uint8_t* mask) { return head_dim == 64 && mask == nullptr; }

int attn_stream_launch(const void* qkv, void* out, int B, int T, int H, int w_left, int w_right, float* lse,
                       cudaStream_t stream) {
  using namespace attn_st;
  Params p;
  const int D = H * kHD;
  int rc = make_tensor_map_3d(&p.tma_qkv, qkv, 3ull * D, T, B, 3ull * D * 2, 3ull * D * 2 * T, kHD, kQ);
  if (rc) return rc;
  rc = make_tensor_map_3d(&p.tma_out, out, D, T, B, 1ull * D * 2, 1ull * D * 2 * T, kHD, kQ);
  if (rc) return rc;
  p.lse = lse;
  p.B = B; p.T = T; p.H = H; p.D = D;
  p.q_tiles = (T + kQ - 1) / kQ;
  p.total_tiles = p.q_tiles * H * B;
  p.w_left = w_left;
  p.w_right = w_right;
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(kHD));
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return set_error(-5, cudaGetErrorString(e));
    configured = true;
  }
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  p.step_q = grid / p.q_tiles;
  p.step_r = grid % p.q_tiles;
  attn_stream_kernel<<<grid, kThreads, kSmemBytes, stream>>>(p);
  OSUDIT_CHECK_LAUNCH();
  return 0;
}

}  // namespace osudit
