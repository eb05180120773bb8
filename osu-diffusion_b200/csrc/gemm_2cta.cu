// 2-CTA (cta_group::2) variant of the bf16 GEMM for the four large per-block contractions
// (models.py:164-170 QKV / out-proj, models.py:112-119 fc1+GELU / fc2).
//
// A CTA pair (one TPC) computes a 256 x BN output tile with UMMA M=256, N=BN (256, or 192 where 256 does not divide
// the width): each CTA stages its own 128 rows of A and its own BN/2 rows (N half) of B, so per MMA every SM reads 8 KB of shared
// memory instead of 12 KB and TMA writes 32 KB instead of 48 KB per k-block — the 1-CTA kernel's
// K=768 shapes sit at ~70 % tensor-pipe activity because of exactly that traffic (profiles/r01).
//
//   both CTAs   warp 0: TMA producer (cp.async.bulk.tensor ... .cta_group::2: completion bytes land on
//               the LEADER's full barrier); warps 2-9: epilogue of the CTA's own 128 accumulator rows
//   leader CTA  warp 1 lane 0: tcgen05.mma.cta_group::2 issuer; tcgen05.commit multicasts to the empty /
//               tmem_full barriers of both CTAs
//   peer CTA    epilogue warps arrive remotely (mapa) on the leader's tmem_empty barrier
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"
#include "ptx.cuh"

namespace osudit {

int make_tensor_map_2d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t stride_bytes,
                       uint32_t b0, uint32_t b1, bool is_f32);

namespace g2 {

constexpr int BM = 128;        // rows per CTA (256 per pair)
// BN (template parameter): columns per pair, 256 or — for widths that 256 does not divide, DiT-XL's 1152 / 3456 and
// DiT-S's 384 / 1152 — 192; each CTA stages BN/2 rows of B.
constexpr int BK = 64;
#ifndef OSUDIT_GELU_FP32  // the forward GELU epilogue works on half2 pairs unless this is defined (see gelu_tanh_pair)
#define OSUDIT_GELU_H2 1
#endif
#ifndef OSUDIT_G2_EPI_GROUPS
#define OSUDIT_G2_EPI_GROUPS 1
#endif
// Epilogue warp groups: each group is 8 warps (2 per TMEM lane quadrant) that own every kEpiGroups-th 64-column
// chunk of a tile, with their own staging buffers, named barrier and TMA-store bookkeeping.  Measured on B200
// (tools/gemm_bench.py, M = 262144): two groups (16 epilogue warps, 5 instead of 6 operand stages) give fc1+GELU
// +2 % (1202 -> 1226 TFLOP/s) but QKV -4 %, out-proj -4 %, fc2 -2 %: the GELU epilogue is not latency-bound, so
// the default stays one group.
constexpr int kEpiGroups = OSUDIT_G2_EPI_GROUPS;
constexpr int kStages = kEpiGroups == 1 ? 6 : 5;
constexpr int kBytesA = BM * BK * 2;
constexpr int kStagingBytes = BM * 128;
constexpr int kThreads = 64 + 256 * kEpiGroups;
template <int BN>
struct Cfg {
  static_assert(BN == 256 || BN == 192, "UMMA N of the pair");
  static constexpr int kBytesB = (BN / 2) * BK * 2;
  static constexpr int kStageBytes = kBytesA + kBytesB;  // 32 KB (28 KB) per CTA, a multiple of the 1 KB swizzle atom
  static constexpr int kBiasBytes = kEpiGroups * 2 * BN * 4;  // bias slice of the current / next tile, per epilogue group
  static constexpr int kSmemBytes = kStages * kStageBytes + 2 * kEpiGroups * kStagingBytes + 1024 + kBiasBytes;
};
// EPI_BF16_DGELU gives one operand stage to the two TMA-staged chunks of its auxiliary operand; EPI_RESID (5) to a
// second pair of fp32 staging boxes (its chunks are 32 KB, double-buffered like the 16 KB bf16 chunks of the others)
__host__ __device__ constexpr int stages_for(int epi) { return (epi == 4 || epi == 5) ? kStages - 1 : kStages; }
template <int BN>
constexpr int smem_bytes_for(int epi) {
  return stages_for(epi) * Cfg<BN>::kStageBytes + ((epi == 4 || epi == 5) ? 2 * kStagingBytes : 0) +
         2 * kEpiGroups * kStagingBytes + 1024 + Cfg<BN>::kBiasBytes;
}

struct Params {
  CUtensorMap tma_a, tma_b, tma_out, tma_aux;
  int kblocks, M, N, m_tiles, n_tiles;
  const float* bias;
  const __nv_bfloat16* aux;  // EPI_BF16_DGELU: the saved GELU derivative, read straight from global memory
  int64_t ld_aux;
  const float* gate;         // EPI_RESID: gate[b, n] at gate + b * gate_ld + n, b = row / rows_per_batch
  int64_t gate_ld;
  int rows_per_batch, batches;
};

// EPI_BF16_GELU_SAVE (training forward of fc1): out = gelu(acc + bias) and aux = gelu'(acc + bias) — the only
// thing the backward needs from the pre-activation, computed here from the same tanh — two stores per chunk
// instead of a separate GELU pass over HBM.
// EPI_BF16_DGELU (training backward through the GELU): out = acc * aux, the data gradient of fc2 multiplied by the
// saved derivative in the epilogue (one multiply per element; ncu showed the tensor pipe at 43 % when the
// derivative was recomputed here), so du is never materialised.
// EPI_RESID (inference): the gated residual update of the block, x += gate * (acc + bias) (models.py:164-174), done
// here as an fp32 TMA reduce-add into the residual stream instead of writing a bf16 branch that the next
// LayerNorm kernel reads back and folds in: that kernel is HBM-bound (12 D bytes per token: x in, branch in, x out,
// h out) and drops to 6 D (x in, h out); the 6 D move into this GEMM, which has HBM headroom.  The branch is no
// longer rounded to bf16 on the way.
enum : int { EPI_BF16 = 1, EPI_BF16_GELU = 2, EPI_BF16_GELU_SAVE = 3, EPI_BF16_DGELU = 4, EPI_RESID = 5 };

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t smem_addr) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(smem_addr));
  return r;
}
// Default semantics (release at CTA scope), as the TMEM hand-off needs no memory ordering beyond the tcgen05 fences:
// the explicit .release.cluster form cost ~2 400 cycles per tile on the epilogue's critical path (tools/gemm_trace.py).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* tmap, uint32_t leader_bar,
                                                 int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {  // arrives on `bar` in BOTH CTAs
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
  return static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF) | (1ull << 16) |
         (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ float gelu_tanh(float x) {
  // 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))) with the hardware tanh (one MUFU op per element;
  // its 2^-11 error is far below the bf16 rounding of the result that follows)
  const float u = x * fmaf(x * x, 0.044715f * 0.7978845608028654f, 0.7978845608028654f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// Two GELUs at once in half2 arithmetic (5 HFMA2-pipe instructions + one tanh.approx.f16x2 for the pair instead of
// 2 x (5 + 1) in fp32): the result is rounded to bf16 (8 mantissa bits) right after, fp16 carries 11.  With the
// epilogue off the critical path the GELU still slowed the MMA issue loop (7.7k vs 6.9k cycles per tile, contention
// for the SM's issue / pipes); halving its instruction count brings fc1 from 1266 to 1314 TFLOP/s.
__device__ __forceinline__ uint32_t gelu_tanh_pair(float a, float b) {
  const __half2 x = __floats2half2_rn(a, b);
  const __half2 x2 = __hmul2(x, x);  // overflows to inf for |x| > 255: tanh(+-inf) = +-1 gives x or 0, as it should
  const __half2 pz = __hfma2(x2, __float2half2_rn(0.044715f * 0.7978845608028654f), __float2half2_rn(0.7978845608028654f));
  const __half2 u = __hmul2(x, pz);
  uint32_t t;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(*reinterpret_cast<const uint32_t*>(&u)));
  const __half2 hx = __hmul2(x, __float2half2_rn(0.5f));
  const float2 g = __half22float2(__hfma2(hx, *reinterpret_cast<const __half2*>(&t), hx));
  return pack_bf16(g.x, g.y);
}

// x -> gelu_tanh(x) (returned in place) and its derivative, from one tanh
__device__ __forceinline__ float gelu_tanh_with_grad(float& x) {
  const float x2 = x * x;
  const float u = x * fmaf(x2, 0.044715f * 0.7978845608028654f, 0.7978845608028654f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float du = fmaf(x2, 3.0f * 0.044715f * 0.7978845608028654f, 0.7978845608028654f);
  const float hx = 0.5f * x;
  const float d = fmaf(hx * du, fmaf(-t, t, 1.0f), fmaf(0.5f, t, 0.5f));
  x = fmaf(hx, t, hx);
  return d;
}

// Optional timeline instrumentation (build with -DOSUDIT_GEMM_TRACE): CTA 0 records clock64() for tiles 4..11 —
// role 0 = MMA issuer (tile start / all MMAs issued), role 1 = epilogue thread 0 (per tile: accumulator ready, then
// per chunk: buffer free, TMEM read done, tile stored to smem, TMA store issued).  osudit_debug_gemm_trace() copies it out.
#ifdef OSUDIT_GEMM_TRACE
__device__ long long g_gemm_trace[2 * 8 * 20];
#define GEMM_TRACE(role, it, ev)                                                                    \
  do {                                                                                              \
    if (blockIdx.x == 0 && (it) >= 4 && (it) < 12) g_gemm_trace[((role) * 8 + (it) - 4) * 20 + (ev)] = clock64(); \
  } while (0)
#else
#define GEMM_TRACE(role, it, ev) do {} while (0)
#endif

template <int EPI, int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_kernel(const __grid_constant__ Params p) {
  constexpr int kBytesB = Cfg<BN>::kBytesB, kStageBytes = Cfg<BN>::kStageBytes;
  constexpr int kSt = stages_for(EPI);
  constexpr bool kAuxTma = EPI == EPI_BF16_DGELU;  // the saved GELU derivative is staged by TMA, two chunks ahead
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* aux_smem = smem + kSt * kStageBytes;  // [2][128 rows][64 cols] bf16, 128-byte swizzled (DGELU only)
  uint8_t* staging = aux_smem + ((kAuxTma || EPI == EPI_RESID) ? 2 * kStagingBytes : 0);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + 2 * kEpiGroups * kStagingBytes);
  uint64_t* empty_bar = full_bar + kSt;
  uint64_t* tmem_full = empty_bar + kSt;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* aux_full = tmem_empty + 2;  // [2] per CTA: an aux chunk has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_full + 2);
  float* s_bias_all = reinterpret_cast<float*>(staging + 2 * kEpiGroups * kStagingBytes + 1024);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_a);
    tma_prefetch_desc(&p.tma_b);
    tma_prefetch_desc(&p.tma_out);
    for (int i = 0; i < kSt; ++i) {
      mbar_init(&full_bar[i], 1);   // leader's: one arrive.expect_tx covering both CTAs' bytes
      mbar_init(&empty_bar[i], 1);  // per CTA: one multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);    // per CTA: one multicast commit
      mbar_init(&tmem_empty[i], 16 * kEpiGroups);  // leader's: every epilogue warp of both CTAs
      mbar_init(&aux_full[i], 1);
    }
    if (kAuxTma) tma_prefetch_desc(&p.tma_aux);
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
        const int m0 = (tile / p.n_tiles) * (2 * BM) + static_cast<int>(rank) * BM;
        const int n0 = (tile % p.n_tiles) * BN + static_cast<int>(rank) * (BN / 2);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          const uint32_t lbar = smem_u32(&full_bar[stage]) & 0xFEFFFFFFu;  // the leader's copy
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * kStageBytes);
          tma_load_2d_pair(sa, &p.tma_a, lbar, kb * BK, m0);
          tma_load_2d_pair(sa + kBytesA, &p.tma_b, lbar, kb * BK, n0);
          if (++stage == kSt) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer (leader only)
    // The whole warp runs the loop converged and one elected lane issues.  Inside an `if (lane == 0)` region the
    // compiler cannot prove the operands of tcgen05.mma / tcgen05.commit warp-uniform (they must sit in uniform
    // registers), and wraps every one of them in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop of ~16 instructions:
    // with four MMAs of 128 tensor cycles per k-block the issue then runs about as long as the execution.  Measured
    // (tools/gemm_bench.py): fc1+GELU 1362 -> 1436, QKV 1410 -> 1435 TFLOP/s, out-proj and fc2 unchanged.  The same
    // change made the 1-CTA kernel (64-cycle MMAs, a __syncwarp per k-block on its critical path) and the attention
    // backward slightly slower (training step 25.56 -> 25.84 ms), so those keep the single-lane loop.
    if (leader) {
      // kind::f16: D fp32, A/B bf16 K-major, N = BN, M = 256 (pair)
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
                                 (static_cast<uint32_t>((2 * BM) >> 4) << 24);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);  // provably warp-uniform
      const uint32_t sbase = smem_u32(smem);
      const uint64_t da0 = desc_sw128(sbase), db0 = desc_sw128(sbase + kBytesA);
      constexpr uint32_t kStageStep = kStageBytes >> 4;  // descriptor address field counts 16-byte units (< 256 KB: no carry)
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++it) {
        if (lane == 0) {
          GEMM_TRACE(0, it, 0);
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
          GEMM_TRACE(0, it, 1);
        }
        __syncwarp();
        tc_fence_after();
        const uint32_t d_tmem = tb + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          if (lane == 0) mbar_wait(&full_bar[stage], phase);
          __syncwarp();
          tc_fence_after();
          const uint64_t da = da0 + static_cast<uint32_t>(stage) * kStageStep;
          const uint64_t db = db0 + static_cast<uint32_t>(stage) * kStageStep;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == kSt) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit_pair(&tmem_full[acc]);
        __syncwarp();
        if (lane == 0) GEMM_TRACE(0, it, 2);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue (both CTAs)
    constexpr int kChunks = BN / 64;
    const int grp = (warp - 2) >> 3;                 // epilogue warp group
    const int ep_tid = (threadIdx.x - 64) & 255;     // thread index within the group
    const int quad = warp & 3;
    const int half = ((warp - 2) >> 2) & 1;
    uint8_t* const my_staging = staging + grp * 2 * kStagingBytes;
    const int row = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t chunk_ctr = 0;
    int it = 0;
    auto aux_issue = [&](uint32_t g) {  // TMA load of running chunk g of this CTA (nothing past the last tile)
      const int tile2 = cluster_id + static_cast<int>(g / kChunks) * n_clusters;
      if (tile2 >= total_tiles) return;
      const int m2 = (tile2 / p.n_tiles) * (2 * BM) + static_cast<int>(rank) * BM;
      const int n2 = (tile2 % p.n_tiles) * BN + static_cast<int>(g % kChunks) * 64;
      mbar_expect_tx(&aux_full[g & 1], kStagingBytes);
      tma_load_2d(aux_smem + (g & 1) * kStagingBytes, &p.tma_aux, &aux_full[g & 1], n2, m2);
    };
    if (kAuxTma && ep_tid == 0) {
      aux_issue(0);
      aux_issue(1);
    }
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++it) {
      const int m0 = (tile / p.n_tiles) * (2 * BM) + static_cast<int>(rank) * BM;
      const int n0 = (tile % p.n_tiles) * BN;
      // EPI_BF16_DGELU: the chunk's [128 rows][64 columns] of the saved GELU derivative arrive by TMA two chunks
      // ahead (issued below, after the barrier that ends a chunk's reads of its buffer).  Per-thread global loads of
      // the thread's own row cost 32 cache lines per load instruction and held this GEMM at 43 % tensor-pipe activity.
      // this tile's 256 bias values go through shared memory (one global load per thread per tile, issued before
      // the wait for the accumulator) instead of 8 dependent __ldg per thread per chunk on the critical path
      // (EPI_RESID: bias in the first buffer, this tile's gate slice in the second; the per-chunk barriers order the
      // previous tile's reads before these writes)
      float* s_bias = s_bias_all + (grp * 2 + (EPI == EPI_RESID ? 0 : (it & 1))) * BN;
      if (ep_tid < BN) s_bias[ep_tid] = p.bias != nullptr ? __ldg(p.bias + n0 + ep_tid) : 0.f;
      if (EPI == EPI_RESID && ep_tid < BN) {
        const int b = min(m0 / p.rows_per_batch, p.batches - 1);  // the CTA's 128 rows lie in one batch row
        s_bias[BN + ep_tid] = __ldg(p.gate + static_cast<int64_t>(b) * p.gate_ld + n0 + ep_tid);
      }
      if (ep_tid == 0) GEMM_TRACE(1, it, 0);
      warp_mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      named_bar_sync(1 + grp, 256);
      if (ep_tid == 0) GEMM_TRACE(1, it, 1);
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * BN);
#pragma unroll 1
      for (int c = grp; c < kChunks; c += kEpiGroups) {
        const int ncol0 = n0 + c * 64;
        uint4 ax[4];
        const uint32_t gchunk = static_cast<uint32_t>(it) * kChunks + c;  // running chunk index of this CTA
        if (kAuxTma) {
          warp_mbar_wait(&aux_full[gchunk & 1], (gchunk >> 1) & 1);
          const uint32_t arow = smem_u32(aux_smem + (gchunk & 1) * kStagingBytes) + row * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) ax[j] = lds128(arow + (((half * 4 + j) ^ (row & 7)) << 4));
        }
        uint32_t r[32];
        if (ep_tid == 0) GEMM_TRACE(1, it, 2 + 4 * c);
        tmem_ld_32x32(t_row + static_cast<uint32_t>(c * 64 + half * 32), r);
        tmem_ld_wait();
        if (c + kEpiGroups >= kChunks) {  // this warp's last read of the accumulator: hand the stage back
          tc_fence_before();
          if (ep_tid == 0 && c == 3) GEMM_TRACE(1, it, 18);
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_rank0(smem_u32(&tmem_empty[acc])));
        }
        if (ep_tid == 0 && c == 3) GEMM_TRACE(1, it, 19);
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b = *reinterpret_cast<const float4*>(s_bias + c * 64 + half * 32 + i);
          v[i + 0] = __uint_as_float(r[i + 0]) + b.x;
          v[i + 1] = __uint_as_float(r[i + 1]) + b.y;
          v[i + 2] = __uint_as_float(r[i + 2]) + b.z;
          v[i + 3] = __uint_as_float(r[i + 3]) + b.w;
        }
        if (EPI == EPI_RESID) {
          const float* s_gate = s_bias + BN + c * 64 + half * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 g = *reinterpret_cast<const float4*>(s_gate + i);
            v[i + 0] *= g.x;
            v[i + 1] *= g.y;
            v[i + 2] *= g.z;
            v[i + 3] *= g.w;
          }
          // two [128 rows][32 columns] fp32 boxes (one per column half): the group's staging pair and the pair in
          // the auxiliary area alternate per chunk, so one chunk's reduce-add may still be reading while the next is
          // written (kEpiGroups == 1 is the shipped configuration; with two groups both would share the second pair)
          static_assert(EPI != EPI_RESID || kEpiGroups == 1, "EPI_RESID: one epilogue group");
          uint8_t* pair = (chunk_ctr & 1) ? aux_smem : my_staging;
          ++chunk_ctr;
          if (ep_tid == 0) tma_store_wait_read<1>();
          named_bar_sync(1 + grp, 256);
          uint8_t* my_row = pair + half * kStagingBytes + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(my_row + ((j ^ (row & 7)) << 4)) =
                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          fence_proxy_async_smem();
          named_bar_sync(1 + grp, 256);
          if (ep_tid == 0) {
            tma_reduce_add_2d(&p.tma_out, pair, ncol0, m0);
            tma_reduce_add_2d(&p.tma_out, pair + kStagingBytes, ncol0 + 32, m0);
            tma_store_commit();
          }
          continue;
        }
        if (EPI == EPI_BF16_DGELU) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t* w = reinterpret_cast<const uint32_t*>(&ax[j]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
              v[8 * j + 2 * k] *= d.x;
              v[8 * j + 2 * k + 1] *= d.y;
            }
          }
        }
        constexpr int kPasses = EPI == EPI_BF16_GELU_SAVE ? 2 : 1;  // SAVE: the derivative first, then the GELU
        float dv[EPI == EPI_BF16_GELU_SAVE ? 32 : 1];
        if (EPI == EPI_BF16_GELU_SAVE) {
#pragma unroll
          for (int i = 0; i < 32; ++i) dv[i] = gelu_tanh_with_grad(v[i]);
        }
#pragma unroll
        for (int pass = 0; pass < kPasses; ++pass, ++chunk_ctr) {
          uint8_t* buf = my_staging + (chunk_ctr & 1) * kStagingBytes;
          if (ep_tid == 0) tma_store_wait_read<1>();
          named_bar_sync(1 + grp, 256);
          if (kAuxTma && ep_tid == 0) aux_issue(gchunk + 2);  // every thread has consumed this chunk's aux buffer
          if (ep_tid == 0) GEMM_TRACE(1, it, 2 + 4 * c + 1);
          uint8_t* my_row = buf + row * 128;
#ifndef OSUDIT_GELU_H2
          if (EPI == EPI_BF16_GELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_tanh(v[i]);
          }
#endif
          const float* src = (EPI == EPI_BF16_GELU_SAVE && pass == 0) ? dv : v;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
#ifdef OSUDIT_GELU_H2
            if (EPI == EPI_BF16_GELU) {
              o.x = gelu_tanh_pair(src[8 * j + 0], src[8 * j + 1]);
              o.y = gelu_tanh_pair(src[8 * j + 2], src[8 * j + 3]);
              o.z = gelu_tanh_pair(src[8 * j + 4], src[8 * j + 5]);
              o.w = gelu_tanh_pair(src[8 * j + 6], src[8 * j + 7]);
            } else
#endif
            {
              o.x = pack_bf16(src[8 * j + 0], src[8 * j + 1]);
              o.y = pack_bf16(src[8 * j + 2], src[8 * j + 3]);
              o.z = pack_bf16(src[8 * j + 4], src[8 * j + 5]);
              o.w = pack_bf16(src[8 * j + 6], src[8 * j + 7]);
            }
            const int jj = half * 4 + j;
            *reinterpret_cast<uint4*>(my_row + ((jj ^ (row & 7)) << 4)) = o;
          }
          if (ep_tid == 0) GEMM_TRACE(1, it, 2 + 4 * c + 2);
          fence_proxy_async_smem();
          named_bar_sync(1 + grp, 256);
          if (ep_tid == 0) GEMM_TRACE(1, it, 2 + 4 * c + 3);
#ifndef OSUDIT_GEMM_NOSTORE  // (timing experiment: the epilogue without its global writes)
          if (ep_tid == 0) {
            tma_store_2d((EPI == EPI_BF16_GELU_SAVE && pass == 0) ? &p.tma_aux : &p.tma_out, buf, ncol0, m0);
            tma_store_commit();
          }
#endif
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (ep_tid == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  cluster_sync_all();  // the peer may still be signalling this CTA's barriers / reading its smem
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

template <int EPI, int BN>
static int launch2(const Params& p, cudaStream_t stream) {
  constexpr int kSmemBytes = smem_bytes_for<BN>(EPI);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_kernel<EPI, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return set_error(-5, cudaGetErrorString(e));
    configured = true;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  int clusters = num_sms() / 2;
  if (tiles < clusters) clusters = tiles;
  gemm2_kernel<EPI, BN><<<2 * clusters, kThreads, kSmemBytes, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(-6, cudaGetErrorString(e));
  return 0;
}

}  // namespace g2

#ifdef OSUDIT_GEMM_TRACE
extern "C" int osudit_debug_gemm_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g2::g_gemm_trace, sizeof(g2::g_gemm_trace)) == cudaSuccess ? 0 : -1;
}
#endif

bool gemm_2cta_resid_applicable(int64_t M, int64_t N, int64_t rows_per_batch) {
  if (!(N % 256 == 0 || N % 192 == 0) || rows_per_batch <= 0 || rows_per_batch % g2::BM != 0 || M % rows_per_batch != 0)
    return false;
  const int64_t tiles = ((M + 255) / 256) * (N % 256 == 0 ? N / 256 : N / 192);
  return M >= 256 && tiles >= 37;
}

// x[M, N] (fp32) += gate[row / rows_per_batch, :] * (a[M, K] b[N, K]^T + bias)
int gemm_2cta_resid_launch(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t K, int64_t M, int64_t N,
                           const float* bias, const float* gate, int64_t gate_ld, int64_t rows_per_batch, float* x,
                           int64_t ldx, cudaStream_t stream) {
  using namespace g2;
  const int BN = N % 256 == 0 ? 256 : 192;
  Params p;
  p.kblocks = static_cast<int>((K + BK - 1) / BK);
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(N);
  p.m_tiles = static_cast<int>((M + 2 * BM - 1) / (2 * BM));
  p.n_tiles = static_cast<int>(N / BN);
  p.bias = bias;
  p.aux = nullptr;
  p.ld_aux = 0;
  p.gate = gate;
  p.gate_ld = gate_ld;
  p.rows_per_batch = static_cast<int>(rows_per_batch);
  p.batches = static_cast<int>(M / rows_per_batch);
  int rc = make_tensor_map_2d(&p.tma_a, a, K, M, lda * 2, BK, BM, false);
  if (rc) return rc;
  rc = make_tensor_map_2d(&p.tma_b, b, K, N, ldb * 2, BK, BN / 2, false);
  if (rc) return rc;
  rc = make_tensor_map_2d(&p.tma_out, x, N, M, ldx * 4, 32, BM, true);  // fp32 boxes of 32 columns = 128 bytes
  if (rc) return rc;
  p.tma_aux = p.tma_out;
  return BN == 256 ? launch2<EPI_RESID, 256>(p, stream) : launch2<EPI_RESID, 192>(p, stream);
}

bool gemm_2cta_applicable(int nseg, int64_t M, int64_t N, int epilogue) {
  if (!(nseg == 1 && epilogue >= g2::EPI_BF16 && epilogue <= g2::EPI_BF16_DGELU && (N % 256 == 0 || N % 192 == 0)))
    return false;
  // at least half a wave of 256-row output tiles over the 74 CTA pairs, otherwise the 1-CTA kernel balances better
  // (a strong-scaling training rank has M = 4096 rows: 16 row panels x 3..12 column tiles)
  const int64_t tiles = ((M + 255) / 256) * (N % 256 == 0 ? N / 256 : N / 192);
  return M >= 256 && tiles >= 37;
}

int gemm_2cta_launch(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t K, int64_t M, int64_t N,
                     const float* bias, int epilogue, void* out, int64_t ldo, void* aux, int64_t ld_aux,
                     cudaStream_t stream) {
  using namespace g2;
  const int BN = N % 256 == 0 ? 256 : 192;
  Params p;
  p.kblocks = static_cast<int>((K + BK - 1) / BK);
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(N);
  p.m_tiles = static_cast<int>((M + 2 * BM - 1) / (2 * BM));
  p.n_tiles = static_cast<int>(N / BN);
  p.bias = bias;
  int rc = make_tensor_map_2d(&p.tma_a, a, K, M, lda * 2, BK, BM, false);
  if (rc) return rc;
  rc = make_tensor_map_2d(&p.tma_b, b, K, N, ldb * 2, BK, BN / 2, false);
  if (rc) return rc;
  rc = make_tensor_map_2d(&p.tma_out, out, N, M, ldo * 2, 64, BM, false);
  if (rc) return rc;
  p.aux = static_cast<const __nv_bfloat16*>(aux);
  p.ld_aux = ld_aux;
  p.gate = nullptr;
  p.gate_ld = 0;
  p.rows_per_batch = 1;
  p.batches = 1;
  p.tma_aux = p.tma_out;
  if (epilogue == EPI_BF16_GELU_SAVE || epilogue == EPI_BF16_DGELU) {
    if (aux == nullptr || (ld_aux % 8) != 0) return set_error(-1, "gemm: aux must be given with ld_aux % 8 == 0");
    rc = make_tensor_map_2d(&p.tma_aux, aux, N, M, ld_aux * 2, 64, BM, false);  // SAVE: stores; DGELU: loads
    if (rc) return rc;
  }
  switch (epilogue) {
    case EPI_BF16: return BN == 256 ? launch2<EPI_BF16, 256>(p, stream) : launch2<EPI_BF16, 192>(p, stream);
    case EPI_BF16_GELU: return BN == 256 ? launch2<EPI_BF16_GELU, 256>(p, stream) : launch2<EPI_BF16_GELU, 192>(p, stream);
    case EPI_BF16_GELU_SAVE:
      return BN == 256 ? launch2<EPI_BF16_GELU_SAVE, 256>(p, stream) : launch2<EPI_BF16_GELU_SAVE, 192>(p, stream);
    default: return BN == 256 ? launch2<EPI_BF16_DGELU, 256>(p, stream) : launch2<EPI_BF16_DGELU, 192>(p, stream);
  }
}

}  // namespace osudit
