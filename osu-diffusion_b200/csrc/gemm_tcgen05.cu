// C[M,N] = sum_s A_s[M,K_s] * B_s[N,K_s]^T (+ bias) with an activation/convert epilogue.
//
// The dense contractions of the DiT denoiser (reference models.py:164-170 QKV/out-proj,
// models.py:112-119 MLP, models.py:233-234 first layer, models.py:152-159,193 adaLN,
// models.py:35-38 t-MLP): nn.Linear stores W as [out, in], activations are [tokens, in], so
// both operands are K-major and feed tcgen05.mma directly.
//
// sm_100a design (one CTA per SM, persistent over 128 x BN output tiles):
//   warp 0      TMA producer: cp.async.bulk.tensor 2D loads, 128B-swizzled [rows][64 bf16] boxes
//   warp 1      owns TMEM (2 accumulator stages x BN fp32 columns) and issues tcgen05.mma
//               (M=128, N=BN, K=16, kind::f16 = bf16 in / fp32 accumulate in TMEM)
//   warps 2..5  epilogue: tcgen05.ld 32x32b -> +bias -> (GELU-tanh) -> fp32|bf16 -> swizzled smem
//               -> TMA store (clips the M/N tails), double-buffered against the next tile's MMAs
// Pipelines: smem full/empty mbarriers (kStages deep), TMEM full/empty mbarriers (2 deep).
// Up to three (A,B) K-segments are chained into one accumulator: that is how the split-bf16
// ("bf16x3": hi*hi + lo*hi + hi*lo) first-layer / adaLN products run on the same kernel.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.h"
#include "ptx.cuh"

namespace osudit {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kMaxSeg = 3;
constexpr int kStageBytesA = BM * BK * 2;
constexpr int kStagingBytes = BM * 128;  // one epilogue chunk: 128 rows x 128 B

struct GemmParams {
  CUtensorMap tma_a[kMaxSeg];
  CUtensorMap tma_b[kMaxSeg];
  CUtensorMap tma_out;
  int kblocks[kMaxSeg];
  int nseg;
  int M, N;
  int m_tiles, n_tiles;
  // EPI_WGRAD: the K range (token dimension) is split across CTAs.  EPI_F32 with k_splits > 1 ("precise"
  // accumulation, osudit_gemm_bf16_splitk): the concatenated k-block sequence of all segments is cut into
  // short chains whose partial tiles are reduce-added in fp32 (round-to-nearest) into a zeroed output.
  int k_splits, kb_per_split;
  const float* bias;
};

template <int BN>
struct Cfg {
  static constexpr int kStageBytesB = BN * BK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kBarrierBytes = 1024;
  static constexpr int kSmemBytes =
      kStages * kStageBytes + 2 * kStagingBytes + kBarrierBytes + 1024 /*align slack*/;
};

// UMMA shared-memory descriptor, K-major operand, SWIZZLE_128B, rows 128 B apart, 8-row groups
// 1024 B apart (SBO); LBO is unused for swizzled K-major layouts (encoded 1); version 1 = sm_100.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, dense.
template <int BN>
__device__ __forceinline__ constexpr uint32_t umma_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
         (static_cast<uint32_t>(BM >> 4) << 24);
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

// EPI_WGRAD: out[M,N] (fp32, reduce-added) += A[K,M]^T . B[K,N] with BOTH operands token-major
// (rows = the contraction index): the weight-gradient GEMM dW = dY^T X read straight from dY and X
// as MN-major UMMA operands, K split across CTAs, partial tiles combined by TMA reduce-add.
enum : int { EPI_F32 = 0, EPI_BF16 = 1, EPI_BF16_GELU = 2, EPI_WGRAD = 3 };

__host__ __device__ constexpr bool epi_is_f32(int epi) { return epi == EPI_F32 || epi == EPI_WGRAD; }
__host__ __device__ constexpr int epi_threads(int epi) { return epi_is_f32(epi) ? 128 : 256; }

template <int BN, int EPI>
__global__ void __launch_bounds__(64 + epi_threads(EPI), 1)
gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  using C = Cfg<BN>;
  constexpr int kEpiCols = epi_is_f32(EPI) ? 32 : 64;  // 128 bytes of output per row per chunk
  constexpr bool kWgrad = EPI == EPI_WGRAD;
  constexpr int kChunks = BN / kEpiCols;
  constexpr int kEpiThreads = epi_threads(EPI);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* staging = smem + C::kStages * C::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + 2 * kStagingBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tmem_full = empty_bar + C::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles * p.k_splits;  // work items
  const bool splitk = !kWgrad && p.k_splits > 1;
  // Work item -> (output tile, K split).  Weight gradients run split-major: all output tiles of one token range are
  // in flight together, so the CTAs of a wave share their dY / X slices in L2 (tile-major order re-read the operands
  // 2.7x from HBM: 678 MB for fc2's 251 MB, ncu r02); the other modes keep the splits of a tile adjacent.
  const int n_out_tiles = p.m_tiles * p.n_tiles;
  auto tile_of = [&](int work) { return kWgrad ? work % n_out_tiles : work / p.k_splits; };
  auto split_of = [&](int work) { return kWgrad ? work / n_out_tiles : work % p.k_splits; };
  int total_kblocks = 0;
  for (int s = 0; s < p.nseg; ++s) total_kblocks += p.kblocks[s];

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) {
      tma_prefetch_desc(&p.tma_a[s]);
      tma_prefetch_desc(&p.tma_b[s]);
    }
    tma_prefetch_desc(&p.tma_out);
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kEpiThreads);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<C::kTmemCols>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int work = blockIdx.x; work < total_tiles; work += gridDim.x) {
        const int tile = tile_of(work);
        const int m0 = (tile / p.n_tiles) * BM;
        const int n0 = (tile % p.n_tiles) * BN;
        if (kWgrad) {
          const int kb0 = split_of(work) * p.kb_per_split;
          const int kb1 = min(kb0 + p.kb_per_split, p.kblocks[0]);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * C::kStageBytes;
            mbar_expect_tx(&full_bar[stage], C::kStageBytes);
            // [64 tokens][64 columns] boxes: rows = K, 128 bytes of M (or N) per row
#pragma unroll
            for (int hh = 0; hh < BM / 64; ++hh)
              tma_load_2d(sa + hh * 8192, &p.tma_a[0], &full_bar[stage], m0 + 64 * hh, kb * BK);
#pragma unroll
            for (int hh = 0; hh < BN / 64; ++hh)
              tma_load_2d(sa + kStageBytesA + hh * 8192, &p.tma_b[0], &full_bar[stage], n0 + 64 * hh, kb * BK);
            if (++stage == C::kStages) { stage = 0; phase ^= 1; }
          }
          continue;
        }
        int g_lo = 0, g_hi = total_kblocks;  // this work item's range of the concatenated k-blocks
        if (splitk) {
          g_lo = split_of(work) * p.kb_per_split;
          g_hi = min(g_lo + p.kb_per_split, total_kblocks);
        }
        int g = 0;
        for (int s = 0; s < p.nseg; ++s) {
          for (int kb = 0; kb < p.kblocks[s]; ++kb, ++g) {
            if (g < g_lo || g >= g_hi) continue;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * C::kStageBytes;
            mbar_expect_tx(&full_bar[stage], C::kStageBytes);
            tma_load_2d(sa, &p.tma_a[s], &full_bar[stage], kb * BK, m0);
            tma_load_2d(sa + kStageBytesA, &p.tma_b[s], &full_bar[stage], kb * BK, n0);
            if (++stage == C::kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      // MN-major operands (EPI_WGRAD): bits 15/16 of the instruction descriptor
      constexpr uint32_t idesc = umma_idesc<BN>() | (kWgrad ? (3u << 15) : 0u);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int work = blockIdx.x; work < total_tiles; work += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        int n_kb = total_kblocks;
        if (kWgrad) {
          const int kb0 = split_of(work) * p.kb_per_split;
          n_kb = min(kb0 + p.kb_per_split, p.kblocks[0]) - kb0;
        } else if (splitk) {
          const int g_lo = split_of(work) * p.kb_per_split;
          n_kb = min(g_lo + p.kb_per_split, total_kblocks) - g_lo;
        }
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          if (kWgrad) {
            // MN-major SWIZZLE_128B: 64-element atoms along M/N are 8 KB apart (LBO), 8-row K groups
            // 1 KB apart (SBO); one K=16 step = 16 rows = 2 KB
            const uint64_t lbo = static_cast<uint64_t>((8192 >> 4) - 1) << 16;  // umma_desc sets 1
            const uint64_t da = umma_desc_sw128(sa) + lbo;
            const uint64_t db = umma_desc_sw128(sa + kStageBytesA) + lbo;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16(d_tmem, da + 128 * k, db + 128 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          } else {
            const uint64_t da = umma_desc_sw128(sa);
            const uint64_t db = umma_desc_sw128(sa + kStageBytesA);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in 16 B units
              umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    // fp32 output: 4 warps, one per TMEM lane quadrant, 32 columns (128 B) per chunk.
    // bf16 output: 8 warps, two per quadrant; each takes 32 of the chunk's 64 columns, which keeps
    // two warps per scheduler in flight to hide the TMEM-load / MUFU (GELU) latencies.
    const int ep_tid = threadIdx.x - 64;
    const int quad = warp & 3;           // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;    // which 32-column half of a bf16 chunk (0 for fp32)
    const int row = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t chunk_ctr = 0;
    for (int work = blockIdx.x; work < total_tiles; work += gridDim.x) {
      const int tile = tile_of(work);
      const int m0 = (tile / p.n_tiles) * BM;
      const int n0 = (tile % p.n_tiles) * BN;
      const bool add_bias = p.bias != nullptr && split_of(work) == 0;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                             static_cast<uint32_t>(acc * BN);
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c, ++chunk_ctr) {
        uint8_t* buf = staging + (chunk_ctr & 1) * kStagingBytes;
        // the store that last read this buffer (2 chunks ago) must have drained it
        if (ep_tid == 0) tma_store_wait_read<1>();
        named_bar_sync(1, kEpiThreads);
        const int ncol0 = n0 + c * kEpiCols;
        const uint32_t my_row = smem_u32(buf) + row * 128;  // shared-space address (generic st would be ST.E)
        uint32_t r[32];
        tmem_ld_32x32(t_row + static_cast<uint32_t>(c * kEpiCols + half * 32), r);
        tmem_ld_wait();
        if (c == kChunks - 1) {
          // accumulator fully read: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          mbar_arrive(&tmem_empty[acc]);
        }
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const int n = ncol0 + half * 32 + i;
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (add_bias && n + 3 < p.N) b = __ldg(reinterpret_cast<const float4*>(p.bias + n));
          v[i + 0] = __uint_as_float(r[i + 0]) + b.x;
          v[i + 1] = __uint_as_float(r[i + 1]) + b.y;
          v[i + 2] = __uint_as_float(r[i + 2]) + b.z;
          v[i + 3] = __uint_as_float(r[i + 3]) + b.w;
        }
        if (EPI == EPI_BF16_GELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = gelu_tanh(v[i]);
        }
        if (epi_is_f32(EPI)) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {  // 8 x 16 B
            sts128(my_row + ((j ^ (row & 7)) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                   __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {  // 4 x 16 B (8 bf16 each) per 32 columns
            uint4 o;
            o.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]);
            o.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
            o.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
            o.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
            const int jj = half * 4 + j;
            sts128(my_row + ((jj ^ (row & 7)) << 4), o.x, o.y, o.z, o.w);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, kEpiThreads);
        if (ep_tid == 0) {
          if (kWgrad || splitk) tma_reduce_add_2d(&p.tma_out, buf, ncol0, m0);
          else tma_store_2d(&p.tma_out, buf, ncol0, m0);
          tma_store_commit();
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (ep_tid == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------- host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t d0, d1, d2, stride1, stride2;
  uint32_t b0, b1, dtype;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && stride1 == o.stride1 &&
           stride2 == o.stride2 && b0 == o.b0 && b1 == o.b1 && dtype == o.dtype;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.d0); mix(k.d1); mix(k.d2); mix(k.stride1); mix(k.stride2); mix(k.b0); mix(k.b1); mix(k.dtype);
    return h;
  }
};

// Row-major tensor [d2][d1 rows][d0 cols] (d2 == 0: plain 2-D), row pitch `stride1_bytes`, plane
// pitch `stride2_bytes`; box = [1][b1 rows][b0 cols], 128-byte swizzle, zero fill out of bounds.
// Descriptors depend only on the key, so a process-wide cache is safe.
static int make_tensor_map_nd(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2,
                              uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1,
                              bool is_f32) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, d0, d1, d2, stride1_bytes, stride2_bytes, b0, b1, is_f32 ? 1u : 0u};
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(-3, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (stride1_bytes & 15) || (stride2_bytes & 15))
    return set_error(-2, "TMA operand must be 16-byte aligned with 16-byte-multiple pitches");
  const cuuint32_t rank = d2 ? 3 : 2;
  cuuint64_t dims[3] = {d0, d1, d2 ? d2 : 1};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                   rank, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(-4, "cuTensorMapEncodeTiled failed");
  std::lock_guard<std::mutex> lock(mu);
  if (cache.size() > 65536) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

static int make_tensor_map(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1,
                           uint64_t stride_bytes, uint32_t b0, uint32_t b1, bool is_f32) {
  return make_tensor_map_nd(out, ptr, d0, d1, 0, stride_bytes, 0, b0, b1, is_f32);
}

int make_tensor_map_2d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t stride_bytes,
                       uint32_t b0, uint32_t b1, bool is_f32) {
  return make_tensor_map_nd(out, ptr, d0, d1, 0, stride_bytes, 0, b0, b1, is_f32);
}

bool gemm_2cta_applicable(int nseg, int64_t M, int64_t N, int epilogue);
int gemm_2cta_launch(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t K, int64_t M, int64_t N,
                     const float* bias, int epilogue, void* out, int64_t ldo, void* aux, int64_t ld_aux,
                     cudaStream_t stream);

int make_tensor_map_3d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2,
                       uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1) {
  return make_tensor_map_nd(out, ptr, d0, d1, d2, stride1_bytes, stride2_bytes, b0, b1, false);
}

template <int BN, int EPI>
static int launch(const GemmParams& p, cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool configured = false;
  auto kern = gemm_tcgen05_kernel<BN, EPI>;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return set_error(-5, cudaGetErrorString(e));
    configured = true;
  }
  const int tiles = p.m_tiles * p.n_tiles * p.k_splits;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, 64 + epi_threads(EPI), C::kSmemBytes, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(-6, cudaGetErrorString(e));
  return 0;
}

}  // namespace osudit

using namespace osudit;

static int gemm_entry(int nseg, const void* const* a, const int64_t* lda, const void* const* b,
                      const int64_t* ldb, const int64_t* k, int64_t M, int64_t N, const float* bias,
                      int epilogue, void* out, int64_t ldo, void* stream, int kb_per_split) {
  if (nseg < 1 || nseg > kMaxSeg) return set_error(-1, "gemm: nseg must be 1..3");
  if (M <= 0 || N <= 0 || (N % 8) != 0) return set_error(-1, "gemm: need M>0, N>0, N%8==0");
  if (epilogue < 0 || epilogue > 2) return set_error(-1, "gemm: unknown epilogue");
  {
    // large bf16-output GEMMs: CTA-pair kernel (gemm_2cta.cu); OSUDIT_GEMM_2CTA=0 forces the 1-CTA one
    static const bool use_pair = [] {
      const char* e = getenv("OSUDIT_GEMM_2CTA");
      return !(e && e[0] == '0');
    }();
    if (use_pair && kb_per_split == 0 && gemm_2cta_applicable(nseg, M, N, epilogue)) {
      if (k[0] <= 0 || (k[0] % 8) != 0) return set_error(-1, "gemm: K must be a positive multiple of 8");
      return gemm_2cta_launch(a[0], lda[0], b[0], ldb[0], k[0], M, N, bias, epilogue, out, ldo, nullptr, 0,
                              static_cast<cudaStream_t>(stream));
    }
  }
  const int BN = (N % 256 == 0) ? 256 : 128;
  GemmParams p;
  p.nseg = nseg;
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(N);
  p.m_tiles = static_cast<int>((M + BM - 1) / BM);
  p.n_tiles = static_cast<int>((N + BN - 1) / BN);
  p.bias = bias;
  p.k_splits = 1;
  p.kb_per_split = 0;
  for (int s = 0; s < kMaxSeg; ++s) p.kblocks[s] = 0;
  for (int s = 0; s < nseg; ++s) {
    if (k[s] <= 0 || (k[s] % 8) != 0) return set_error(-1, "gemm: K must be a positive multiple of 8");
    p.kblocks[s] = static_cast<int>((k[s] + BK - 1) / BK);
    int rc = make_tensor_map(&p.tma_a[s], a[s], k[s], M, lda[s] * 2, BK, BM, false);
    if (rc) return rc;
    rc = make_tensor_map(&p.tma_b[s], b[s], k[s], N, ldb[s] * 2, BK, BN, false);
    if (rc) return rc;
  }
  const bool f32 = epilogue == EPI_F32;
  int rc = make_tensor_map(&p.tma_out, out, N, M, ldo * (f32 ? 4 : 2), f32 ? 32 : 64, BM, f32);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (kb_per_split > 0) {  // precise accumulation: short tensor-core chains, fp32 reduce-add into zeros
    if (!f32) return set_error(-1, "gemm_splitk: fp32 output only");
    int total = 0;
    for (int s = 0; s < nseg; ++s) total += p.kblocks[s];
    p.kb_per_split = kb_per_split;
    p.k_splits = (total + kb_per_split - 1) / kb_per_split;
    cudaError_t e = cudaMemset2DAsync(out, static_cast<size_t>(ldo) * 4, 0, static_cast<size_t>(N) * 4,
                                      static_cast<size_t>(M), st);
    if (e != cudaSuccess) return set_error(-6, cudaGetErrorString(e));
  }
  if (BN == 256) {
    if (epilogue == EPI_F32) return launch<256, EPI_F32>(p, st);
    if (epilogue == EPI_BF16) return launch<256, EPI_BF16>(p, st);
    return launch<256, EPI_BF16_GELU>(p, st);
  }
  if (epilogue == EPI_F32) return launch<128, EPI_F32>(p, st);
  if (epilogue == EPI_BF16) return launch<128, EPI_BF16>(p, st);
  return launch<128, EPI_BF16_GELU>(p, st);
}

extern "C" int osudit_gemm_bf16(int nseg, const void* const* a, const int64_t* lda,
                                const void* const* b, const int64_t* ldb, const int64_t* k,
                                int64_t M, int64_t N, const float* bias, int epilogue, void* out,
                                int64_t ldo, void* stream) {
  return gemm_entry(nseg, a, lda, b, ldb, k, M, N, bias, epilogue, out, ldo, stream, 0);
}

namespace osudit {
int gelu_aux_launch(void* aux, void* out, int64_t n, int mode, cudaStream_t st);  // backward.cu
}

extern "C" int osudit_gemm_bf16_aux(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t K, int64_t M,
                                    int64_t N, const float* bias, int epilogue, void* out, int64_t ldo, void* aux,
                                    int64_t ld_aux, void* stream) {
  if (epilogue != 3 && epilogue != 4) return set_error(-1, "gemm_aux: epilogue must be GELU_SAVE (3) or DGELU (4)");
  if (aux == nullptr || out == nullptr) return set_error(-1, "gemm_aux: out and aux are required");
  if (M <= 0 || N <= 0 || K <= 0 || (K % 8) || (N % 8)) return set_error(-1, "gemm_aux: bad shape");
  static const bool use_pair = [] {
    const char* e = getenv("OSUDIT_GEMM_2CTA");
    return !(e && e[0] == '0');
  }();
  if (use_pair && gemm_2cta_applicable(1, M, N, epilogue))
    return gemm_2cta_launch(a, lda, b, ldb, K, M, N, bias, epilogue, out, ldo, aux, ld_aux,
                            static_cast<cudaStream_t>(stream));
  // shapes the CTA-pair kernel does not take: the same result in two launches
  if (ldo != N || ld_aux != N) return set_error(-1, "gemm_aux: the two-launch path needs contiguous out / aux");
  const void* as[1] = {a};
  const void* bs[1] = {b};
  const int64_t ldas[1] = {lda}, ldbs[1] = {ldb}, ks[1] = {K};
  if (epilogue == 3) {
    int rc = gemm_entry(1, as, ldas, bs, ldbs, ks, M, N, bias, EPI_BF16, aux, ld_aux, stream, 0);
    if (rc) return rc;
    return gelu_aux_launch(aux, out, M * N, 0, static_cast<cudaStream_t>(stream));
  }
  int rc = gemm_entry(1, as, ldas, bs, ldbs, ks, M, N, bias, EPI_BF16, out, ldo, stream, 0);
  if (rc) return rc;
  return gelu_aux_launch(aux, out, M * N, 1, static_cast<cudaStream_t>(stream));
}

namespace osudit {
bool gemm_2cta_resid_applicable(int64_t M, int64_t N, int64_t rows_per_batch);
int gemm_2cta_resid_launch(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t K, int64_t M, int64_t N,
                           const float* bias, const float* gate, int64_t gate_ld, int64_t rows_per_batch, float* x,
                           int64_t ldx, cudaStream_t stream);
}

extern "C" int osudit_gemm_gated_residual_applicable(int64_t M, int64_t N, int64_t rows_per_batch) {
  return gemm_2cta_resid_applicable(M, N, rows_per_batch) ? 1 : 0;
}

extern "C" int osudit_gemm_gated_residual(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t K, int64_t M,
                                          int64_t N, const float* bias, const float* gate, int64_t gate_ld,
                                          int64_t rows_per_batch, float* x, int64_t ldx, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || (K % 8) || (N % 8) || (lda % 8) || (ldb % 8) || (ldx % 4))
    return set_error(-1, "gemm_gated_residual: bad shape");
  if (a == nullptr || b == nullptr || gate == nullptr || x == nullptr) return set_error(-1, "gemm_gated_residual: null operand");
  if (!gemm_2cta_resid_applicable(M, N, rows_per_batch))
    return set_error(-1, "gemm_gated_residual: needs N % 256 == 0 or N % 192 == 0, rows_per_batch % 128 == 0 dividing M, "
                         "and at least 37 output tiles (ask osudit_gemm_gated_residual_applicable first)");
  return gemm_2cta_resid_launch(a, lda, b, ldb, K, M, N, bias, gate, gate_ld, rows_per_batch, x, ldx,
                                static_cast<cudaStream_t>(stream));
}

extern "C" int osudit_gemm_bf16_splitk(int nseg, const void* const* a, const int64_t* lda,
                                       const void* const* b, const int64_t* ldb, const int64_t* k,
                                       int64_t M, int64_t N, const float* bias, int kb_per_split,
                                       float* out, int64_t ldo, void* stream) {
  if (kb_per_split < 1) return set_error(-1, "gemm_splitk: kb_per_split must be >= 1");
  return gemm_entry(nseg, a, lda, b, ldb, k, M, N, bias, EPI_F32, out, ldo, stream, kb_per_split);
}

// out[M,N] (fp32) += dY[rows,M]^T . X[rows,N]: the weight gradient of y = x W^T (W is [M,N] = [out,in]),
// read straight from the token-major activations (no transposes), K = rows split across CTAs.
extern "C" int osudit_gemm_wgrad(const void* dy, int64_t ld_dy, const void* x, int64_t ld_x, int64_t rows,
                                 int64_t M, int64_t N, float* out, int64_t ldo, void* stream) {
  if (rows <= 0 || M <= 0 || N <= 0 || (M % 8) || (N % 8)) return set_error(-1, "gemm_wgrad: bad shape");
  const int BN = (N % 256 == 0) ? 256 : 128;
  GemmParams p;
  p.nseg = 1;
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(N);
  p.m_tiles = static_cast<int>((M + BM - 1) / BM);
  p.n_tiles = static_cast<int>((N + BN - 1) / BN);
  p.bias = nullptr;
  for (int s = 0; s < kMaxSeg; ++s) p.kblocks[s] = 0;
  p.kblocks[0] = static_cast<int>((rows + BK - 1) / BK);
  const int tiles = p.m_tiles * p.n_tiles;
  int splits = (2 * num_sms() + tiles - 1) / tiles;
  if (splits > p.kblocks[0]) splits = p.kblocks[0];
  if (splits < 1) splits = 1;
  p.kb_per_split = (p.kblocks[0] + splits - 1) / splits;
  p.k_splits = (p.kblocks[0] + p.kb_per_split - 1) / p.kb_per_split;
  int rc = make_tensor_map(&p.tma_a[0], dy, M, rows, ld_dy * 2, 64, 64, false);
  if (rc) return rc;
  rc = make_tensor_map(&p.tma_b[0], x, N, rows, ld_x * 2, 64, 64, false);
  if (rc) return rc;
  rc = make_tensor_map(&p.tma_out, out, N, M, ldo * 4, 32, BM, true);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  using C256 = Cfg<256>;
  using C128 = Cfg<128>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<256, EPI_WGRAD>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C256::kSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_tcgen05_kernel<128, EPI_WGRAD>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, C128::kSmemBytes);
    if (e != cudaSuccess) return set_error(-5, cudaGetErrorString(e));
    configured = true;
  }
  const int work = tiles * p.k_splits;
  const int grid = work < num_sms() ? work : num_sms();
  if (BN == 256)
    gemm_tcgen05_kernel<256, EPI_WGRAD><<<grid, 64 + epi_threads(EPI_WGRAD), C256::kSmemBytes, st>>>(p);
  else
    gemm_tcgen05_kernel<128, EPI_WGRAD><<<grid, 64 + epi_threads(EPI_WGRAD), C128::kSmemBytes, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(-6, cudaGetErrorString(e));
  return 0;
}
