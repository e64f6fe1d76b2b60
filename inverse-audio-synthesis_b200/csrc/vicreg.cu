// VICReg loss for sm_100a: invariance / variance / covariance terms of vicreg.VICReg.loss (vicreg.py:35-58) and
// off_diagonal (vicreg.py:73-76), over the gathered batch.
//
// Forward pipeline (all on the caller's stream, no host sync):
//   k_colsum        column sums of x and y (deterministic two-level), sum (x-y)^2 over the local rows
//   k_center_pack   subtract the column mean, split every element into tf32 hi + tf32 lo, and write the transposed
//                   operand X_c^T [D][B] as ready-made 128x32 shared-memory tile images (K-major, 128-byte swizzle)
//   k_gram_tc       tcgen05 Gram: per CTA one 128x128 tile of X_c^T X_c over one K-split; operands arrive by
//                   cp.async.bulk (TMA 1-D bulk copy) through a 3-stage mbarrier ring, 3 tf32 MMAs per k-step
//                   (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM), epilogue tcgen05.ld -> partial tile
//   k_cov_reduce    sum the K-split partials; off-diagonal squares (x2 for tiles above the diagonal), diagonal out
//   k_finalize      variance -> std hinge, covariance scale, the four scalars
// The only dense contraction of the hot path is the Gram, so it is the only tensor-core kernel (north star).
#include "ias_common.cuh"

#include <stdint.h>

namespace ias {
namespace {

constexpr int TILE = 128;        // Gram output tile (UMMA M = N = 128)
constexpr int KBLK = 32;         // batch rows per K-block: 32 tf32 = 128 B = one swizzle row
constexpr int TILE_FLOATS = TILE * KBLK;          // 4096 floats = 16 KiB per operand tile image
constexpr int TILE_BYTES = TILE_FLOATS * 4;
constexpr int STAGES = 3;
constexpr int COLSUM_ROWS = 128;  // rows per k_colsum CTA
constexpr int COV_ROWS = 4;       // Gram rows per k_cov_reduce CTA (128-column tile): 2 elements per thread
constexpr int PACK_KB = 1;        // K-blocks (32 batch rows) per k_center_pack CTA: B/32 x 2 CTAs keep every SM busy
constexpr int MAX_SPLITS = 64;

struct Plan {
  int B, D, DT, Dp, KB, ntiles, splits, P, MT, NV;
  size_t off_partial, off_varpart, off_repr, off_mean, off_packed, off_gram, off_covp, off_diag, off_stats, off_gfull, off_rowpack, off_gpack, total;
};

__host__ inline Plan make_plan(int B, int D) {
  Plan p;
  p.B = B;
  p.D = D;
  p.DT = (D + TILE - 1) / TILE;
  p.Dp = p.DT * TILE;
  p.KB = (B + KBLK - 1) / KBLK;
  p.ntiles = p.DT * (p.DT + 1) / 2;
  // enough K-splits to cover ~one wave of 148 SMs, at least 2 K-blocks per split
  int want = (148 + 2 * p.ntiles - 1) / (2 * p.ntiles);
  int max_by_k = (p.KB + 1) / 2;
  p.splits = want < 1 ? 1 : want;
  if (p.splits > max_by_k) p.splits = max_by_k < 1 ? 1 : max_by_k;
  if (p.splits > MAX_SPLITS) p.splits = MAX_SPLITS;
  p.P = (B + COLSUM_ROWS - 1) / COLSUM_ROWS;
  size_t o = 0;
  auto take = [&](size_t nfloats) {
    size_t r = o;
    o += (nfloats + 255) / 256 * 256;  // 1 KiB granularity keeps every region 1024-byte aligned
    return r;
  };
  p.off_partial = take((size_t)p.P * 2 * D);
  p.NV = (p.KB + PACK_KB - 1) / PACK_KB;  // k_center_pack CTAs per side = centred second-moment partials
  p.off_varpart = take((size_t)p.NV * 2 * D);
  p.off_repr = take((size_t)p.P);
  p.off_mean = take((size_t)2 * D);
  p.off_packed = take((size_t)2 * 2 * p.DT * p.KB * TILE_FLOATS);
  p.off_gram = take((size_t)2 * p.ntiles * p.splits * TILE * TILE);
  p.off_covp = take((size_t)2 * p.ntiles * (TILE / COV_ROWS));
  p.off_diag = take((size_t)2 * p.Dp);
  p.off_stats = take((size_t)8 * p.Dp);
  p.MT = (B + TILE - 1) / TILE;
  p.off_gfull = take((size_t)2 * p.Dp * p.Dp);
  p.off_rowpack = take((size_t)2 * 2 * p.MT * (p.Dp / KBLK) * TILE_FLOATS);
  p.off_gpack = take((size_t)2 * 2 * p.DT * (p.Dp / KBLK) * TILE_FLOATS);
  p.total = o;
  return p;
}

// ------------------------------------------------------------------------------------------------------------
// k_colsum: grid = P, block = 256.  partial[p][s][d] = sum over the CTA's rows; repr[p] = sum (x-y)^2 over local rows
// ------------------------------------------------------------------------------------------------------------
// Where the rows of the (global) batch live: rank q owns rows [q*rows_per, (q+1)*rows_per).  Single process: one
// entry.  Multi-GPU fused gather: the entries are peer pointers into every rank's symmetric-memory buffer, read over
// NVLink by the statistics kernel itself (no separate all-gather launch, no NCCL on the data path).
constexpr int MAX_PEERS = 16;
struct RowSrc {
  const float* x[MAX_PEERS];
  const float* y[MAX_PEERS];
  int rows_per;
};

// Scalar version: any D, any alignment.
__global__ void __launch_bounds__(256) k_colsum(RowSrc src, int B, int D, int local_row0, int local_rows,
                                                float* __restrict__ partial, float* __restrict__ repr,
                                                float* __restrict__ xg, float* __restrict__ yg) {
  __shared__ float s_red[8];
  const int r0 = blockIdx.x * COLSUM_ROWS;
  const int r1 = min(r0 + COLSUM_ROWS, B);
  float rp = 0.0f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float sx = 0.0f, sy = 0.0f;
    for (int r = r0; r < r1; ++r) {
      const int q = r / src.rows_per;
      const size_t lo = (size_t)(r - q * src.rows_per) * D + d;
      // plain (coherent) loads: the peers' buffers were written by other GPUs just before the barrier
      const float a = src.x[q][lo], c = src.y[q][lo];
      if (xg) {  // fused all-gather: keep a local copy of the gathered batch for the passes that follow
        xg[(size_t)r * D + d] = a;
        yg[(size_t)r * D + d] = c;
      }
      sx += a;
      sy += c;
      if (r >= local_row0 && r < local_row0 + local_rows) {
        const float e = a - c;
        rp = fmaf(e, e, rp);
      }
    }
    partial[((size_t)blockIdx.x * 2 + 0) * D + d] = sx;
    partial[((size_t)blockIdx.x * 2 + 1) * D + d] = sy;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rp += __shfl_xor_sync(0xffffffffu, rp, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = rp;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    repr[blockIdx.x] = t;
  }
}

// 128-bit version (D % 4 == 0, 16-byte aligned rows): the CTA's 128 rows are split over 4 row groups of 64 threads,
// each thread owns one column quad and keeps 8 independent 16-byte loads in flight per matrix, so the pass runs at
// memory latency / 8 per row instead of one round trip per row.  Same outputs as k_colsum (the per-column sums add
// the same values in a different order).
constexpr int CS_CG = 64;                 // column quads per pass
constexpr int CS_THREADS = 512;
constexpr int CS_RG = CS_THREADS / CS_CG;  // row groups: 8 x 8 rows in flight = half of a CTA's 128 rows per round trip
__global__ void __launch_bounds__(CS_THREADS) k_colsum_v4(RowSrc src, int B, int D, int local_row0, int local_rows,
                                                   float* __restrict__ partial, float* __restrict__ repr,
                                                   float* __restrict__ xg, float* __restrict__ yg) {
  __shared__ float4 s_part[2][CS_RG][CS_CG];
  __shared__ float s_red[CS_THREADS / 32];
  const int r0 = blockIdx.x * COLSUM_ROWS;
  const int r1 = min(r0 + COLSUM_ROWS, B);
  const int cg = threadIdx.x % CS_CG, rg = threadIdx.x / CS_CG;
  const int D4 = D / 4;
  // rows of one CTA normally belong to one rank (rows_per % 128 == 0): resolve the peer once
  const int q0 = r0 / src.rows_per;
  const bool one_peer = (r1 - 1) / src.rows_per == q0;
  float rp = 0.0f;
  for (int c0 = 0; c0 < D4; c0 += CS_CG) {
    const int c4 = c0 + cg;
    float4 sx = make_float4(0.f, 0.f, 0.f, 0.f), sy = sx;
    if (c4 < D4) {
      constexpr int U = 8;
      for (int rb = r0 + rg; rb < r1; rb += CS_RG * U) {
        float4 a[U], c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int r = rb + u * CS_RG;
          if (r < r1) {
            const int q = one_peer ? q0 : r / src.rows_per;
            const size_t lo = (size_t)(r - q * src.rows_per) * D4 + c4;
            // plain (coherent) loads: the peers' buffers were written by other GPUs just before the barrier
            a[u] = reinterpret_cast<const float4*>(src.x[q])[lo];
            c[u] = reinterpret_cast<const float4*>(src.y[q])[lo];
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int r = rb + u * CS_RG;
          if (r < r1) {
            if (xg) {  // fused all-gather: keep a local copy of the gathered batch for the passes that follow
              reinterpret_cast<float4*>(xg)[(size_t)r * D4 + c4] = a[u];
              reinterpret_cast<float4*>(yg)[(size_t)r * D4 + c4] = c[u];
            }
            sx.x += a[u].x; sx.y += a[u].y; sx.z += a[u].z; sx.w += a[u].w;
            sy.x += c[u].x; sy.y += c[u].y; sy.z += c[u].z; sy.w += c[u].w;
            if (r >= local_row0 && r < local_row0 + local_rows) {
              const float e0 = a[u].x - c[u].x, e1 = a[u].y - c[u].y, e2 = a[u].z - c[u].z, e3 = a[u].w - c[u].w;
              rp = fmaf(e0, e0, rp); rp = fmaf(e1, e1, rp); rp = fmaf(e2, e2, rp); rp = fmaf(e3, e3, rp);
            }
          }
        }
      }
    }
    s_part[0][rg][cg] = sx;
    s_part[1][rg][cg] = sy;
    __syncthreads();
    if (rg < 2 && c4 < D4) {  // row group 0 finishes x, row group 1 finishes y
      float4 t = s_part[rg][0][cg];
#pragma unroll
      for (int g = 1; g < CS_RG; ++g) {
        const float4 v = s_part[rg][g][cg];
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
      }
      reinterpret_cast<float4*>(partial + ((size_t)blockIdx.x * 2 + rg) * D)[c4] = t;
    }
    __syncthreads();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rp += __shfl_xor_sync(0xffffffffu, rp, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = rp;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < CS_THREADS / 32; ++w) t += s_red[w];
    repr[blockIdx.x] = t;
  }
}

// ------------------------------------------------------------------------------------------------------------
// k_center_pack: grid = (KB, 2), block = 256.  One K-block (32 batch rows) of one matrix -> DT tile images, hi and lo.
// Tile image (K-major, SWIZZLE_128B): element (row d, k) at byte (d/8)*1024 + (d%8)*128 + (((k/4) ^ (d%8))*16) + (k%4)*4.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}


__global__ void __launch_bounds__(256) k_center_pack(const float* __restrict__ x, const float* __restrict__ y, int B,
                                                     int D, int P, int DT, int KB, const float* __restrict__ partial,
                                                     float* __restrict__ mean_out, float* __restrict__ varpart,
                                                     float* __restrict__ packed) {
  const int pb = blockIdx.x, s = blockIdx.y;
  const float* src = s ? y : x;
  const float invB = 1.0f / (float)B;
  for (int d = threadIdx.x; d < DT * TILE; d += blockDim.x) {
    float mean = 0.0f;
    if (d < D) {
      // P column-sum partials (64 at B = 8192): eight loads in flight, summed in index order as before
      float t = 0.0f;
      int p = 0;
      for (; p + 8 <= P; p += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = partial[((size_t)(p + u) * 2 + s) * D + d];
#pragma unroll
        for (int u = 0; u < 8; ++u) t += v[u];
      }
      for (; p < P; ++p) t += partial[((size_t)p * 2 + s) * D + d];
      mean = t * invB;
      if (pb == 0) mean_out[s * D + d] = mean;
    }
    const int dt = d / TILE, dl = d % TILE;
    const int rowbase = (dl >> 3) * 256 + (dl & 7) * 32;  // in floats
    float sq4[PACK_KB];
#pragma unroll
    for (int q = 0; q < PACK_KB; ++q) {
      const int kb = pb * PACK_KB + q;
      sq4[q] = 0.0f;
      if (kb >= KB) continue;
      float* hi = packed + ((((size_t)s * 2 + 0) * DT + dt) * KB + kb) * TILE_FLOATS;
      float* lo = packed + ((((size_t)s * 2 + 1) * DT + dt) * KB + kb) * TILE_FLOATS;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 h, l;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = kb * KBLK + c * 4 + e;
          v[e] = (d < D && r < B) ? __ldg(src + (size_t)r * D + d) - mean : 0.0f;
          sq4[q] = fmaf(v[e], v[e], sq4[q]);
        }
        h.x = to_tf32(v[0]); h.y = to_tf32(v[1]); h.z = to_tf32(v[2]); h.w = to_tf32(v[3]);
        l.x = to_tf32(v[0] - h.x); l.y = to_tf32(v[1] - h.y); l.z = to_tf32(v[2] - h.z); l.w = to_tf32(v[3] - h.w);
        const int chunk = (c ^ (dl & 7)) * 4;
        *reinterpret_cast<float4*>(hi + rowbase + chunk) = h;
        *reinterpret_cast<float4*>(lo + rowbase + chunk) = l;
      }
    }
    // centred second moment of this CTA's 128 rows (fp32, two-level); the variance comes from these, not from the
    // Gram diagonal: the hinge 1 - std amplifies a relative error of the variance by ~1/(1 - std)
    if (d < D) {
      float sq = 0.0f;
#pragma unroll
      for (int q = 0; q < PACK_KB; ++q) sq += sq4[q];
      varpart[((size_t)pb * 2 + s) * D + d] = sq;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, bulk copy, tcgen05
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a lost arrival traps (error surfaces at the next CUDA call) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format, cute::UMMA::SmemDescriptor):
// start>>4 | LBO(ignored)=1 <<16 | SBO = 1024 B (8 rows x 128 B) >>4 <<32 | version 1 <<46 | layout SWIZZLE_128B (2) <<61
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// kind::tf32, fp32 accumulate, A and B K-major, M = N = 128 (cute::UMMA::InstrDescriptor bit layout)
constexpr uint32_t IDESC_TF32_128x128 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TILE >> 3) << 17) |
                                        ((uint32_t)(TILE >> 4) << 24);

struct GramSmem {
  float tile[STAGES][4][TILE_FLOATS];  // [stage][A_hi, A_lo, B_hi, B_lo], each 16 KiB, 1024-byte aligned
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t done;
  uint32_t tmem_base;
};

// One CTA-wide tcgen05 pass: D[128 x 128] (TMEM, fp32) = sum over nkb K-blocks of A_kb (128 x 32) * B_kb (128 x 32)^T
// with every operand split into tf32 hi + lo tile images (3 MMAs per k-step: hi*hi + hi*lo + lo*hi).
// Warp 0 lane 0 feeds a STAGES-deep ring of 16 KiB cp.async.bulk copies, warp 1 lane 0 issues the MMAs, warp 2 owns
// the TMEM allocation.  Returns the TMEM base address once the accumulator is complete and visible to all threads.
struct TileStream {
  const float* hi;    // first K-block of the hi tile images
  const float* lo;    // first K-block of the lo tile images
  size_t kb_stride;   // floats between consecutive K-blocks
};

__device__ __forceinline__ uint32_t tc_setup(GramSmem& sm) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&sm.full[i], 1);
      mbar_init(&sm.empty[i], 1);
    }
    mbar_init(&sm.done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {  // one warp allocates 128 TMEM columns (128 lanes x 128 fp32 accumulators)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                 "n"(TILE)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return sm.tmem_base;
}

__device__ __forceinline__ void tc_mainloop(GramSmem& sm, uint32_t tmem, TileStream A, TileStream Bm, bool same,
                                            int nkb) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    // ---- producer: one elected thread feeds the ring with 16 KiB bulk copies ----
    for (int i = 0; i < nkb; ++i) {
      const int st = i % STAGES;
      if (i >= STAGES) mbar_wait(&sm.empty[st], ((i / STAGES) - 1) & 1);
      mbar_expect_tx(&sm.full[st], same ? 2 * TILE_BYTES : 4 * TILE_BYTES);
      bulk_g2s(sm.tile[st][0], A.hi + (size_t)i * A.kb_stride, TILE_BYTES, &sm.full[st]);
      bulk_g2s(sm.tile[st][1], A.lo + (size_t)i * A.kb_stride, TILE_BYTES, &sm.full[st]);
      if (!same) {
        bulk_g2s(sm.tile[st][2], Bm.hi + (size_t)i * Bm.kb_stride, TILE_BYTES, &sm.full[st]);
        bulk_g2s(sm.tile[st][3], Bm.lo + (size_t)i * Bm.kb_stride, TILE_BYTES, &sm.full[st]);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ---- MMA issuer: a single thread drives the tensor core ----
    for (int i = 0; i < nkb; ++i) {
      const int st = i % STAGES;
      mbar_wait(&sm.full[st], (i / STAGES) & 1);
      tc_fence_after();
      const uint32_t ah = smem_u32(sm.tile[st][0]), al = smem_u32(sm.tile[st][1]);
      const uint32_t bh = same ? ah : smem_u32(sm.tile[st][2]);
      const uint32_t bl = same ? al : smem_u32(sm.tile[st][3]);
#pragma unroll
      for (int k = 0; k < KBLK / 8; ++k) {  // UMMA_K = 8 for tf32 = 32 bytes along the swizzle row
        const uint64_t dah = umma_desc_sw128(ah + 32 * k), dal = umma_desc_sw128(al + 32 * k);
        const uint64_t dbh = umma_desc_sw128(bh + 32 * k), dbl = umma_desc_sw128(bl + 32 * k);
        tc_mma_tf32(tmem, dah, dbh, IDESC_TF32_128x128, (i | k) != 0);
        tc_mma_tf32(tmem, dah, dbl, IDESC_TF32_128x128, 1);
        tc_mma_tf32(tmem, dal, dbh, IDESC_TF32_128x128, 1);
      }
      tc_commit(&sm.empty[st]);  // frees the smem slot once these MMAs have read it
    }
    tc_commit(&sm.done);  // accumulator complete
  }
  __syncwarp();
  mbar_wait(&sm.done, 0);
  tc_fence_after();
}

// 32 fp32 accumulators of this thread's TMEM lane (= output row), columns [c0, c0 + 32)
__device__ __forceinline__ void tc_load32(uint32_t tmem, int c0, uint32_t (&r)[32]) {
  const uint32_t taddr = tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16) + (uint32_t)c0;
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tc_teardown(uint32_t tmem) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TILE) : "memory");
  }
}

__device__ __forceinline__ GramSmem& aligned_smem(uint8_t* raw) {
  return *reinterpret_cast<GramSmem*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
}

// Forward Gram.  grid = 2 * ntiles * splits, block = 128.
__global__ void __launch_bounds__(128, 1) k_gram_tc(const float* __restrict__ packed, float* __restrict__ gram_partial,
                                                    int DT, int KB, int ntiles, int splits) {
  extern __shared__ uint8_t smem_raw[];
  GramSmem& sm = aligned_smem(smem_raw);
  int u = blockIdx.x;
  const int split = u % splits;
  u /= splits;
  const int t = u % ntiles;
  const int s = u / ntiles;
  int tm = 0, rem = t;  // upper-triangular tile index -> (tm, tn), tn >= tm
  while (rem >= DT - tm) {
    rem -= DT - tm;
    ++tm;
  }
  const int tn = tm + rem;
  const int kb0 = (int)((long long)KB * split / splits);
  const int kb1 = (int)((long long)KB * (split + 1) / splits);
  const int nkb = kb1 - kb0;
  const uint32_t tmem = tc_setup(sm);

  const size_t part_stride = (size_t)DT * KB * TILE_FLOATS;  // hi -> lo
  TileStream A, Bm;
  A.hi = packed + ((((size_t)s * 2 + 0) * DT + tm) * KB + kb0) * TILE_FLOATS;
  A.lo = A.hi + part_stride;
  A.kb_stride = TILE_FLOATS;
  Bm.hi = packed + ((((size_t)s * 2 + 0) * DT + tn) * KB + kb0) * TILE_FLOATS;
  Bm.lo = Bm.hi + part_stride;
  Bm.kb_stride = TILE_FLOATS;
  tc_mainloop(sm, tmem, A, Bm, tm == tn, nkb);

  // ---- epilogue: TMEM -> registers -> partial tile (row = TMEM lane = thread) ----
  float* dst = gram_partial + ((((size_t)s * ntiles + t) * splits + split) * TILE + threadIdx.x) * TILE;
#pragma unroll
  for (int c0 = 0; c0 < TILE; c0 += 32) {
    uint32_t r[32];
    tc_load32(tmem, c0, r);
#pragma unroll
    for (int c = 0; c < 32; c += 4)
      *reinterpret_cast<float4*>(dst + c0 + c) =
          nkb > 0 ? make_float4(__uint_as_float(r[c]), __uint_as_float(r[c + 1]), __uint_as_float(r[c + 2]),
                                __uint_as_float(r[c + 3]))
                  : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  tc_teardown(tmem);
}

// ------------------------------------------------------------------------------------------------------------
// Backward: d loss / d x, d y.
//   gx[r][d] = kappa * sum_j xc[r][j] * offG_x[j][d]  +  xc[r][d] * s_x[d]  +  c_r * (x[r][d] - y[r][d])
//   gy likewise with -c_r;  offG = Gram with zero diagonal, s[d] = c_std * (std < 1 ? -1 / (2 D std (B-1)) : 0),
//   kappa = c_cov * 4 / (embeddim (n-1)^2),  c_r = c_repr * 2 / (B_local D) on the local rows, 0 elsewhere.
// The contraction is the same tcgen05 pass with A = 128 batch rows of X_c (K-major as stored) and B = 128 rows of offG.
// ------------------------------------------------------------------------------------------------------------
// k_pack_rows: grid = (MT, 2), block = 256: centred rows -> tile images [side][part][mt][kbd]
__global__ void __launch_bounds__(256) k_pack_rows(const float* __restrict__ x, const float* __restrict__ y, int B, int D,
                                                   int Dp, const float* __restrict__ mean, float* __restrict__ rowpack,
                                                   int MT, int mt0) {
  const int mt = blockIdx.x, s = blockIdx.y;  // mt counts from the first packed row tile mt0
  const float* src = s ? y : x;
  const int KBD = Dp / KBLK;
  const int f4_per_row = Dp / 4;
  for (int f = threadIdx.x; f < TILE * f4_per_row; f += blockDim.x) {
    const int rl = f / f4_per_row, c4 = f - rl * f4_per_row;
    const int d = 4 * c4, row = (mt0 + mt) * TILE + rl;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      v[e] = (row < B && d + e < D) ? __ldg(src + (size_t)row * D + d + e) - mean[s * D + d + e] : 0.0f;
    float4 h, l;
    h.x = to_tf32(v[0]); h.y = to_tf32(v[1]); h.z = to_tf32(v[2]); h.w = to_tf32(v[3]);
    l.x = to_tf32(v[0] - h.x); l.y = to_tf32(v[1] - h.y); l.z = to_tf32(v[2] - h.z); l.w = to_tf32(v[3] - h.w);
    const int kb = d / KBLK, c = (d % KBLK) / 4;
    const size_t off = (size_t)(rl >> 3) * 256 + (rl & 7) * 32 + ((c ^ (rl & 7)) * 4);
    float* hi = rowpack + ((((size_t)s * 2 + 0) * MT + mt) * KBD + kb) * TILE_FLOATS;
    float* lo = rowpack + ((((size_t)s * 2 + 1) * MT + mt) * KBD + kb) * TILE_FLOATS;
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
  }
}

// k_pack_offg: grid = (DT, 2), block = 256: Gram with zero diagonal -> tile images [side][part][nt][kbd]
__global__ void __launch_bounds__(256) k_pack_offg(const float* __restrict__ gram_full, int Dp,
                                                   float* __restrict__ gpack, int DT) {
  const int nt = blockIdx.x, s = blockIdx.y;
  const int KBD = Dp / KBLK;
  const int f4_per_row = Dp / 4;
  for (int f = threadIdx.x; f < TILE * f4_per_row; f += blockDim.x) {
    const int rl = f / f4_per_row, c4 = f - rl * f4_per_row;
    const int d = 4 * c4, row = nt * TILE + rl;
    float4 g = *reinterpret_cast<const float4*>(gram_full + ((size_t)s * Dp + row) * Dp + d);
    float v[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (d + e == row) v[e] = 0.0f;
    float4 h, l;
    h.x = to_tf32(v[0]); h.y = to_tf32(v[1]); h.z = to_tf32(v[2]); h.w = to_tf32(v[3]);
    l.x = to_tf32(v[0] - h.x); l.y = to_tf32(v[1] - h.y); l.z = to_tf32(v[2] - h.z); l.w = to_tf32(v[3] - h.w);
    const int kb = d / KBLK, c = (d % KBLK) / 4;
    const size_t off = (size_t)(rl >> 3) * 256 + (rl & 7) * 32 + ((c ^ (rl & 7)) * 4);
    float* hi = gpack + ((((size_t)s * 2 + 0) * DT + nt) * KBD + kb) * TILE_FLOATS;
    float* lo = gpack + ((((size_t)s * 2 + 1) * DT + nt) * KBD + kb) * TILE_FLOATS;
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
  }
}

struct BwdArgs {
  const float* x;
  const float* y;
  const float* mean;     // [2][D]
  const float* stdv;     // [2][Dp]
  const float* gout4;    // device: upstream grads of {loss, repr, std, cov}
  const float* rowpack;
  const float* gpack;
  float* gx;
  float* gy;
  int B, D, Dp, DT, MT, local_row0, B_local, cfgB, embeddim;
  int Bstat;      // rows behind the batch statistics (== B, or the global batch when x/y hold one rank's rows only)
  int mt0;        // first row tile computed (MT tiles from there)
  int out_row0;   // gx/gy hold rows [out_row0, out_row0 + out_rows) of the batch
  int out_rows;
  float gscale;   // multiplies the std/cov gradient (world size in the fused-gather path)
  float sim, stdc, covc;
};

// grid = 2 * MT * DT, block = 128
__global__ void __launch_bounds__(128, 1) k_bwd_tc(BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  GramSmem& sm = aligned_smem(smem_raw);
  int u = blockIdx.x;
  const int nt = u % a.DT;
  u /= a.DT;
  const int mt = u % a.MT;
  const int s = u / a.MT;
  const int KBD = a.Dp / KBLK;
  const uint32_t tmem = tc_setup(sm);
  TileStream A, Bm;
  A.hi = a.rowpack + ((((size_t)s * 2 + 0) * a.MT + mt) * KBD) * TILE_FLOATS;
  A.lo = a.rowpack + ((((size_t)s * 2 + 1) * a.MT + mt) * KBD) * TILE_FLOATS;
  A.kb_stride = TILE_FLOATS;
  Bm.hi = a.gpack + ((((size_t)s * 2 + 0) * a.DT + nt) * KBD) * TILE_FLOATS;
  Bm.lo = a.gpack + ((((size_t)s * 2 + 1) * a.DT + nt) * KBD) * TILE_FLOATS;
  Bm.kb_stride = TILE_FLOATS;
  tc_mainloop(sm, tmem, A, Bm, false, KBD);

  const float g0 = a.gout4[0], g1 = a.gout4[1], g2 = a.gout4[2], g3 = a.gout4[3];
  const float c_std = g0 * a.stdc + g2;
  const float n1 = (float)(a.cfgB - 1);
  const float kappa = (g0 * a.covc + g3) * 4.0f / ((float)a.embeddim * n1 * n1);
  const int row = (a.mt0 + mt) * TILE + threadIdx.x;
  const bool local = row >= a.local_row0 && row < a.local_row0 + a.B_local;
  float c_r = local ? (g0 * a.sim + g1) * 2.0f / ((float)a.B_local * (float)a.D) : 0.0f;
  if (s) c_r = -c_r;
  const float s_scale = -a.gscale * c_std / (2.0f * (float)a.D * (float)(a.Bstat - 1));
  const float kap = kappa * a.gscale;
  float* dst = s ? a.gy : a.gx;
#pragma unroll
  for (int c0 = 0; c0 < TILE; c0 += 32) {
    uint32_t r[32];
    tc_load32(tmem, c0, r);
    if (row < a.B && row >= a.out_row0 && row < a.out_row0 + a.out_rows) {
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int d = nt * TILE + c0 + c;
        if (d < a.D) {
          const float xv = __ldg(a.x + (size_t)row * a.D + d), yv = __ldg(a.y + (size_t)row * a.D + d);
          const float ov = s ? yv : xv;
          const float sd = a.stdv[s * a.Dp + d];
          const float sg = sd < 1.0f ? s_scale / sd : 0.0f;
          dst[(size_t)(row - a.out_row0) * a.D + d] =
              kap * __uint_as_float(r[c]) + (ov - a.mean[s * a.D + d]) * sg + c_r * (xv - yv);
        }
      }
    }
  }
  tc_teardown(tmem);
}

// Plain CUDA-core Gram of the same packed operands' source (test hook): gram[D][D] = xc^T xc in fp32.
__global__ void __launch_bounds__(256) k_gram_simt(const float* __restrict__ x, const float* __restrict__ mean, int B,
                                                   int D, float* __restrict__ gram) {
  __shared__ float sa[16][17], sb[16][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i = blockIdx.y * 16 + ty, j = blockIdx.x * 16 + tx;
  float acc = 0.0f;
  for (int r0 = 0; r0 < B; r0 += 16) {
    const int r = r0 + ty;
    const int ca = blockIdx.y * 16 + tx, cb = blockIdx.x * 16 + tx;
    sa[ty][tx] = (r < B && ca < D) ? x[(size_t)r * D + ca] - mean[ca] : 0.0f;
    sb[ty][tx] = (r < B && cb < D) ? x[(size_t)r * D + cb] - mean[cb] : 0.0f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fmaf(sa[k][ty], sb[k][tx], acc);
    __syncthreads();
  }
  if (i < D && j < D) gram[(size_t)i * D + j] = acc;
}

// ------------------------------------------------------------------------------------------------------------
// k_cov_reduce: grid = 2 * ntiles * (TILE/16), block = 256: 16 rows x 128 cols of one tile.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cov_reduce(const float* __restrict__ gram_partial, int DT, int ntiles,
                                                    int splits, int Dp, float* __restrict__ covp,
                                                    float* __restrict__ diag_out, float* __restrict__ gram_full) {
  __shared__ float s_red[8];
  int u = blockIdx.x;
  const int rb = u % (TILE / COV_ROWS);
  u /= (TILE / COV_ROWS);
  const int t = u % ntiles;
  const int s = u / ntiles;
  int tm = 0, rem = t;
  while (rem >= DT - tm) {
    rem -= DT - tm;
    ++tm;
  }
  const int tn = tm + rem;
  const float* base = gram_partial + (((size_t)s * ntiles + t) * splits) * TILE * TILE;
  float sq = 0.0f;
  for (int e = threadIdx.x; e < COV_ROWS * TILE; e += 256) {
    const int r = rb * COV_ROWS + e / TILE, c = e % TILE;
    float g = 0.0f;
    for (int k = 0; k < splits; ++k) g += base[(size_t)k * TILE * TILE + (size_t)r * TILE + c];
    const int gi = tm * TILE + r, gj = tn * TILE + c;
    if (gi == gj)
      diag_out[s * Dp + gi] = g;
    else
      sq = fmaf(g, g, sq);
    if (gram_full) {
      gram_full[((size_t)s * Dp + gi) * Dp + gj] = g;
      if (tm != tn) gram_full[((size_t)s * Dp + gj) * Dp + gi] = g;  // tiles below the diagonal are never computed
    }
  }
  if (tm != tn) sq *= 2.0f;  // the mirrored tile below the diagonal
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tsum = 0.0f;
    for (int w = 0; w < 8; ++w) tsum += s_red[w];
    covp[blockIdx.x] = tsum;
  }
}

// ------------------------------------------------------------------------------------------------------------
// k_finalize: one CTA.  stats[s][0][d] = std, used by the backward pass.
// ------------------------------------------------------------------------------------------------------------
struct FinalArgs {
  const float* repr;   // [P]
  const float* covp;   // [2 * ntiles * 8]
  const float* varpart;  // [NV][2][D] centred second-moment partials
  float* stats;        // [2][Dp] std
  float* out4;
  int P, NV, ncovp_per_side, D, Dp, B, B_local, cfgB, embeddim;
  float sim, stdc, covc;
  int* epoch;          // statistics exchange: step counter advanced once the step's results are final (else null)
};

constexpr int FIN_THREADS = 1024;
__global__ void __launch_bounds__(FIN_THREADS) k_finalize(FinalArgs a) {
  // variance: NV partials (one per 32 batch rows) per dimension and side; four threads share a dimension so that the
  // 2*NV*D loads of a large gathered batch (NV = 256 at B = 8192) are spread over the whole CTA
  __shared__ float s_var[4][256];
  const int t = threadIdx.x, col = t & 255, part = t >> 8;
  float hinge[2] = {0.0f, 0.0f};
  for (int d0 = 0; d0 < a.D; d0 += 256) {
    const int d = d0 + col;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      float acc0 = 0.0f, acc1 = 0.0f;
      if (d < a.D) {
        constexpr int UV = 8;  // partials in flight per thread: the loop is a chain of L2 round trips otherwise
        int p = part;
        for (; p + 4 * (UV - 1) < a.NV; p += 4 * UV) {
          float v[UV];
#pragma unroll
          for (int u = 0; u < UV; ++u) v[u] = a.varpart[((size_t)(p + 4 * u) * 2 + s) * a.D + d];
#pragma unroll
          for (int u = 0; u < UV; u += 2) {
            acc0 += v[u];
            acc1 += v[u + 1];
          }
        }
        for (; p < a.NV; p += 4) acc0 += a.varpart[((size_t)p * 2 + s) * a.D + d];
      }
      s_var[part][col] = acc0 + acc1;
      __syncthreads();
      if (part == 0 && d < a.D) {
        const float ss = (s_var[0][col] + s_var[1][col]) + (s_var[2][col] + s_var[3][col]);
        const float var = ss / (float)(a.B - 1);  // NaN for B == 1, as torch.var
        const float sd = sqrtf(var + 0.0001f);
        a.stats[s * a.Dp + d] = sd;
        hinge[s] += fmaxf(1.0f - sd, 0.0f);
      }
      __syncthreads();
    }
  }
  // cov_x and cov_y partials are summed separately, then divided by embeddim, then added (vicreg.py:49-51)
  float covx = 0.0f, covy = 0.0f;
  for (int i = t; i < a.ncovp_per_side; i += FIN_THREADS) {
    covx += a.covp[i];
    covy += a.covp[a.ncovp_per_side + i];
  }
  float rp = 0.0f;
  for (int i = t; i < a.P; i += FIN_THREADS) rp += a.repr[i];
  float v[5] = {hinge[0], hinge[1], covx, covy, rp};
  __shared__ float s_all[5][FIN_THREADS / 32];
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    float tt = v[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tt += __shfl_xor_sync(0xffffffffu, tt, o);
    if ((t & 31) == 0) s_all[q][t >> 5] = tt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot[5];
    for (int q = 0; q < 5; ++q) {
      float tt = 0.0f;
      for (int w = 0; w < FIN_THREADS / 32; ++w) tt += s_all[q][w];
      tot[q] = tt;
    }
    const float repr_loss = tot[4] / ((float)a.B_local * (float)a.D);
    const float std_loss = (tot[0] / (float)a.D) / 2.0f + (tot[1] / (float)a.D) / 2.0f;
    const float n1 = (float)(a.cfgB - 1);
    const float cov_loss = (tot[2] / (n1 * n1)) / (float)a.embeddim + (tot[3] / (n1 * n1)) / (float)a.embeddim;
    a.out4[0] = a.sim * repr_loss + a.stdc * std_loss + a.covc * cov_loss;
    a.out4[1] = repr_loss;
    a.out4[2] = std_loss;
    a.out4[3] = cov_loss;
    if (a.epoch) *a.epoch += 1;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Statistics exchange (multi-GPU): instead of gathering the [B_local, D] embeddings and recomputing the
// global statistics on every rank, each rank reduces its own rows with the single-GPU kernels above (local mean,
// locally-centred second moments and Gram) and sends that summary -- a `packet` of 4*Dp + 2*ntiles*128*128 floats,
// 0.39 MB at D = 256 -- to every peer.  The global statistics follow from the exact pooled formulas
//     mu = sum_q B_q mu_q / B,    G = sum_q [ G_q + B_q (mu_q - mu)(mu_q - mu)^T ],    m2 likewise on the diagonal,
// which add a small between-rank term to well-conditioned within-rank sums (no X^T X - B mu mu^T cancellation).
//
// Transport: one peer-mapped buffer per rank (e.g. torch symmetric memory) laid out as
//     [ header: int flags[MAX_PEERS] | int epoch @32 | int done @33 ]  (1 KiB)   [ inbox[world][2][packet] ]
// k_stats_publish (rank r, step e) stores its packet into inbox[r][e & 1] of EVERY rank with plain 16-byte stores over
// NVLink, then -- last CTA done, after a system-scope fence -- writes e into flags[r] of every rank.
// k_stats_combine waits until flags[q] >= e for every q (bounded spin, trap on timeout), then reads only its own
// memory.  No separate barrier or collective launch; double-buffered by the parity of e, so a rank can only overwrite
// inbox[r][e & 1] at step e + 2, which it reaches after its own combine of step e + 1 saw every peer's flag e + 1,
// i.e. after every peer finished reading step e (stream order).  k_finalize advances the epoch.
// ------------------------------------------------------------------------------------------------------------
constexpr int XCHG_HDR_FLOATS = 256;
constexpr int XCHG_EPOCH = 32, XCHG_DONE = 33;

__host__ __device__ inline size_t packet_floats(int Dp, int ntiles) {
  return ((size_t)4 * Dp + (size_t)2 * ntiles * TILE * TILE + 255) / 256 * 256;
}

struct PeerTable {
  float* base[MAX_PEERS];
};

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// block = 256 = 64 float4 items of the packet x 4 cooperating lanes: lane j of an item adds up the K-split partials
// (or variance partials) j, j + 4, ... with all its loads in flight at once, two shuffle steps finish the sum in a fixed
// order, and lane j stores the result to peers j, j + 4, ...  grid = ceil(items / 64).
constexpr int PUB_PARTS = 4;
__global__ void __launch_bounds__(256) k_stats_publish(const float* __restrict__ gram_partial,
                                                       const float* __restrict__ varpart,
                                                       const float* __restrict__ mean, int D, int Dp, int ntiles,
                                                       int splits, int NV, PeerTable peers, int world, int rank) {
  float* own = peers.base[rank];
  const int epoch = *reinterpret_cast<const int*>(own + XCHG_EPOCH);  // advanced by k_finalize after this step
  const size_t pf = packet_floats(Dp, ntiles);
  const size_t slot = XCHG_HDR_FLOATS + ((size_t)rank * 2 + (epoch & 1)) * pf;
  const int n4 = (int)(pf / 4);
  const int part = threadIdx.x & (PUB_PARTS - 1);
  const int i4 = blockIdx.x * (256 / PUB_PARTS) + (threadIdx.x >> 2);
  const int f = i4 * 4;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i4 < n4) {
    if (f < 2 * Dp) {  // local column means, [2][Dp]
      const int s = f / Dp, d = f - s * Dp;
      if (part == 0) {
        float t[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) t[e] = (d + e < D) ? mean[s * D + d + e] : 0.0f;
        v = make_float4(t[0], t[1], t[2], t[3]);
      }
    } else if (f < 4 * Dp) {  // locally-centred second moments, [2][Dp]: sum of the NV k_center_pack partials
      const int g = f - 2 * Dp, s = g / Dp, d = g - s * Dp;
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      for (int p = part; p < NV; p += PUB_PARTS)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (d + e < D) t[e] += varpart[((size_t)p * 2 + s) * D + d + e];
      v = make_float4(t[0], t[1], t[2], t[3]);
    } else if ((size_t)f < (size_t)4 * Dp + (size_t)2 * ntiles * TILE * TILE) {  // Gram tiles, K-split partials summed
      const size_t g = (size_t)f - 4 * Dp;
      const size_t st = g / (TILE * TILE), e = g - st * (TILE * TILE);  // st = s * ntiles + t
      const float4* src = reinterpret_cast<const float4*>(gram_partial + st * splits * TILE * TILE + e);
      constexpr int U = 4;  // up to 16 splits: every load of a lane is in flight at once
      int k = part;
      for (; k + (U - 1) * PUB_PARTS < splits; k += U * PUB_PARTS) {
        float4 a[U];
#pragma unroll
        for (int u = 0; u < U; ++u) a[u] = src[(size_t)(k + u * PUB_PARTS) * (TILE * TILE / 4)];
#pragma unroll
        for (int u = 0; u < U; ++u) { v.x += a[u].x; v.y += a[u].y; v.z += a[u].z; v.w += a[u].w; }
      }
      for (; k < splits; k += PUB_PARTS) {
        const float4 a = src[(size_t)k * (TILE * TILE / 4)];
        v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
      }
    }
  }
#pragma unroll
  for (int o = 1; o < PUB_PARTS; o <<= 1) {  // lanes of an item are adjacent: fixed-order tree, every lane gets the sum
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
    v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
  }
  if (i4 < n4)
    for (int q = part; q < world; q += PUB_PARTS) *reinterpret_cast<float4*>(peers.base[q] + slot + f) = v;
  // last CTA done: every CTA's packet stores are ordered before its arrival, the last arrival publishes the flag.
  // ONE system-scope fence per CTA, by the thread that arrives: the barrier orders the other threads' stores before it
  // (cumulativity), and a fence per thread -- 99 k MEMBAR.SYS per launch -- was most of this kernel's time at N = 8.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    int* done = reinterpret_cast<int*>(own + XCHG_DONE);
    if (atomicAdd(done, 1) == (int)gridDim.x - 1) {
      *done = 0;
      __threadfence_system();
      for (int q = 0; q < world; ++q) st_release_sys(reinterpret_cast<int*>(peers.base[q]) + rank, epoch);
    }
  }
}

// grid = 2 * ntiles * (TILE / COV_ROWS), block = 256: COV_ROWS rows x 128 cols of one Gram tile (as k_cov_reduce).
__global__ void __launch_bounds__(256) k_stats_combine(const float* __restrict__ own, int world, int rows_per_rank,
                                                       int D, int Dp, int DT, int ntiles, float* __restrict__ covp,
                                                       float* __restrict__ gram_full, float* __restrict__ mean_out,
                                                       float* __restrict__ m2_out) {
  __shared__ float s_red[8];
  extern __shared__ float s_dev[];  // [world][TILE] column deviations, then [world][COV_ROWS] row deviations
  float (*s_dcol)[TILE] = reinterpret_cast<float (*)[TILE]>(s_dev);
  float (*s_drow)[COV_ROWS] = reinterpret_cast<float (*)[COV_ROWS]>(s_dev + world * TILE);
  const int epoch = *reinterpret_cast<const int*>(own + XCHG_EPOCH);
  if ((int)threadIdx.x < world) {  // wait for every rank's packet of this step (bounded: a lost peer traps, no hang)
    const int* flag = reinterpret_cast<const int*>(own) + threadIdx.x;
    unsigned spin = 0;
    unsigned long long t0 = 0;
    while (ld_acquire_sys(flag) < epoch) {
      if ((++spin & 1023u) == 0) {  // ranks may be seconds apart on their first step (module loading): 30 s budget
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 30000000000ull) __trap();
      }
      __nanosleep(100);
    }
  }
  __syncthreads();
  int u = blockIdx.x;
  const int rb = u % (TILE / COV_ROWS);
  u /= (TILE / COV_ROWS);
  const int t = u % ntiles;
  const int s = u / ntiles;
  int tm = 0, rem = t;
  while (rem >= DT - tm) {
    rem -= DT - tm;
    ++tm;
  }
  const int tn = tm + rem;
  const size_t pf = packet_floats(Dp, ntiles);
  const float* inbox = own + XCHG_HDR_FLOATS + (size_t)(epoch & 1) * pf;  // + q * 2 * pf for rank q
  const float Bl = (float)rows_per_rank, invW = 1.0f / (float)world;
  // between-rank deviations of the column means for this block's 128 columns and COV_ROWS rows.  Peers wrote the
  // inbox over NVLink: read it through L2 (ld.cg), never through this SM's L1.
  if (threadIdx.x < TILE + COV_ROWS) {
    const int d = threadIdx.x < TILE ? tn * TILE + threadIdx.x : tm * TILE + rb * COV_ROWS + (threadIdx.x - TILE);
    float m[MAX_PEERS], mu = 0.0f;
    for (int q = 0; q < world; ++q) {
      m[q] = __ldcg(inbox + (size_t)q * 2 * pf + (size_t)s * Dp + d);
      mu += m[q];
    }
    mu *= invW;
    for (int q = 0; q < world; ++q) {
      if (threadIdx.x < TILE) s_dcol[q][threadIdx.x] = m[q] - mu;
      else s_drow[q][threadIdx.x - TILE] = m[q] - mu;
    }
    if (threadIdx.x < TILE && tm == tn && rb == 0 && d < D) {  // one block per (side, diagonal tile) also emits these
      float m2 = 0.0f;
      for (int q = 0; q < world; ++q) {
        const float dq = m[q] - mu;
        m2 += __ldcg(inbox + (size_t)q * 2 * pf + 2 * Dp + (size_t)s * Dp + d) + Bl * dq * dq;
      }
      mean_out[s * D + d] = mu;
      m2_out[s * D + d] = m2;
    }
  }
  __syncthreads();
  float sq = 0.0f;
  for (int e = threadIdx.x; e < COV_ROWS * TILE; e += 256) {
    const int rl = e / TILE, c = e % TILE;
    const int r = rb * COV_ROWS + rl;
    const size_t off = (size_t)4 * Dp + ((size_t)s * ntiles + t) * (TILE * TILE) + (size_t)r * TILE + c;
    float g = 0.0f;
    for (int q = 0; q < world; ++q) g += __ldcg(inbox + (size_t)q * 2 * pf + off) + Bl * s_drow[q][rl] * s_dcol[q][c];
    const int gi = tm * TILE + r, gj = tn * TILE + c;
    if (gi != gj) sq = fmaf(g, g, sq);
    gram_full[((size_t)s * Dp + gi) * Dp + gj] = g;
    if (tm != tn) gram_full[((size_t)s * Dp + gj) * Dp + gi] = g;
  }
  if (tm != tn) sq *= 2.0f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tsum = 0.0f;
    for (int w = 0; w < 8; ++w) tsum += s_red[w];
    covp[blockIdx.x] = tsum;
  }
}

int check_common(const float* x, int B, int D, void* ws, size_t ws_bytes, const char* who) {
  IAS_REQUIRE(B > 0 && D > 0, IAS_ERR_INVALID, "%s: B=%d D=%d", who, B, D);
  IAS_REQUIRE(x != nullptr, IAS_ERR_INVALID, "%s: NULL input", who);
  IAS_REQUIRE(D <= 16384, IAS_ERR_UNSUPPORTED, "%s: D=%d > 16384", who, D);
  const Plan p = make_plan(B, D);
  IAS_REQUIRE(ws && ws_bytes >= p.total * sizeof(float), IAS_ERR_WORKSPACE, "%s: workspace %zu < %zu bytes", who,
              ws_bytes, p.total * sizeof(float));
  IAS_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 1023u) == 0, IAS_ERR_INVALID, "%s: workspace must be 1024-byte aligned",
              who);
  return IAS_OK;
}

int launch_gram(const Plan& p, float* w, cudaStream_t st) {
  static unsigned long long attr_devs = 0;
  const int smem = (int)sizeof(GramSmem) + 1024;
  if (ias_first_use_on_device(attr_devs)) {
    IAS_CUDA(cudaFuncSetAttribute(k_gram_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  {
    ProfScope prof_(K_VICREG_GRAM_TC, st);
    k_gram_tc<<<2 * p.ntiles * p.splits, 128, smem, st>>>(w + p.off_packed, w + p.off_gram, p.DT, p.KB, p.ntiles,
                                                          p.splits);
  }
  IAS_LAUNCH_CHECK("k_gram_tc");
  return IAS_OK;
}

RowSrc single_src(const float* x, const float* y, int B) {
  RowSrc s;
  for (int i = 0; i < MAX_PEERS; ++i) s.x[i] = s.y[i] = nullptr;
  s.x[0] = x;
  s.y[0] = y;
  s.rows_per = B;
  return s;
}

// x, y: where the centring/packing passes read the batch (the gathered copy when `src` is a peer table)
int run_stats_and_gram(const RowSrc& src, const float* x, const float* y, float* xg, float* yg, const Plan& p,
                       int local_row0, int B_local, float* w, float* gram_full, cudaStream_t st) {
  {
    bool v4 = (p.D % 4 == 0) && ias_aligned16(xg) && ias_aligned16(yg);
    for (int q = 0; q < MAX_PEERS; ++q) v4 = v4 && ias_aligned16(src.x[q]) && ias_aligned16(src.y[q]);
    ProfScope prof_(K_VICREG_COLSUM, st);
    if (v4)
      k_colsum_v4<<<p.P, CS_THREADS, 0, st>>>(src, p.B, p.D, local_row0, B_local, w + p.off_partial, w + p.off_repr, xg, yg);
    else
      k_colsum<<<p.P, 256, 0, st>>>(src, p.B, p.D, local_row0, B_local, w + p.off_partial, w + p.off_repr, xg, yg);
  }
  IAS_LAUNCH_CHECK("k_colsum");
  {
    ProfScope prof_(K_VICREG_PACK, st);
    k_center_pack<<<dim3(p.NV, 2), 256, 0, st>>>(x, y, p.B, p.D, p.P, p.DT, p.KB, w + p.off_partial, w + p.off_mean,
                                                w + p.off_varpart, w + p.off_packed);
  }
  IAS_LAUNCH_CHECK("k_center_pack");
  int rc = launch_gram(p, w, st);
  if (rc) return rc;
  {
    ProfScope prof_(K_VICREG_COV_REDUCE, st);
    k_cov_reduce<<<2 * p.ntiles * (TILE / COV_ROWS), 256, 0, st>>>(w + p.off_gram, p.DT, p.ntiles, p.splits, p.Dp,
                                                             w + p.off_covp, w + p.off_diag, gram_full);
  }
  IAS_LAUNCH_CHECK("k_cov_reduce");
  return IAS_OK;
}

}  // namespace
}  // namespace ias

using namespace ias;

extern "C" size_t ias_vicreg_workspace_bytes(int B, int D) {
  if (B <= 0 || D <= 0) return 0;
  return make_plan(B, D).total * sizeof(float);
}

extern "C" int ias_vicreg_loss(const float* x, const float* y, int B, int local_row0, int B_local, int cfg_batch_size,
                               int D, int embeddim, float sim_coeff, float std_coeff, float cov_coeff, float* out4,
                               void* workspace, size_t workspace_bytes, ias_stream_t stream) {
  int rc = check_common(x, B, D, workspace, workspace_bytes, "ias_vicreg_loss");
  if (rc) return rc;
  IAS_REQUIRE(y && out4, IAS_ERR_INVALID, "ias_vicreg_loss: NULL pointer");
  IAS_REQUIRE(local_row0 >= 0 && B_local > 0 && local_row0 + B_local <= B, IAS_ERR_INVALID,
              "ias_vicreg_loss: local rows [%d,%d) outside [0,%d)", local_row0, local_row0 + B_local, B);
  IAS_REQUIRE(cfg_batch_size != 1 && embeddim > 0, IAS_ERR_INVALID, "ias_vicreg_loss: cfg_batch_size=%d embeddim=%d",
              cfg_batch_size, embeddim);
  const Plan p = make_plan(B, D);
  float* w = reinterpret_cast<float*>(workspace);
  cudaStream_t st = as_stream(stream);
  rc = run_stats_and_gram(single_src(x, y, B), x, y, nullptr, nullptr, p, local_row0, B_local, w, w + p.off_gfull, st);
  if (rc) return rc;
  FinalArgs a;
  a.repr = w + p.off_repr;
  a.covp = w + p.off_covp;
  a.varpart = w + p.off_varpart;
  a.stats = w + p.off_stats;
  a.out4 = out4;
  a.P = p.P;
  a.NV = p.NV;
  a.ncovp_per_side = p.ntiles * (TILE / COV_ROWS);
  a.D = D; a.Dp = p.Dp; a.B = B; a.B_local = B_local; a.cfgB = cfg_batch_size; a.embeddim = embeddim;
  a.sim = sim_coeff; a.stdc = std_coeff; a.covc = cov_coeff;
  a.epoch = nullptr;
  {
    ProfScope prof_(K_VICREG_FINALIZE, st);
    k_finalize<<<1, FIN_THREADS, 0, st>>>(a);
  }
  IAS_LAUNCH_CHECK("k_finalize");
  return IAS_OK;
}

extern "C" int ias_vicreg_gram_tc(const float* x, int B, int D, float* gram, void* workspace, size_t workspace_bytes,
                                  ias_stream_t stream) {
  int rc = check_common(x, B, D, workspace, workspace_bytes, "ias_vicreg_gram_tc");
  if (rc) return rc;
  IAS_REQUIRE(gram, IAS_ERR_INVALID, "ias_vicreg_gram_tc: NULL output");
  IAS_REQUIRE(D % TILE == 0, IAS_ERR_UNSUPPORTED, "ias_vicreg_gram_tc: test hook needs D %% 128 == 0 (D=%d)", D);
  const Plan p = make_plan(B, D);
  float* w = reinterpret_cast<float*>(workspace);
  cudaStream_t st = as_stream(stream);
  // the hook runs x against itself; `gram` receives both (identical) sides, [2][D][D]
  return run_stats_and_gram(single_src(x, x, B), x, x, nullptr, nullptr, p, 0, B, w, gram, st);
}

extern "C" int ias_vicreg_gram_reference(const float* x, int B, int D, float* gram, void* workspace,
                                         size_t workspace_bytes, ias_stream_t stream) {
  int rc = check_common(x, B, D, workspace, workspace_bytes, "ias_vicreg_gram_reference");
  if (rc) return rc;
  IAS_REQUIRE(gram, IAS_ERR_INVALID, "ias_vicreg_gram_reference: NULL output");
  const Plan p = make_plan(B, D);
  float* w = reinterpret_cast<float*>(workspace);
  cudaStream_t st = as_stream(stream);
  k_colsum<<<p.P, 256, 0, st>>>(single_src(x, x, B), p.B, p.D, 0, B, w + p.off_partial, w + p.off_repr, nullptr, nullptr);
  IAS_LAUNCH_CHECK("k_colsum");
  k_center_pack<<<dim3(p.NV, 2), 256, 0, st>>>(x, x, p.B, p.D, p.P, p.DT, p.KB, w + p.off_partial, w + p.off_mean,
                                              w + p.off_varpart, w + p.off_packed);
  IAS_LAUNCH_CHECK("k_center_pack");
  {
    ProfScope prof_(K_VICREG_GRAM_SIMT, st);
    k_gram_simt<<<dim3((D + 15) / 16, (D + 15) / 16), 256, 0, st>>>(x, w + p.off_mean, B, D, gram);
  }
  IAS_LAUNCH_CHECK("k_gram_simt");
  return IAS_OK;
}

namespace ias {
namespace {
int run_backward(const float* x, const float* y, const Plan& p, int local_row0, int B_local, int cfgB, int embeddim,
                 float sim, float stdc, float covc, const float* gout4, float* gx, float* gy, int out_row0, int out_rows,
                 float gscale, float* w, cudaStream_t st, int Bstat = 0) {
  const int mt0 = out_row0 / TILE;
  const int mt1 = (out_row0 + out_rows + TILE - 1) / TILE;
  const int MTn = mt1 - mt0;
  {
    ProfScope prof_(K_VICREG_BWD, st);
    k_pack_rows<<<dim3(MTn, 2), 256, 0, st>>>(x, y, p.B, p.D, p.Dp, w + p.off_mean, w + p.off_rowpack, MTn, mt0);
  }
  IAS_LAUNCH_CHECK("k_pack_rows");
  {
    ProfScope prof_(K_VICREG_BWD, st);
    k_pack_offg<<<dim3(p.DT, 2), 256, 0, st>>>(w + p.off_gfull, p.Dp, w + p.off_gpack, p.DT);
  }
  IAS_LAUNCH_CHECK("k_pack_offg");
  static unsigned long long attr_devs = 0;
  const int smem = (int)sizeof(GramSmem) + 1024;
  if (ias_first_use_on_device(attr_devs)) {
    IAS_CUDA(cudaFuncSetAttribute(k_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  BwdArgs a;
  a.x = x; a.y = y;
  a.mean = w + p.off_mean;
  a.stdv = w + p.off_stats;
  a.gout4 = gout4;
  a.rowpack = w + p.off_rowpack;
  a.gpack = w + p.off_gpack;
  a.gx = gx; a.gy = gy;
  a.B = p.B; a.D = p.D; a.Dp = p.Dp; a.DT = p.DT; a.MT = MTn;
  a.local_row0 = local_row0; a.B_local = B_local; a.cfgB = cfgB; a.embeddim = embeddim;
  a.mt0 = mt0; a.out_row0 = out_row0; a.out_rows = out_rows; a.gscale = gscale;
  a.Bstat = Bstat > 0 ? Bstat : p.B;
  a.sim = sim; a.stdc = stdc; a.covc = covc;
  {
    ProfScope prof_(K_VICREG_BWD, st);
    k_bwd_tc<<<2 * MTn * p.DT, 128, smem, st>>>(a);
  }
  IAS_LAUNCH_CHECK("k_bwd_tc");
  return IAS_OK;
}

Plan gather_plan(int world, int B_local, int D, size_t* off_xg, size_t* off_yg, size_t* total) {
  const Plan p = make_plan(world * B_local, D);
  const size_t n = ((size_t)world * B_local * D + 255) / 256 * 256;
  *off_xg = p.total;
  *off_yg = p.total + n;
  *total = p.total + 2 * n;
  return p;
}
}  // namespace
}  // namespace ias

extern "C" int ias_vicreg_loss_backward(const float* x, const float* y, int B, int local_row0, int B_local,
                                        int cfg_batch_size, int D, int embeddim, float sim_coeff, float std_coeff,
                                        float cov_coeff, const float* gout4, float* gx, float* gy, void* workspace,
                                        size_t workspace_bytes, ias_stream_t stream) {
  int rc = check_common(x, B, D, workspace, workspace_bytes, "ias_vicreg_loss_backward");
  if (rc) return rc;
  IAS_REQUIRE(y && gout4 && gx && gy, IAS_ERR_INVALID, "ias_vicreg_loss_backward: NULL pointer");
  IAS_REQUIRE(local_row0 >= 0 && B_local > 0 && local_row0 + B_local <= B, IAS_ERR_INVALID,
              "ias_vicreg_loss_backward: local rows [%d,%d) outside [0,%d)", local_row0, local_row0 + B_local, B);
  IAS_REQUIRE(B > 1 && cfg_batch_size != 1 && embeddim > 0, IAS_ERR_INVALID,
              "ias_vicreg_loss_backward: B=%d cfg_batch_size=%d embeddim=%d", B, cfg_batch_size, embeddim);
  const Plan p = make_plan(B, D);
  return run_backward(x, y, p, local_row0, B_local, cfg_batch_size, embeddim, sim_coeff, std_coeff, cov_coeff, gout4,
                      gx, gy, 0, B, 1.0f, reinterpret_cast<float*>(workspace), as_stream(stream));
}

// ---- fused gather path ---------------------------------------------------------------------------------------------
extern "C" size_t ias_vicreg_gather_workspace_bytes(int world, int B_local, int D) {
  if (world <= 0 || B_local <= 0 || D <= 0) return 0;
  size_t a, b, total;
  gather_plan(world, B_local, D, &a, &b, &total);
  return total * sizeof(float);
}

extern "C" int ias_vicreg_loss_gather(const float* const* x_peers_host, const float* const* y_peers_host, int world,
                                      int rank, int B_local, int cfg_batch_size, int D, int embeddim, float sim_coeff,
                                      float std_coeff, float cov_coeff, float* out4, void* workspace,
                                      size_t workspace_bytes, ias_stream_t stream) {
  IAS_REQUIRE(x_peers_host && y_peers_host && out4, IAS_ERR_INVALID, "ias_vicreg_loss_gather: NULL pointer");
  IAS_REQUIRE(world >= 1 && world <= MAX_PEERS && rank >= 0 && rank < world && B_local > 0 && D > 0, IAS_ERR_INVALID,
              "ias_vicreg_loss_gather: world=%d rank=%d B_local=%d D=%d (at most %d peers)", world, rank, B_local, D,
              MAX_PEERS);
  IAS_REQUIRE(cfg_batch_size != 1 && embeddim > 0, IAS_ERR_INVALID, "ias_vicreg_loss_gather: cfg_batch_size=%d", cfg_batch_size);
  size_t off_xg, off_yg, total;
  const Plan p = gather_plan(world, B_local, D, &off_xg, &off_yg, &total);
  IAS_REQUIRE(workspace && workspace_bytes >= total * sizeof(float), IAS_ERR_WORKSPACE,
              "ias_vicreg_loss_gather: workspace %zu < %zu bytes", workspace_bytes, total * sizeof(float));
  IAS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023u) == 0, IAS_ERR_INVALID,
              "ias_vicreg_loss_gather: workspace must be 1024-byte aligned");
  RowSrc src;
  for (int i = 0; i < MAX_PEERS; ++i) src.x[i] = src.y[i] = nullptr;
  for (int i = 0; i < world; ++i) {
    IAS_REQUIRE(x_peers_host[i] && y_peers_host[i], IAS_ERR_INVALID, "ias_vicreg_loss_gather: peer %d pointer is NULL", i);
    src.x[i] = x_peers_host[i];
    src.y[i] = y_peers_host[i];
  }
  src.rows_per = B_local;
  float* w = reinterpret_cast<float*>(workspace);
  cudaStream_t st = as_stream(stream);
  int rc = run_stats_and_gram(src, w + off_xg, w + off_yg, w + off_xg, w + off_yg, p, rank * B_local, B_local, w,
                              w + p.off_gfull, st);
  if (rc) return rc;
  FinalArgs a;
  a.repr = w + p.off_repr;
  a.covp = w + p.off_covp;
  a.varpart = w + p.off_varpart;
  a.stats = w + p.off_stats;
  a.out4 = out4;
  a.P = p.P;
  a.NV = p.NV;
  a.ncovp_per_side = p.ntiles * (TILE / COV_ROWS);
  a.D = D; a.Dp = p.Dp; a.B = p.B; a.B_local = B_local; a.cfgB = cfg_batch_size; a.embeddim = embeddim;
  a.sim = sim_coeff; a.stdc = std_coeff; a.covc = cov_coeff;
  a.epoch = nullptr;
  {
    ProfScope prof_(K_VICREG_FINALIZE, st);
    k_finalize<<<1, FIN_THREADS, 0, st>>>(a);
  }
  IAS_LAUNCH_CHECK("k_finalize");
  return IAS_OK;
}

extern "C" int ias_vicreg_loss_gather_backward(int world, int rank, int B_local, int cfg_batch_size, int D, int embeddim,
                                               float sim_coeff, float std_coeff, float cov_coeff, const float* gout4,
                                               float* gx_local, float* gy_local, void* workspace,
                                               size_t workspace_bytes, ias_stream_t stream) {
  IAS_REQUIRE(gout4 && gx_local && gy_local, IAS_ERR_INVALID, "ias_vicreg_loss_gather_backward: NULL pointer");
  IAS_REQUIRE(world >= 1 && world <= MAX_PEERS && rank >= 0 && rank < world && B_local > 0 && D > 0, IAS_ERR_INVALID,
              "ias_vicreg_loss_gather_backward: world=%d rank=%d B_local=%d D=%d", world, rank, B_local, D);
  size_t off_xg, off_yg, total;
  const Plan p = gather_plan(world, B_local, D, &off_xg, &off_yg, &total);
  IAS_REQUIRE(workspace && workspace_bytes >= total * sizeof(float), IAS_ERR_WORKSPACE,
              "ias_vicreg_loss_gather_backward: workspace %zu < %zu bytes", workspace_bytes, total * sizeof(float));
  IAS_REQUIRE(p.B > 1, IAS_ERR_INVALID, "ias_vicreg_loss_gather_backward: global batch of 1");
  float* w = reinterpret_cast<float*>(workspace);
  // Every rank holds the same std/cov terms, so the sum over ranks of their gradients w.r.t. this rank's rows is
  // `world` times the own-row slice: no reduce-scatter is needed (FullGatherLayer.backward semantics, vicreg.py:92-95).
  return run_backward(w + off_xg, w + off_yg, p, rank * B_local, B_local, cfg_batch_size, embeddim, sim_coeff, std_coeff,
                      cov_coeff, gout4, gx_local, gy_local, rank * B_local, B_local, (float)world, w, as_stream(stream));
}

// ---- statistics exchange path ---------------------------------------------------------------------------------------
extern "C" size_t ias_vicreg_stats_buffer_bytes(int world, int D) {
  if (world <= 0 || D <= 0) return 0;
  const Plan p = make_plan(TILE, D);
  return ((size_t)XCHG_HDR_FLOATS + (size_t)world * 2 * packet_floats(p.Dp, p.ntiles)) * sizeof(float);
}

extern "C" int ias_vicreg_loss_stats(const float* x, const float* y, float* const* buffers_host, int world, int rank,
                                     int B_local, int cfg_batch_size, int D, int embeddim, float sim_coeff,
                                     float std_coeff, float cov_coeff, float* out4, void* workspace,
                                     size_t workspace_bytes, ias_stream_t stream) {
  return ias_vicreg_loss_stats_stages(x, y, buffers_host, world, rank, B_local, cfg_batch_size, D, embeddim, sim_coeff,
                                      std_coeff, cov_coeff, out4, workspace, workspace_bytes,
                                      IAS_STATS_STAGE_PUBLISH | IAS_STATS_STAGE_COMBINE, stream);
}

extern "C" int ias_vicreg_loss_stats_stages(const float* x, const float* y, float* const* buffers_host, int world,
                                            int rank, int B_local, int cfg_batch_size, int D, int embeddim,
                                            float sim_coeff, float std_coeff, float cov_coeff, float* out4,
                                            void* workspace, size_t workspace_bytes, int stages, ias_stream_t stream) {
  const bool do_publish = (stages & IAS_STATS_STAGE_PUBLISH) != 0, do_combine = (stages & IAS_STATS_STAGE_COMBINE) != 0;
  IAS_REQUIRE(do_publish || do_combine, IAS_ERR_INVALID, "ias_vicreg_loss_stats: stages=%d selects nothing", stages);
  int rc = check_common(x, B_local, D, workspace, workspace_bytes, "ias_vicreg_loss_stats");
  if (rc) return rc;
  IAS_REQUIRE(y && out4 && buffers_host, IAS_ERR_INVALID, "ias_vicreg_loss_stats: NULL pointer");
  IAS_REQUIRE(world >= 1 && world <= MAX_PEERS && rank >= 0 && rank < world, IAS_ERR_INVALID,
              "ias_vicreg_loss_stats: world=%d rank=%d (at most %d peers)", world, rank, MAX_PEERS);
  IAS_REQUIRE(cfg_batch_size != 1 && embeddim > 0 && (long long)world * B_local > 1, IAS_ERR_INVALID,
              "ias_vicreg_loss_stats: cfg_batch_size=%d embeddim=%d", cfg_batch_size, embeddim);
  const Plan p = make_plan(B_local, D);
  PeerTable peers;
  for (int i = 0; i < MAX_PEERS; ++i) peers.base[i] = nullptr;
  for (int i = 0; i < world; ++i) {
    IAS_REQUIRE(buffers_host[i] && ias_aligned16(buffers_host[i]), IAS_ERR_INVALID,
                "ias_vicreg_loss_stats: peer buffer %d is NULL or not 16-byte aligned", i);
    peers.base[i] = buffers_host[i];
  }
  float* w = reinterpret_cast<float*>(workspace);
  cudaStream_t st = as_stream(stream);
  // local rows -> local mean, centred second moments, K-split Gram partials (the single-GPU kernels, B = B_local)
  if (do_publish) {
    bool v4 = (p.D % 4 == 0) && ias_aligned16(x) && ias_aligned16(y);
    ProfScope prof_(K_VICREG_COLSUM, st);
    if (v4)
      k_colsum_v4<<<p.P, CS_THREADS, 0, st>>>(single_src(x, y, B_local), p.B, p.D, 0, B_local, w + p.off_partial,
                                              w + p.off_repr, nullptr, nullptr);
    else
      k_colsum<<<p.P, 256, 0, st>>>(single_src(x, y, B_local), p.B, p.D, 0, B_local, w + p.off_partial, w + p.off_repr,
                                    nullptr, nullptr);
  }
  IAS_LAUNCH_CHECK("k_colsum");
  if (do_publish) {
    ProfScope prof_(K_VICREG_PACK, st);
    k_center_pack<<<dim3(p.NV, 2), 256, 0, st>>>(x, y, p.B, p.D, p.P, p.DT, p.KB, w + p.off_partial, w + p.off_mean,
                                                w + p.off_varpart, w + p.off_packed);
  }
  IAS_LAUNCH_CHECK("k_center_pack");
  if (do_publish) {
    rc = launch_gram(p, w, st);
    if (rc) return rc;
  }
  const size_t pf = packet_floats(p.Dp, p.ntiles);
  if (do_publish) {
    ProfScope prof_(K_VICREG_STATS_PUBLISH, st);
    const int grid = (int)((pf / 4 + 256 / PUB_PARTS - 1) / (256 / PUB_PARTS));
    k_stats_publish<<<grid, 256, 0, st>>>(w + p.off_gram, w + p.off_varpart, w + p.off_mean, D, p.Dp, p.ntiles, p.splits,
                                          p.NV, peers, world, rank);
  }
  IAS_LAUNCH_CHECK("k_stats_publish");
  if (!do_combine) return IAS_OK;
  {
    // the combine overwrites the local mean with the global one and leaves the global centred second moments where
    // k_finalize reads variance partials (as a single partial)
    ProfScope prof_(K_VICREG_STATS_COMBINE, st);
    k_stats_combine<<<2 * p.ntiles * (TILE / COV_ROWS), 256, (size_t)world * (TILE + COV_ROWS) * sizeof(float), st>>>(peers.base[rank], world, B_local, D, p.Dp, p.DT,
                                                                   p.ntiles, w + p.off_covp, w + p.off_gfull,
                                                                   w + p.off_mean, w + p.off_varpart);
  }
  IAS_LAUNCH_CHECK("k_stats_combine");
  FinalArgs a;
  a.repr = w + p.off_repr;
  a.covp = w + p.off_covp;
  a.varpart = w + p.off_varpart;
  a.stats = w + p.off_stats;
  a.out4 = out4;
  a.P = p.P;
  a.NV = 1;
  a.ncovp_per_side = p.ntiles * (TILE / COV_ROWS);
  a.D = D; a.Dp = p.Dp; a.B = world * B_local; a.B_local = B_local; a.cfgB = cfg_batch_size; a.embeddim = embeddim;
  a.sim = sim_coeff; a.stdc = std_coeff; a.covc = cov_coeff;
  a.epoch = reinterpret_cast<int*>(peers.base[rank] + XCHG_EPOCH);
  {
    ProfScope prof_(K_VICREG_FINALIZE, st);
    k_finalize<<<1, FIN_THREADS, 0, st>>>(a);
  }
  IAS_LAUNCH_CHECK("k_finalize");
  return IAS_OK;
}

extern "C" int ias_vicreg_loss_stats_backward(const float* x, const float* y, int world, int B_local,
                                              int cfg_batch_size, int D, int embeddim, float sim_coeff,
                                              float std_coeff, float cov_coeff, const float* gout4, float* gx_local,
                                              float* gy_local, void* workspace, size_t workspace_bytes,
                                              ias_stream_t stream) {
  int rc = check_common(x, B_local, D, workspace, workspace_bytes, "ias_vicreg_loss_stats_backward");
  if (rc) return rc;
  IAS_REQUIRE(y && gout4 && gx_local && gy_local, IAS_ERR_INVALID, "ias_vicreg_loss_stats_backward: NULL pointer");
  IAS_REQUIRE(world >= 1 && (long long)world * B_local > 1 && cfg_batch_size != 1 && embeddim > 0, IAS_ERR_INVALID,
              "ias_vicreg_loss_stats_backward: world=%d B_local=%d cfg_batch_size=%d", world, B_local, cfg_batch_size);
  const Plan p = make_plan(B_local, D);
  // every rank holds the same std/cov terms: the sum over ranks of their gradients w.r.t. this rank's rows is `world`
  // times the own-row slice (FullGatherLayer.backward semantics, vicreg.py:92-95) -- no communication
  return run_backward(x, y, p, 0, B_local, cfg_batch_size, embeddim, sim_coeff, std_coeff, cov_coeff, gout4, gx_local,
                      gy_local, 0, B_local, (float)world, reinterpret_cast<float*>(workspace), as_stream(stream),
                      world * B_local);
}
