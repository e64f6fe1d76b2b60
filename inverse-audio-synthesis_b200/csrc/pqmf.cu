// PQMF analysis / synthesis for sm_100a.
//
// Replaces pqmf.PQMF.analysis / forward (pqmf.py:46-50: F.conv1d(x, H, padding=taps//2, stride=N)) and
// pqmf.PQMF.synthesis (pqmf.py:52-55: conv_transpose1d zero-stuffing with gain N, then F.conv1d with G).
//
// Both directions are register-tiled FIR kernels.  A thread owns Q consecutive decimated time steps for all N
// bands, so every input sample it needs is read from shared memory into a register once and then used by all
// (up to 63) taps that touch it; the taps are kernel arguments, i.e. live in the constant bank and are FFMA
// operands with no load instruction.  The signal tile is staged global -> shared with 128-bit coalesced loads
// and a padded layout (stride Q*N made odd) so the per-thread sliding windows are bank-conflict free.  The
// zero-stuffed [B,N,L*N] intermediate of the reference's synthesis is never materialised (polyphase indexing).
// HBM traffic: 4T read + 4T written per sound in each direction (algorithmic minimum).
#include "ias_common.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>
#include <vector>

namespace ias {
namespace {

constexpr int PQ_THREADS = 128;

// ---- TMA bulk copy (cp.async.bulk, 1-D) + mbarrier: stages a linear signal tile with ONE instruction of one thread
#ifndef IAS_PQMF_BULK
#define IAS_PQMF_BULK 1
#endif
__device__ __forceinline__ uint32_t pq_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pq_bulk_stage(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  const uint32_t b = pq_smem_u32(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   pq_smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(b)
               : "memory");
}
// Bounded wait on phase 0: a lost completion traps instead of hanging the GPU.
__device__ __forceinline__ void pq_bulk_wait(uint64_t* bar) {
  const uint32_t b = pq_smem_u32(bar);
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(b)
        : "memory");
    if (done) return;
  }
  __trap();
}

template <int N, int K>
struct Taps {
  float h[N * K];
};

// Cosine-modulated factorisation of the analysis bank (pqmf.py:21-30): H[k][j] = g[j] * c[k][j mod 2N] with
// g[j] = (-1)^floor(j/2N) * 2*prototype[j] and c[k][r] = cos((2k+1)*pi/(2N) * (r - (taps-1)/2) + (-1)^k*pi/4).
template <int N, int K>
struct TapsCM {
  float g[K + (K & 1)];  // padded so that c starts 8-byte aligned (read as float2 pairs by the packed stages)
  float c[N * 2 * N];
  // the prototype again, as the tap pairs the packed (FFMA2) polyphase stage multiplies two adjacent samples by:
  // ge[m] = (g[2m], g[2m+1]) and go[m] = (g[2m-1], g[2m]), zero outside [0, K)
  float2 ge[(K + 1) / 2];
  float2 go[(K + 1) / 2];
};

// packed fp32 pair (sm_100 fma.rn.f32x2 -> FFMA2: two IEEE fp32 lanes per instruction, same rounding as scalar FFMA)
struct P2 {
  unsigned long long v;
};
__device__ __forceinline__ P2 p2(float lo, float hi) {
  P2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void p2_unpack(P2 a, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
}
__device__ __forceinline__ P2 p2_fma(P2 a, P2 b, P2 c) {
  P2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}

// For N >= 8 the N x 2N cosine modulation is evaluated as one size-N DCT-IV on folded partial sums.  With
// t = r - (K-2)/2, every residue r maps to t' = t + 2N*q in {+-(m + 1/2)}, m < N, with sign (-1)^q (the cosine flips
// sign every 2N taps for all k).  Writing P[m] / Q[m] for the signed partial sums at t' = +(m+1/2) / -(m+1/2):
//   out[k] = sum_m cos((2k+1)(2m+1)pi/(4N)) / sqrt(2) * ( (P[m] + Q[m]) - (P[N-1-m] - Q[N-1-m]) )
// (the (-1)^k pi/4 phase of pqmf.py:28 turns the sine part into the index-reversed cosine part): N^2 + 3N instead of
// 2N^2 operations per time step.  For N >= 8 TapsCM::c holds that N x N matrix in its first N*N entries.
template <int N, int K>
struct FoldCM {
  __host__ __device__ static constexpr int wrap(int x) { return ((x % (2 * N)) + 2 * N) % (2 * N); }
  __host__ __device__ static constexpr int rP(int m) { return wrap(m + (K - 1) / 2); }
  __host__ __device__ static constexpr int rQ(int m) { return wrap((K - 3) / 2 - m); }
  __host__ __device__ static constexpr bool negP(int m) { return (((m + (K - 1) / 2) - rP(m)) / (2 * N)) & 1; }
  __host__ __device__ static constexpr bool negQ(int m) { return ((((K - 3) / 2 - m) - rQ(m)) / (2 * N)) & 1; }
};
constexpr int FOLD_MIN_N = 8;

// A thread's run of CNT consecutive output floats (CNT % 4 == 0, 16-byte aligned): 256-bit stores (sm_100
// st.global.v8.f32 -> STG.E.256) where the run is 32-byte aligned, so each store fills whole 32-byte sectors even
// though the lanes of a warp are CNT floats apart; 128-bit stores otherwise.
template <int CNT>
__device__ __forceinline__ void store_run(float* __restrict__ dst, const float (&v)[CNT], bool wide) {
  static_assert(CNT % 4 == 0, "run length");
  if (CNT % 8 == 0 && wide) {
#pragma unroll
    for (int c = 0; c + 8 <= CNT; c += 8)
      asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + c), "f"(v[c]), "f"(v[c + 1]),
                   "f"(v[c + 2]), "f"(v[c + 3]), "f"(v[c + 4]), "f"(v[c + 5]), "f"(v[c + 6]), "f"(v[c + 7])
                   : "memory");
  } else {
#pragma unroll
    for (int c = 0; c < CNT; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
  }
}

// Optional per-band epilogue (band - mean[k]) / std[k]: torchvision.transforms.Normalize on the [B,3,240,245] image
// view the reference takes of the bands (audioembed.py:41,49; constants vicreg_audio_params.py:60-62).
template <int N>
struct BandNorm {
  float mean[N];
  float std[N];
  int on;
};

// Optional pooled-magnitude epilogue (harness bridge, SURVEY 8d: adaptive_avg_pool1d(|bands|.reshape(B,1,N*L), P)):
// while a CTA still holds its band values in registers it adds up |value| per pooling bin.  A CTA's TILE_N steps of one
// band are TILE_N consecutive elements of the flattened [N*L] axis; with TILE_N + 1 <= floor(N*L / P) they touch at
// most two bins, ilo = floor(first * P / (N*L)) and ilo + 1 (torch's bins [floor(i S/P), ceil((i+1) S/P)) overlap by
// up to one element, which then counts for both).  Every warp adds its 32*Q values in 2^-22 fixed point (one
// redux.sync) into the 64-bit accumulators acc[b][ilo], acc[b][ilo + 1] (red.global.add.u64; integer adds commute, so
// the result is bit-identical from run to run); k_pool_finalize_fixed converts and divides by the bin width.  The bands
// are not read back from HBM.  (Round 1 reduced per-CTA float partials through shared memory in a fixed order: same
// determinism, one barrier, a second phase and ~80 more instructions per thread: 0.333 vs 0.3265 ms.)
struct PoolReq {  // host-side request of the pooled epilogue
  float* feat;  // [B][P]
  int P;
  void* workspace;
  size_t workspace_bytes;
};
#define IAS_POOL_SCALE 4194304.0f  /* 2^22: fixed-point resolution of the pooled sums */
// a thread's contribution saturates at 2^27 - 1 (a sum of 32 over its Q steps) so that a warp's 32 contributions cannot
// wrap 32 bits; bands of audio within [-1, 1] stay below sum_j |H[k][j]| < 2 per step
__device__ __forceinline__ unsigned pool_fixed(float t) { return min(__float2uint_rn(t * IAS_POOL_SCALE), 134217727u); }
struct PoolArgs {
  unsigned long long* acc;  // [B][P] fixed-point bin accumulators, zeroed by the launcher
  int P;
  int S;  // N * L  (S * P < 2^31 is checked by the caller)
  // floor(x / S) and floor(x / P) for x < 2^31 as __umulhi(x, m) >> sh with m = ceil(2^(32+sh) / d), 2^(sh+1) >= d
  // (Granlund-Montgomery: exact for 31-bit x); the divisors are launch constants, so the host prepares (m, sh)
  unsigned mS, shS, mP, shP;
};
// red.global.add.u64 on an address rebuilt from two 32-bit halves (the pooled epilogue keeps its accumulator addresses
// in shared memory): through a C++ pointer nvcc no longer knows the address space and emits the generic-atomic dispatch
__device__ __forceinline__ void red_global_add_u64(unsigned long long addr, unsigned long long v) {
  asm volatile("red.global.add.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned magic_div(unsigned x, unsigned m, unsigned sh) { return __umulhi(x, m) >> sh; }
inline void magic_for(unsigned d, unsigned& m, unsigned& sh) {  // d >= 2
  unsigned l = 1;
  while ((1ull << l) < d) ++l;
  m = (unsigned)(((1ull << (31 + l)) + d - 1) / d);
  sh = l - 1;
}

// Shared-memory layout of a staged signal row.  Thread t reads a sliding window that starts at linear index t*S.
//   S % 4 == 0: rows are stored in chunks of S floats at a padded pitch SP (a multiple of 4 with SP/4 odd), so the
//               128-bit window loads of a quarter warp hit 8 distinct 16-byte bank groups;
//   otherwise (S = 2: N = 16 synthesis): plain linear layout.  Staging stays a 128-bit copy with no index arithmetic;
//               the scalar window loads see a 2-way bank conflict, on ~80 loads per thread against ~2000 FMAs.
template <int S>
struct Pad {
  static constexpr bool VEC = (S % 4 == 0);
  static constexpr int SP = VEC ? ((((S / 4) & 1) == 1) ? S : S + 4) : S;
  __host__ __device__ static constexpr int at(int i) { return VEC ? (i / S) * SP + (i % S) : i; }
  __host__ __device__ static constexpr int floats(int span) { return VEC ? (span / S + 2) * SP : span + 8; }
};

// Stage `SPAN4` float4 groups of one row into shared memory: global index g0 + 4*i4 (g0 % 4 == 0), zero outside
// [0, len), scaled by `gain`.
template <int S, int SPAN4>
__device__ __forceinline__ void stage_row(float* __restrict__ dst, const float* __restrict__ row, int g0, int len,
                                          float gain, bool vec_ok) {
  using P = Pad<S>;
  constexpr int STEP = 4 * PQ_THREADS;
  int q = (4 * (int)threadIdx.x) / S, r = (4 * (int)threadIdx.x) % S;
  for (int i4 = threadIdx.x; i4 < SPAN4; i4 += PQ_THREADS) {
    const int g = g0 + 4 * i4;
    float4 v;
    if (vec_ok && g >= 0 && g + 3 < len) {
      v = __ldg(reinterpret_cast<const float4*>(row + g));
    } else {
      v.x = (g + 0 >= 0 && g + 0 < len) ? __ldg(row + g + 0) : 0.0f;
      v.y = (g + 1 >= 0 && g + 1 < len) ? __ldg(row + g + 1) : 0.0f;
      v.z = (g + 2 >= 0 && g + 2 < len) ? __ldg(row + g + 2) : 0.0f;
      v.w = (g + 3 >= 0 && g + 3 < len) ? __ldg(row + g + 3) : 0.0f;
    }
    v.x *= gain; v.y *= gain; v.z *= gain; v.w *= gain;
    if constexpr (P::VEC) {
      *reinterpret_cast<float4*>(dst + q * P::SP + r) = v;  // S % 4 == 0: a group never straddles a chunk
      q += STEP / S;
      r += STEP % S;
      if (r >= S) {
        r -= S;
        ++q;
      }
    } else {
      *reinterpret_cast<float4*>(dst + 4 * i4) = v;  // linear layout
    }
  }
}

// Interior tiles (every float4 group inside the row, 16-byte aligned, no gain): all global loads are issued before
// the first shared store, no bounds logic.  `src` points at global index g0 of the row.
template <int S, int SPAN4>
__device__ __forceinline__ void stage_row_interior(float* __restrict__ dst, const float* __restrict__ src) {
  using P = Pad<S>;
  constexpr int ITERS = (SPAN4 + PQ_THREADS - 1) / PQ_THREADS;
  const float4* s4 = reinterpret_cast<const float4*>(src) + threadIdx.x;
  float4 v[ITERS];
#pragma unroll
  for (int i = 0; i < ITERS; ++i)
    if ((i + 1) * PQ_THREADS <= SPAN4 || (int)threadIdx.x + i * PQ_THREADS < SPAN4) v[i] = __ldg(s4 + i * PQ_THREADS);
#pragma unroll
  for (int i = 0; i < ITERS; ++i)
    if ((i + 1) * PQ_THREADS <= SPAN4 || (int)threadIdx.x + i * PQ_THREADS < SPAN4) {
      const int i4 = (int)threadIdx.x + i * PQ_THREADS;
      *reinterpret_cast<float4*>(dst + P::at(4 * i4)) = v[i];  // linear layout when !VEC: at() is the identity
    }
}

// The same window as raw 128-bit groups: wr[i] = staged float t*S + i for i < NRAW (NRAW % 4 == 0), so that callers
// know which elements share an aligned register pair (i even).
template <int S, int NRAW>
__device__ __forceinline__ void load_window_raw(const float* __restrict__ src, float (&wr)[NRAW]) {
  using P = Pad<S>;
  static_assert(P::VEC && NRAW % 4 == 0, "raw windows need the 128-bit layout");
  const float* base = src + threadIdx.x * P::SP;
#pragma unroll
  for (int i = 0; i < NRAW / 4; ++i) {
    const float4 v = *reinterpret_cast<const float4*>(base + P::at(4 * i));
    wr[4 * i] = v.x; wr[4 * i + 1] = v.y; wr[4 * i + 2] = v.z; wr[4 * i + 3] = v.w;
  }
}

// Sliding window of this thread: linear floats [t*S + OFF, t*S + OFF + WIN) of the staged row -> registers.
template <int S, int OFF, int WIN>
__device__ __forceinline__ void load_window(const float* __restrict__ src, float (&w)[WIN]) {
  using P = Pad<S>;
  if constexpr (P::VEC) {
    constexpr int G = (OFF + WIN + 3) / 4;
    float4 g[G];
    const float* base = src + threadIdx.x * P::SP;
#pragma unroll
    for (int i = 0; i < G; ++i) g[i] = *reinterpret_cast<const float4*>(base + P::at(4 * i));
#pragma unroll
    for (int c = 0; c < WIN; ++c) {
      const float4 v = g[(OFF + c) / 4];
      const int e = (OFF + c) % 4;
      w[c] = e == 0 ? v.x : (e == 1 ? v.y : (e == 2 ? v.z : v.w));
    }
  } else {
#pragma unroll
    for (int c = 0; c < WIN; ++c) w[c] = src[P::at((int)threadIdx.x * S + OFF + c)];
  }
}

// Row (sound) and tile of a CTA.  The launchers use a 2-D grid (x = tile, y = row) whenever the batch fits gridDim.y:
// a 1-D grid costs every CTA an integer division by a kernel argument -- I2F, MUFU.RCP, F2I and a dozen dependent
// integer instructions before the tile's bulk copy can be issued (12 % of the analysis kernel's stall samples sat on
// that chain, profiles/r3z).  Batches beyond 65535 rows keep the 1-D grid and the division.
__device__ __forceinline__ void row_and_tile(int tiles_per_row, int& b, int& tile) {
  if (gridDim.x == (unsigned)tiles_per_row) {
    b = blockIdx.y;
    tile = blockIdx.x;
  } else {
    b = blockIdx.x / tiles_per_row;
    tile = blockIdx.x - b * tiles_per_row;
  }
}
inline dim3 row_tile_grid(int B, int tiles) {
  return B <= 65535 ? dim3((unsigned)tiles, (unsigned)B) : dim3((unsigned)((size_t)B * tiles));
}

// ------------------------------------------------------------------------------------------------------------
// analysis: out[b][k][n] = sum_j H[k][j] * x[b][n*N + j - PAD]
// ------------------------------------------------------------------------------------------------------------
template <int N, int K, int Q, class TapsT, bool POOL>
__global__ void __launch_bounds__(PQ_THREADS)
k_pqmf_analysis(const float* __restrict__ x, const float* __restrict__ row_scale, float* __restrict__ out, int T, int L,
                int tiles_per_row, TapsT taps, BandNorm<N> norm, PoolArgs pool) {
  constexpr int PAD = (K - 1) / 2;
  constexpr int S = Q * N;                 // input samples consumed per thread
  constexpr int WIN = (Q - 1) * N + K;     // input window of one thread
  constexpr int TILE_N = PQ_THREADS * Q;   // output steps per CTA
  static_assert((TILE_N * N) % 4 == 0, "tile start must keep 16-byte alignment");
  constexpr int OFF = (4 - PAD % 4) % 4;   // staged window starts at a multiple of 4 <= first needed sample
  constexpr int SPAN = TILE_N * N + K - N + 4;
  constexpr int SPAN4 = (SPAN + 3) / 4;
  __shared__ __align__(16) float xs[Pad<S>::floats(SPAN4 * 4)];
  // per band: start of bin ilo+1, end of bin ilo, and the address of this sound's accumulator of bin ilo (8-byte
  // aligned; bit 0 set when ilo is the last bin, i.e. there is no accumulator after it) -- one 128-bit load per band
  __shared__ __align__(16) int4 s_bound[POOL ? N : 1];

  int b, tile;
  row_and_tile(tiles_per_row, b, tile);
  const int n_tile = tile * TILE_N;
  const float scale = row_scale ? row_scale[b] : 1.0f;
  const int g0 = n_tile * N - PAD - OFF;  // multiple of 4
  const bool vec_ok = ((T & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0);
  // the per-row gain is applied to the accumulators (linear), so the staged samples are the raw input.
  // Interior tiles whose shared-memory layout is linear (chunk pitch == chunk size: N = 3) arrive by ONE bulk copy
  // (cp.async.bulk, the TMA engine) issued by thread 0 and tracked by an mbarrier -- no LDG / STS instructions, no
  // shared-memory store wavefronts; other layouts and the first / last tile of a row take the register path.
  constexpr bool LINEAR = Pad<S>::VEC && Pad<S>::SP == S && IAS_PQMF_BULK;
  __shared__ __align__(8) uint64_t s_bar;
  const bool interior = vec_ok && g0 >= 0 && g0 + 4 * SPAN4 <= T;  // CTA-uniform
  const bool bulk = LINEAR && interior;
  if (bulk) {
    if (threadIdx.x == 0) pq_bulk_stage(xs, x + (size_t)b * T + g0, SPAN4 * 16, &s_bar);
  } else if (interior) {
    stage_row_interior<S, SPAN4>(xs, x + (size_t)b * T + g0);
  } else {
    stage_row<S, SPAN4>(xs, x + (size_t)b * T, g0, T, 1.0f, vec_ok);
  }
  if constexpr (POOL) {
    if (threadIdx.x < N) {  // 32-bit arithmetic: S * P < 2^31 (checked by the launcher)
      const unsigned first = (unsigned)((int)threadIdx.x * L + n_tile);
      const unsigned ilo = magic_div(first * (unsigned)pool.P, pool.mS, pool.shS);
      const unsigned num = (ilo + 1u) * (unsigned)pool.S;
      const unsigned q = magic_div(num, pool.mP, pool.shP);
      const unsigned long long a =
          (unsigned long long)__cvta_generic_to_global(pool.acc + (size_t)b * pool.P + ilo) | (ilo + 1u < (unsigned)pool.P ? 0ull : 1ull);
      s_bound[threadIdx.x] = make_int4((int)q, (int)(q + (num - q * (unsigned)pool.P != 0u ? 1u : 0u)),
                                       (int)(unsigned)a, (int)(unsigned)(a >> 32));
    }
  }
  __syncthreads();  // staged tile (register path), bin bounds and the mbarrier's initialisation are visible
  if (bulk) pq_bulk_wait(&s_bar);

  float acc[Q][N];
#pragma unroll
  for (int q = 0; q < Q; ++q)
#pragma unroll
    for (int k = 0; k < N; ++k) acc[q][k] = 0.0f;
  if constexpr (std::is_same<TapsT, Taps<N, K>>::value) {
    // direct form: 63*N FMA per time step; valid for any H (e.g. loaded from a checkpoint)
    float w[WIN];
    load_window<S, OFF, WIN>(xs, w);
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
      for (int k = 0; k < N; ++k)
#pragma unroll
        for (int q = 0; q < Q; ++q) acc[q][k] = fmaf(taps.h[j * N + k], w[q * N + j], acc[q][k]);  // tap-major copy
  } else {
    // polyphase form: fold the 63 taps into 2N partial sums shared by all bands (63 FMA per time step), then the
    // N x 2N cosine modulation (2N*N FMA per time step): 63 + 2N^2 instead of 63N
    float ps[Q][2 * N];
    if constexpr (K % 2 == 1 && Pad<S>::VEC) {
      // two adjacent taps per FFMA2: the kernel is issue bound, and the 63 prototype multiply-adds per time step are
      // most of its instructions.  A pair needs its two samples in one aligned register pair, i.e. at an even raw
      // index of the 128-bit window groups: steps whose first sample sits at an even index pair taps (0,1), (2,3),
      // ..., steps at an odd index pair (-1,0), (1,2), ... (tap -1 and tap K are zero, their sample lane is a
      // literal 0).  Each lane is the same fma chain, in the same tap order, as the scalar form.
      constexpr int NRAW = (OFF + WIN + 3) / 4 * 4;
      constexpr int NP = (K + 1) / 2;
      float wr[NRAW];
      load_window_raw<S, NRAW>(xs, wr);
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const int base = q * N + OFF;  // raw index of the sample tap 0 multiplies (compile-time after unrolling)
        P2 pp[N];
#pragma unroll
        for (int i = 0; i < N; ++i) pp[i] = p2(0.0f, 0.0f);
        if (base % 2 == 0) {
#pragma unroll
          for (int m = 0; m < NP; ++m) {
            const float2 t = taps.ge[m];
            const P2 x = (2 * m + 1 < K) ? p2(wr[base + 2 * m], wr[base + 2 * m + 1]) : p2(wr[base + 2 * m], 0.0f);
            const int i = ((2 * m) % (2 * N)) / 2;
            pp[i] = p2_fma(p2(t.x, t.y), x, pp[i]);
          }
#pragma unroll
          for (int i = 0; i < N; ++i) p2_unpack(pp[i], ps[q][2 * i], ps[q][2 * i + 1]);
        } else {
#pragma unroll
          for (int m = 0; m < NP; ++m) {
            const float2 t = taps.go[m];
            const P2 x = (m > 0) ? p2(wr[base + 2 * m - 1], wr[base + 2 * m]) : p2(0.0f, wr[base]);
            const int i = (((2 * m - 1) % (2 * N) + 2 * N) % (2 * N) - 1) / 2;  // lanes (2i+1, 2i+2 mod 2N)
            pp[i] = p2_fma(p2(t.x, t.y), x, pp[i]);
          }
#pragma unroll
          for (int i = 0; i < N; ++i) p2_unpack(pp[i], ps[q][2 * i + 1], ps[q][(2 * i + 2) % (2 * N)]);
        }
      }
    } else {
      float w[WIN];
      load_window<S, OFF, WIN>(xs, w);
#pragma unroll
      for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int r = 0; r < 2 * N; ++r) ps[q][r] = 0.0f;
#pragma unroll
      for (int j = 0; j < K; ++j)
#pragma unroll
        for (int q = 0; q < Q; ++q) ps[q][j % (2 * N)] = fmaf(taps.g[j], w[q * N + j], ps[q][j % (2 * N)]);
    }
    if constexpr (N >= FOLD_MIN_N) {
      using F = FoldCM<N, K>;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        float sum[N], dif[N];  // P + Q and P - Q with the residue signs applied
#pragma unroll
        for (int m = 0; m < N; ++m) {
          const float pm = F::negP(m) ? -ps[q][F::rP(m)] : ps[q][F::rP(m)];
          const float qm = F::negQ(m) ? -ps[q][F::rQ(m)] : ps[q][F::rQ(m)];
          sum[m] = pm + qm;
          dif[m] = pm - qm;
        }
        // C is symmetric: row m is contiguous (vectorised constant loads); two bands per FFMA2
        P2 ap[N / 2];
#pragma unroll
        for (int k2 = 0; k2 < N / 2; ++k2) ap[k2] = p2(0.0f, 0.0f);
#pragma unroll
        for (int m = 0; m < N; ++m) {
          const float v = sum[m] - dif[N - 1 - m];
          const P2 vv = p2(v, v);
#pragma unroll
          for (int k2 = 0; k2 < N / 2; ++k2) {
            const float2 t = reinterpret_cast<const float2*>(taps.c)[(m * N) / 2 + k2];
            ap[k2] = p2_fma(p2(t.x, t.y), vv, ap[k2]);
          }
        }
#pragma unroll
        for (int k2 = 0; k2 < N / 2; ++k2) p2_unpack(ap[k2], acc[q][2 * k2], acc[q][2 * k2 + 1]);
      }
    } else {
#pragma unroll
      for (int r = 0; r < 2 * N; ++r)
#pragma unroll
        for (int k = 0; k < N; ++k)
#pragma unroll
          for (int q = 0; q < Q; ++q) acc[q][k] = fmaf(taps.c[k * 2 * N + r], ps[q][r], acc[q][k]);
    }
  }

  if (row_scale) {
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
      for (int k = 0; k < N; ++k) acc[q][k] *= scale;
  }

  if (norm.on) {  // tensor.sub_(mean).div_(std), fp32 ops as torch issues them
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
      for (int k = 0; k < N; ++k) acc[q][k] = __fdiv_rn(__fsub_rn(acc[q][k], norm.mean[k]), norm.std[k]);
  }

  const int n0 = n_tile + threadIdx.x * Q;
  float* ob = out + (size_t)b * N * L;
  const bool st_vec = ((L & 3) == 0) && (Q % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
#pragma unroll
  for (int k = 0; k < N; ++k) {
    float* o = ob + (size_t)k * L + n0;
    if (st_vec && n0 + Q <= L) {
#pragma unroll
      for (int q = 0; q < Q; q += 4)
        *reinterpret_cast<float4*>(o + q) = make_float4(acc[q][k], acc[q + 1][k], acc[q + 2][k], acc[q + 3][k]);
    } else {
#pragma unroll
      for (int q = 0; q < Q; ++q)
        if (n0 + q < L) o[q] = acc[q][k];
    }
  }

  if constexpr (POOL) {
    // Fixed-point pooling: a thread's sum of |v| over its Q steps of one band is rounded to a multiple of 2^-22 (bands
    // of a normalised clip are bounded by sum|H| < 2, so a warp's 32 sums fit 32 bits), the warp adds them with ONE
    // redux.sync, and lane 0 adds the warp's total to the 64-bit accumulator of its bin with one red.global.add.u64.
    // Integer adds commute, so the result does not depend on the order in which warps and CTAs arrive -- bit-identical
    // from run to run like the shared-memory tree this replaces, with no shared memory, no barrier and no second
    // phase.  A warp whose 32*Q elements touch the bin boundary (one in ~5) sends two masked sums.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (n_tile + TILE_N > L) {  // last tile of a row (CTA-uniform, rare): steps past the end contribute nothing
      const int nv = L - n0;
#pragma unroll
      for (int q = 0; q < Q; ++q)
        if (q >= nv) {
#pragma unroll
          for (int k = 0; k < N; ++k) acc[q][k] = 0.0f;
        }
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
      float a[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) a[q] = fabsf(acc[q][k]);
      const int4 bd = s_bound[k];
      const int s1 = bd.x, e0 = bd.y;
      const unsigned long long tagged = (unsigned long long)(unsigned)bd.z | ((unsigned long long)(unsigned)bd.w << 32);
      const unsigned long long bin = tagged & ~7ull;  // global address of the accumulator of bin ilo
      const int wfirst = k * L + n_tile + warp * 32 * Q, wlast = wfirst + 32 * Q - 1;
      if (wlast < s1 || wfirst >= e0) {  // warp-uniform: the whole warp lies in one bin
        float t = a[0];
#pragma unroll
        for (int q = 1; q < Q; ++q) t += a[q];
        const unsigned tot = __reduce_add_sync(0xffffffffu, pool_fixed(t));
        if (lane == 0) red_global_add_u64(bin + (wfirst >= e0 ? 8ull : 0ull), (unsigned long long)tot);
      } else {
        const int f0 = k * L + n0;
        float sum0 = 0.0f, sum1 = 0.0f;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          sum0 += (f0 + q < e0) ? a[q] : 0.0f;
          sum1 += (f0 + q >= s1) ? a[q] : 0.0f;
        }
        const unsigned t0s = __reduce_add_sync(0xffffffffu, pool_fixed(sum0));
        const unsigned t1s = __reduce_add_sync(0xffffffffu, pool_fixed(sum1));
        if (lane == 0) {
          red_global_add_u64(bin, (unsigned long long)t0s);
          if ((tagged & 1ull) == 0) red_global_add_u64(bin + 8ull, (unsigned long long)t1s);
        }
      }
    }
  }
}

// feat[b][i] = mean of |bands_flat[b][s_i .. e_i)| from the fixed-point bin accumulators of the pooled analysis epilogue.
__global__ void k_pool_finalize_fixed(const unsigned long long* __restrict__ acc, float* __restrict__ feat, int B, int N,
                                      int L, int P) {
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (unsigned)(B * P)) return;
  const unsigned uP = (unsigned)P, S = (unsigned)N * (unsigned)L;  // S * P < 2^31 (checked by the launcher)
  const unsigned i = idx % uP;
  const unsigned s = (i * S) / uP, e = ((i + 1u) * S + uP - 1u) / uP;
  feat[idx] = (float)((double)acc[idx] * (1.0 / (double)IAS_POOL_SCALE) / (double)(e - s));
}

// Any (N, K): one thread per output, taps from global memory.  Correct for shapes without a specialised kernel.
__global__ void k_pqmf_analysis_generic(const float* __restrict__ x, const float* __restrict__ H,
                                        const float* __restrict__ row_scale, const float* __restrict__ norm_dev,
                                        float* __restrict__ out, int B, int T, int N, int K, int L) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * N * L;
  if (idx >= total) return;
  const int n = (int)(idx % L);
  const int k = (int)((idx / L) % N);
  const int b = (int)(idx / ((size_t)L * N));
  const int pad = (K - 1) / 2;
  const float* xr = x + (size_t)b * T;
  const float scale = row_scale ? row_scale[b] : 1.0f;
  float acc = 0.0f;
  for (int j = 0; j < K; ++j) {
    const int g = n * N + j - pad;
    if (g >= 0 && g < T) acc = fmaf(H[k * K + j], xr[g] * scale, acc);
  }
  if (norm_dev) acc = __fdiv_rn(__fsub_rn(acc, norm_dev[k]), norm_dev[N + k]);  // [mean[N] | std[N]]
  out[idx] = acc;
}

// ------------------------------------------------------------------------------------------------------------
// synthesis: y[b][n*N + p] = sum_k sum_{j : p(j) = p} G[k][j] * (N * z[b][k][n + o(j)])
//   p(j) = (PAD - j) mod N,  o(j) = (p(j) + j - PAD) / N   (tap j only ever meets a non-zero of the zero-stuffed
//   signal on output phase p(j))
// ------------------------------------------------------------------------------------------------------------
template <int N, int K>
struct SynthGeom {
  static constexpr int PAD = (K - 1) / 2;
  // tap j contributes to output phase p(j) with source offset o(j): y[n*N + p] += G[k][j] * N*z[k][n + o]
  __host__ __device__ static constexpr int phase(int j) { return (((PAD - j) % N) + N) % N; }
  __host__ __device__ static constexpr int offset(int j) { return (phase(j) + j - PAD) / N; }  // exact division
  __host__ __device__ static constexpr int omin() {
    int m = 1000000;
    for (int j = 0; j < K; ++j) m = offset(j) < m ? offset(j) : m;
    return m;
  }
  __host__ __device__ static constexpr int omax() {
    int m = -1000000;
    for (int j = 0; j < K; ++j) m = offset(j) > m ? offset(j) : m;
    return m;
  }
};

template <int N, int K, int Q>
__global__ void __launch_bounds__(PQ_THREADS)
k_pqmf_synthesis(const float* __restrict__ z, float* __restrict__ y, int L, int tiles_per_row, Taps<N, K> taps) {
  using Geo = SynthGeom<N, K>;
  constexpr int DMIN = Geo::omin();
  constexpr int WIN = Q + Geo::omax() - DMIN;  // per-band window of one thread: n0+DMIN .. n0+Q-1+omax
  constexpr int TILE_N = PQ_THREADS * Q;
  static_assert(TILE_N % 4 == 0, "tile start must keep 16-byte alignment");
  constexpr int OFF = (((DMIN % 4) + 4) % 4);  // staged window starts at a multiple of 4 <= n_tile + DMIN
  constexpr int SPAN = TILE_N + WIN - Q + 4;
  constexpr int SPAN4 = (SPAN + 3) / 4;
  constexpr int ROW = Pad<Q>::floats(SPAN4 * 4);
  __shared__ __align__(16) float zs[N * ROW];

  int b, tile;
  row_and_tile(tiles_per_row, b, tile);
  const int n_tile = tile * TILE_N;
  const int g0 = n_tile + DMIN - OFF;  // multiple of 4
  const bool vec_ok = ((L & 3) == 0) && ((reinterpret_cast<uintptr_t>(z) & 15u) == 0);
  // fp32 product N*z like conv_transpose1d with the updown filter scaled by N (pqmf.py:53)
  for (int k = 0; k < N; ++k) stage_row<Q, SPAN4>(zs + k * ROW, z + ((size_t)b * N + k) * L, g0, L, (float)N, vec_ok);
  __syncthreads();

  float acc[Q][N];
#pragma unroll
  for (int q = 0; q < Q; ++q)
#pragma unroll
    for (int p = 0; p < N; ++p) acc[q][p] = 0.0f;

#pragma unroll
  for (int k = 0; k < N; ++k) {
    float w[WIN];
    load_window<Q, OFF, WIN>(zs + k * ROW, w);
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const float g = taps.h[k * K + j];
#pragma unroll
      for (int q = 0; q < Q; ++q)
        acc[q][Geo::phase(j)] = fmaf(g, w[q + Geo::offset(j) - DMIN], acc[q][Geo::phase(j)]);
    }
  }

  const int n0 = n_tile + threadIdx.x * Q;
  float* yo = y + (size_t)b * L * N + (size_t)n0 * N;
  const bool st_vec = (((size_t)L * N) % 4 == 0) && ((Q * N) % 4 == 0) && ((reinterpret_cast<uintptr_t>(y) & 15u) == 0);
  if (st_vec && n0 + Q <= L) {
    float flat[Q * N];
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
      for (int p = 0; p < N; ++p) flat[q * N + p] = acc[q][p];
    store_run<Q * N>(yo, flat, (reinterpret_cast<uintptr_t>(yo) & 31u) == 0);
  } else {
#pragma unroll
    for (int q = 0; q < Q; ++q)
      if (n0 + q < L) {
#pragma unroll
        for (int p = 0; p < N; ++p) yo[q * N + p] = acc[q][p];
      }
  }
}

// ------------------------------------------------------------------------------------------------------------
// synthesis, cosine-modulated form (N >= 8, G = the filter PQMF.__init__ designs, pqmf.py:18-30):
//   G[k][j] = g[j] * cos(theta_k (r - (K-2)/2) - (-1)^k pi/4), r = j mod 2N, g[j] = (-1)^floor(j/2N) 2 prototype[j]
//   v[m][r]  = sum_k cos(...) * N z[k][m]                         (modulation of time step m, shared by all taps)
//   y[nN+p]  = sum_{j : p(j) = p} g[j] * v[n + o(j)][j mod 2N]    (63 multiply-adds per time step)
// and the N -> 2N modulation is one size-N DCT-IV U[m'] = sum_k C[k][m'] z[k] unfolded with signs: with
// r - (K-2)/2 + 2N q = +-(m' + 1/2),  v[r] = (-1)^q (N / sqrt 2) (U[m'] +- U[N-1-m'])  (the -(-1)^k pi/4 phase turns
// the sine part into the index-reversed cosine part).  N^2 + 2N + 63 instead of 63 N operations per time step
// (N = 16: 351 instead of 1008).
// Phase 1: thread = one time step of the tile + halo; its N band values come straight from global memory (coalesced
// over time), DCT-IV from the constant bank, 2N modulated values to shared memory (row pitch 2N+4 floats: the 128-bit
// accesses of a quarter warp fall into 8 distinct bank groups).  Phase 2: thread = one time step, N outputs, its
// 63 (row, r) operands read as 128-bit shared loads.
// ------------------------------------------------------------------------------------------------------------
template <int N, int K>
struct TapsSynCM {
  float g[K + (K & 1)];  // padded so that c starts 8-byte aligned
  float c[N * N];  // (N / sqrt 2) cos((2k+1)(2m+1) pi / (4N)), [k][m]
};

template <int N, int K>
struct UnfoldCM {
  // twice (r - (K-2)/2) wrapped into (-2N, 2N): an odd integer
  __host__ __device__ static constexpr int tt(int r) {
    return ((2 * r - (K - 2) + 2 * N) % (4 * N) + 4 * N) % (4 * N) - 2 * N;
  }
  __host__ __device__ static constexpr int mp(int r) { return ((tt(r) < 0 ? -tt(r) : tt(r)) - 1) / 2; }
  __host__ __device__ static constexpr bool pos(int r) { return tt(r) > 0; }
  __host__ __device__ static constexpr bool neg(int r) { return (((tt(r) - (2 * r - (K - 2))) / (4 * N)) & 1) != 0; }
};

template <int N, int K>
__global__ void __launch_bounds__(PQ_THREADS)
k_pqmf_synthesis_cm(const float* __restrict__ z, float* __restrict__ y, int L, int tiles_per_row, TapsSynCM<N, K> taps) {
  using Geo = SynthGeom<N, K>;
  using UF = UnfoldCM<N, K>;
  constexpr int DMIN = Geo::omin();
  constexpr int HALO = Geo::omax() - DMIN;
  constexpr int TILE_N = PQ_THREADS - HALO;  // time steps per CTA; every thread modulates one row
  constexpr int PITCH = 2 * N + 4;
  static_assert(N % 4 == 0, "128-bit rows");
  __shared__ __align__(16) float vs[PQ_THREADS * PITCH];

  int b, tile;
  row_and_tile(tiles_per_row, b, tile);
  const int n_tile = tile * TILE_N;

  {
    const int m = n_tile + DMIN + (int)threadIdx.x;
    float zk[N];
    const bool in = (m >= 0 && m < L);
    // one pointer walked from band to band: `zp + (size_t)k * L` costs a 64-bit multiply-add chain per load (ten
    // integer instructions each, a quarter of the kernel's instructions at N = 16: ncu, run r4x; N = 8 0.338 -> 0.328 ms,
    // N = 16 0.278 -> 0.276 ms -- that kernel waits on its stores, not on issue slots)
    const float* zp = z + (size_t)b * N * L + (in ? m : 0);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      zk[k] = in ? __ldg(zp) : 0.0f;
      zp += L;
    }
    float u[N];
    {
      P2 up[N / 2];  // two DCT outputs per FFMA2
#pragma unroll
      for (int m2 = 0; m2 < N / 2; ++m2) up[m2] = p2(0.0f, 0.0f);
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const P2 zz = p2(zk[k], zk[k]);
#pragma unroll
        for (int m2 = 0; m2 < N / 2; ++m2) {
          const float2 t = reinterpret_cast<const float2*>(taps.c)[(k * N) / 2 + m2];
          up[m2] = p2_fma(p2(t.x, t.y), zz, up[m2]);
        }
      }
#pragma unroll
      for (int m2 = 0; m2 < N / 2; ++m2) p2_unpack(up[m2], u[2 * m2], u[2 * m2 + 1]);
    }
    float v[2 * N];
#pragma unroll
    for (int r = 0; r < 2 * N; ++r) {
      const float a = u[UF::mp(r)], bb = u[N - 1 - UF::mp(r)];
      const float s = UF::pos(r) ? a + bb : a - bb;
      v[r] = UF::neg(r) ? -s : s;
    }
    float* row = vs + threadIdx.x * PITCH;
#pragma unroll
    for (int r = 0; r < 2 * N; r += 4) *reinterpret_cast<float4*>(row + r) = make_float4(v[r], v[r + 1], v[r + 2], v[r + 3]);
  }
  __syncthreads();

  if ((int)threadIdx.x >= TILE_N) return;
  const int n0 = n_tile + (int)threadIdx.x;
  if (n0 >= L) return;
  float acc[N];
#pragma unroll
  for (int p = 0; p < N; ++p) acc[p] = 0.0f;
#pragma unroll
  for (int rho = 0; rho <= HALO; ++rho) {
    // operands of this row: taps j with o(j) - DMIN == rho, at r = j mod 2N
    const float* row = vs + ((int)threadIdx.x + rho) * PITCH;
    float w[2 * N];
#pragma unroll
    for (int r4 = 0; r4 < 2 * N; r4 += 4) {
      bool need = false;
#pragma unroll
      for (int j = 0; j < K; ++j)
        if (Geo::offset(j) - DMIN == rho && (j % (2 * N)) / 4 == r4 / 4) need = true;
      if (need) {
        const float4 t = *reinterpret_cast<const float4*>(row + r4);
        w[r4] = t.x; w[r4 + 1] = t.y; w[r4 + 2] = t.z; w[r4 + 3] = t.w;
      }
    }
#pragma unroll
    for (int j = 0; j < K; ++j)
      if (Geo::offset(j) - DMIN == rho) acc[Geo::phase(j)] = fmaf(taps.g[j], w[j % (2 * N)], acc[Geo::phase(j)]);
  }
  float* yo = y + (size_t)b * L * N + (size_t)n0 * N;
  if ((reinterpret_cast<uintptr_t>(y) & 15u) == 0 && (((size_t)L * N) % 4 == 0)) {
    store_run<N>(yo, acc, (reinterpret_cast<uintptr_t>(yo) & 31u) == 0);
  } else {
#pragma unroll
    for (int p = 0; p < N; ++p) yo[p] = acc[p];
  }
}

// ------------------------------------------------------------------------------------------------------------
// synthesis, cosine-modulated form for N >= 8, second formulation of the FIR phase (round 2, session 3).
// k_pqmf_synthesis_cm runs at 87 % of the L1 / shared-memory wavefront peak (ncu, run r4x), 52 % of those wavefronts
// being the FIR phase's row loads: a time step reads HALO + 1 row halves (taps N rho .. N rho + N - 1 meet row
// s + rho, half rho & 1), each of which a second step reads again.  Here a thread owns TWO steps of one parity,
// s and s + 2, which share all but two of their row halves at the same tap alignment: HALO + 3 instead of
// 2 (HALO + 1) half loads per two steps (N = 16: 6 instead of 8, N = 8: 10 instead of 16).  Lanes 2p / 2p+1 own the
// even / odd steps of a four-step block; a CTA modulates 128 rows with all threads and runs the FIR phase on
// (128 - HALO) / 4 * 4 steps with half of them.  Row halves are stored REVERSED, so that element pairs line up with
// accumulator pairs in output order and the FIR phase is FFMA2s on (tap pair from a uniform register) x (pair of a
// 128-bit shared load): acc[k], acc[k+1] += g[N rho + N-1-k], g[N rho + N-2-k] * half'[k], half'[k+1].
// Row r starts at 16-byte unit (2N/4 + 1) r + 2 (r/8) + 4 (r/16): found by exhaustive search over small pads, free of
// bank conflicts for the phase-1 stores (lanes = consecutive rows) and the phase-2 loads (lane pairs 4 rows apart).
// Per output the taps are applied in ascending order as in the first formulation: bit-identical results.
// ------------------------------------------------------------------------------------------------------------
template <int N, int K>
struct TapsSynCM2 {
  static constexpr int NR = SynthGeom<N, K>::omax() - SynthGeom<N, K>::omin() + 1;
  float c[N * N];            // (N / sqrt 2) cos((2k+1)(2m+1) pi / (4N)), [k][m]
  float2 gp[NR * (N / 2)];   // [rho][k/2] = (g[N rho + N-1-k], g[N rho + N-2-k]), k even; 0 beyond tap K-1
};

template <int N>
__host__ __device__ constexpr int cm2_row_unit(int r) { return (2 * N / 4 + 1) * r + 2 * (r / 8) + 4 * (r / 16); }

template <int N, int K>
__global__ void __launch_bounds__(PQ_THREADS)
k_pqmf_synthesis_cm2(const float* __restrict__ z, float* __restrict__ y, int L, int tiles_per_row, TapsSynCM2<N, K> taps) {
  using Geo = SynthGeom<N, K>;
  using UF = UnfoldCM<N, K>;
  constexpr int DMIN = Geo::omin();
  constexpr int HALO = Geo::omax() - DMIN;
  constexpr int TILE_N = (PQ_THREADS - HALO) / 4 * 4;  // time steps per CTA (whole four-step blocks)
  constexpr int ACTIVE = TILE_N / 2;                   // FIR-phase threads
  static_assert(N % 4 == 0 && Geo::phase(0) == N - 1 && Geo::offset(N) - Geo::offset(0) == 1, "tap blocks of N per row");
  static_assert(TILE_N - 1 + HALO < PQ_THREADS, "every row is modulated by one thread");
  __shared__ __align__(16) float vs[(cm2_row_unit<N>(PQ_THREADS - 1) + 2 * N / 4) * 4];

  int b, tile;
  row_and_tile(tiles_per_row, b, tile);
  const int n_tile = tile * TILE_N;

  {
    const int m = n_tile + DMIN + (int)threadIdx.x;
    float zk[N];
    const bool in = (m >= 0 && m < L);
    const float* zp = z + (size_t)b * N * L + (in ? m : 0);  // walked from band to band (no 64-bit multiply per load)
#pragma unroll
    for (int k = 0; k < N; ++k) {
      zk[k] = in ? __ldg(zp) : 0.0f;
      zp += L;
    }
    float u[N];
    {
      P2 up[N / 2];  // two DCT outputs per FFMA2
#pragma unroll
      for (int m2 = 0; m2 < N / 2; ++m2) up[m2] = p2(0.0f, 0.0f);
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const P2 zz = p2(zk[k], zk[k]);
#pragma unroll
        for (int m2 = 0; m2 < N / 2; ++m2) {
          const float2 t = reinterpret_cast<const float2*>(taps.c)[(k * N) / 2 + m2];
          up[m2] = p2_fma(p2(t.x, t.y), zz, up[m2]);
        }
      }
#pragma unroll
      for (int m2 = 0; m2 < N / 2; ++m2) p2_unpack(up[m2], u[2 * m2], u[2 * m2 + 1]);
    }
    float v[2 * N];
#pragma unroll
    for (int r = 0; r < 2 * N; ++r) {
      const float a = u[UF::mp(r)], bb = u[N - 1 - UF::mp(r)];
      const float sgn = UF::pos(r) ? a + bb : a - bb;
      v[r] = UF::neg(r) ? -sgn : sgn;
    }
    // half h reversed: position k of the half holds v[N h + N-1-k]
    float* row = vs + cm2_row_unit<N>((int)threadIdx.x) * 4;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int k = 0; k < N; k += 4)
        *reinterpret_cast<float4*>(row + h * N + k) = make_float4(v[N * h + N - 1 - k], v[N * h + N - 2 - k],
                                                                  v[N * h + N - 3 - k], v[N * h + N - 4 - k]);
  }
  __syncthreads();

  if ((int)threadIdx.x >= ACTIVE) return;
  const int pr = (int)threadIdx.x >> 1, par = (int)threadIdx.x & 1;
  const int sA = 4 * pr + par;  // tile-relative first step; the second one is sA + 2
  if (n_tile + sA >= L) return;
  P2 acc[2][N / 2];
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int k2 = 0; k2 < N / 2; ++k2) acc[q][k2] = p2(0.0f, 0.0f);
#pragma unroll
  for (int j = 0; j < HALO + 3; ++j) {
    // row sA + j is met by step sA at rho = j and by step sA + 2 at rho = j - 2: same parity, same half
    const float* half = vs + cm2_row_unit<N>(sA + j) * 4 + (j & 1) * N;
    P2 w[N / 2];
#pragma unroll
    for (int k = 0; k < N; k += 4) {
      const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(half + k);
      w[k / 2].v = t.x;
      w[k / 2 + 1].v = t.y;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int rho = j - 2 * q;
      if (rho >= 0 && rho <= HALO) {
#pragma unroll
        for (int k2 = 0; k2 < N / 2; ++k2) {
          const float2 g = taps.gp[rho * (N / 2) + k2];
          acc[q][k2] = p2_fma(w[k2], p2(g.x, g.y), acc[q][k2]);
        }
      }
    }
  }
  const bool st_vec = (reinterpret_cast<uintptr_t>(y) & 15u) == 0;  // L * N and the run starts are multiples of 4 floats
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int n = n_tile + sA + 2 * q;
    if (n < L) {
      float f[N];
#pragma unroll
      for (int k2 = 0; k2 < N / 2; ++k2) p2_unpack(acc[q][k2], f[2 * k2], f[2 * k2 + 1]);
      float* yo = y + (size_t)b * L * N + (size_t)n * N;
      if (st_vec) {
        store_run<N>(yo, f, (reinterpret_cast<uintptr_t>(yo) & 31u) == 0);
      } else {
#pragma unroll
        for (int p = 0; p < N; ++p) yo[p] = f[p];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// synthesis, cosine-modulated form for small N (2, 3, 4): same factorisation as above with the N -> 2N modulation
// done directly (2 N^2 multiply-adds per time step), and Q consecutive time steps per thread in the FIR phase so
// that each modulated row read from shared memory serves Q time steps.  The N taps that meet one (time step, row)
// pair are consecutive, and their residues j mod 2N alternate between two fixed halves of the row as the offset
// advances; the row is stored as those two halves, each padded to 4 floats, so a (row, half) operand set is one
// 128-bit shared load.  16 bytes of padding after every Q rows make both the phase-1 stores (one row per thread)
// and the phase-2 loads (thread stride Q rows) conflict free.  63 + 2 N^2 instead of 63 N multiply-adds per step.
// ------------------------------------------------------------------------------------------------------------
template <int N, int K>
struct TapsSynSmall {
  float g[K];
  float c[N * 2 * N];  // N cos(theta_k (r - (K-2)/2) - (-1)^k pi/4), [k][r]
};

template <int N, int K>
struct SynRows {
  using Geo = SynthGeom<N, K>;
  static constexpr int PAD = (K - 1) / 2;
  static constexpr int DMIN = Geo::omin();
  static constexpr int HALO = Geo::omax() - DMIN;
  static constexpr int N4 = (N + 3) / 4 * 4;
  // first tap of the block that meets offset o (may be negative / beyond K-1 at the ends)
  __host__ __device__ static constexpr int jb(int o) { return PAD + N * (o - 1) + 1; }
  __host__ __device__ static constexpr int ra() { return ((jb(DMIN) % (2 * N)) + 2 * N) % (2 * N); }
  // residue stored at (half h, element e) of a row
  __host__ __device__ static constexpr int res(int h, int e) { return (ra() + h * N + e) % (2 * N); }
};

template <int N, int K, int Q>
__global__ void __launch_bounds__(PQ_THREADS)
k_pqmf_synthesis_small(const float* __restrict__ z, float* __restrict__ y, int L, int tiles_per_row,
                       TapsSynSmall<N, K> taps) {
  using R = SynRows<N, K>;
  constexpr int DMIN = R::DMIN, HALO = R::HALO, N4 = R::N4;
  constexpr int TILE_N = PQ_THREADS * Q;
  constexpr int ROWS = TILE_N + HALO;
  constexpr int PASSES = (ROWS + PQ_THREADS - 1) / PQ_THREADS;
  // Shared layout: one plane per row half, the half of row r at 16-byte unit r + r/Q of its plane (N4 == 4).  Phase 2
  // (lane t reads rows Q t + i: units (Q + 1) t + i + i/Q, Q + 1 odd) is free of bank conflicts and its offsets are
  // compile-time constants; phase 1 (lane = consecutive rows) is conflict free for Q = 8 and has one 2-way pair per
  // quarter warp for Q = 4.  The interleaved [row][half] layout this replaces had 2-way conflicts on every phase-1
  // store (ncu: 24.9 M conflicts at N = 3).
  static_assert(N4 == 4, "plane layout assumes one 16-byte unit per row half");
  constexpr int UNITS = ROWS + ROWS / Q + 1;
  __shared__ __align__(16) float vs[2 * UNITS * 4];

  int b, tile;
  row_and_tile(tiles_per_row, b, tile);
  const int n_tile = tile * TILE_N;
  const float* zb = z + (size_t)b * N * L;

  // ---- phase 1: modulate rows n_tile + DMIN + [0, ROWS) ----
  float zk[PASSES][N];
#pragma unroll
  for (int i = 0; i < PASSES; ++i) {
    const int row = (int)threadIdx.x + i * PQ_THREADS;
    const int m = n_tile + DMIN + row;
    const bool in = row < ROWS && m >= 0 && m < L;
#pragma unroll
    for (int k = 0; k < N; ++k) zk[i][k] = in ? __ldg(zb + (size_t)k * L + m) : 0.0f;  // (a walked pointer: N=4 0.389 vs 0.372 ms)
  }
#pragma unroll
  for (int i = 0; i < PASSES; ++i) {
    const int row = (int)threadIdx.x + i * PQ_THREADS;
    if (row < ROWS) {
      float* dst = vs + (row + row / Q) * 4;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v[N4];
#pragma unroll
        for (int e = 0; e < N4; ++e) {
          v[e] = 0.0f;
          if (e < N) {
#pragma unroll
            for (int k = 0; k < N; ++k) v[e] = fmaf(taps.c[k * 2 * N + R::res(h, e)], zk[i][k], v[e]);
          }
        }
        *reinterpret_cast<float4*>(dst + h * UNITS * 4) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
  }
  __syncthreads();

  // ---- phase 2: Q time steps per thread, rows streamed once ----
  const int n0 = n_tile + (int)threadIdx.x * Q;
  if (n0 >= L) return;
  float acc[Q][N];
#pragma unroll
  for (int q = 0; q < Q; ++q)
#pragma unroll
    for (int p = 0; p < N; ++p) acc[q][p] = 0.0f;
  const float* base = vs + (int)threadIdx.x * (Q + 1) * 4;  // unit of row threadIdx.x * Q
#pragma unroll
  for (int i = 0; i < Q + HALO; ++i) {
    const float* row = base + (i + i / Q) * 4;
    float w[2][N4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      // half h of this row is used by the time steps q with offset o = DMIN + i - q in range and (i - q) parity h
      bool need = false;
#pragma unroll
      for (int q = 0; q < Q; ++q)
        if (i - q >= 0 && i - q <= HALO && ((i - q) & 1) == h) need = true;
      if (need) {
        const float4 t = *reinterpret_cast<const float4*>(row + h * UNITS * 4);
        w[h][0] = t.x; w[h][1] = t.y; w[h][2] = t.z; w[h][3] = t.w;
      }
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int d = i - q;  // o - DMIN
      if (d >= 0 && d <= HALO) {
#pragma unroll
        for (int e = 0; e < N; ++e) {
          const int j = R::jb(DMIN + d) + e;
          if (j >= 0 && j < K) acc[q][R::Geo::phase(j)] = fmaf(taps.g[j], w[d & 1][e], acc[q][R::Geo::phase(j)]);
        }
      }
    }
  }

  float* yo = y + (size_t)b * L * N + (size_t)n0 * N;
  const bool st_vec = (((size_t)L * N) % 4 == 0) && ((Q * N) % 4 == 0) && ((reinterpret_cast<uintptr_t>(y) & 15u) == 0);
  if (st_vec && n0 + Q <= L) {
    float flat[Q * N];
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
      for (int p = 0; p < N; ++p) flat[q * N + p] = acc[q][p];
    store_run<Q * N>(yo, flat, (reinterpret_cast<uintptr_t>(yo) & 31u) == 0);
  } else {
#pragma unroll
    for (int q = 0; q < Q; ++q)
      if (n0 + q < L) {
#pragma unroll
        for (int p = 0; p < N; ++p) yo[q * N + p] = acc[q][p];
      }
  }
}

// ------------------------------------------------------------------------------------------------------------
// synthesis, N = 3 (the reference's bank, vicreg_audio_params.py:40), cosine-modulated form in packed fp32 (FFMA2):
// same rows, same taps and the same accumulation order per output as k_pqmf_synthesis_small<3>, so the results are
// bit-identical to it.  That kernel is bound by the shared-memory / L1 data pipe (58 B of traffic per 4-byte output),
// not by instruction issue, so the point of the packing is the LAYOUT it allows: 24 instead of 32 bytes per row.
// The three taps that meet (time step q, row i), d = i - q, are j = 3d-1, 3d, 3d+1 on output phases 2, 1, 0 and read
// elements 0, 1, 2 of row half d & 1.  Two of them are adjacent in the row and in the tap table:
//   (y[q][2], y[q][1]) += (g[3d-1], g[3d]) * (half[0], half[1])
// and the third is paired with the neighbouring time step, which reads the OTHER half of the same row:
//   (y[q][0], y[q+1][0]) += (g[3d+1], g[3d-2]) * (half_d[2], half_{d-1}[2])          (q even)
// For even q the parity of d is the parity of the row index (Q is even), so a row is only ever read with one order
// of that pair and stores it in that order:
//   plane 0, 16 bytes per row: a0 a1 | b0 b1          plane 1, 8 bytes per row: a2 b2 (even rows) / b2 a2 (odd rows)
// (a = half 0, b = half 1): one 128-bit and one 64-bit shared load per row, every operand an aligned register pair.
// Taps outside [0, K) are zeros of the pair tables (their products add +0 to a finite accumulator).
// ------------------------------------------------------------------------------------------------------------
// tuning builds (tools/gpu_r4e.sh, one box, 1024 x 4 s): plane-1 padding every Q rows as in plane 0 (PAD1=0) / register
// cap through __launch_bounds__(128, MINB).  Measured: uncapped (75 registers, 6 CTAs per SM) 0.288 ms; MINB=7
// (70 registers, 7 CTAs per SM) 0.324 ms -- ptxas pays for the five registers with a serialised load schedule that
// costs far more than the seventh CTA brings; MINB=7 with PAD1=0 0.318; static instead of dynamic shared memory: no
// difference (not kept).
#ifndef IAS_SYN3_PAD1
#define IAS_SYN3_PAD1 1
#endif
#ifndef IAS_SYN3_MINB
#define IAS_SYN3_MINB 1
#endif
template <int K, int Q>
struct SynN3PSmem {
  static constexpr int ROWS = PQ_THREADS * Q + SynRows<3, K>::HALO;
  static constexpr int UNITS = ROWS + ROWS / Q + 1;         // plane 0: 16-byte units, row r at unit r + r/Q
  // plane 1: 8-byte units, row r at unit u1(r) = r + m + 8 (m / 8), m = r / (2Q)  (PAD1 = 0: r + r/Q as in plane 0)
  __host__ __device__ static constexpr int u1(int r) {
    return IAS_SYN3_PAD1 ? r + r / (2 * Q) + 8 * (r / (16 * Q)) : r + r / Q;
  }
  static constexpr int UNITS1 = u1(ROWS - 1) + 1;
  static constexpr size_t BYTES = (size_t)UNITS * 16 + (size_t)((UNITS1 + 1) / 2 * 2) * 8;
};

template <int K>
struct TapsSynN3P {
  static constexpr int ND = SynRows<3, K>::HALO + 1;
  // modulation pairs per band k: with c[k][r] as TapsSynSmall<3, K>::c and a_e = c[k][res(0, e)], b_e = c[k][res(1, e)]:
  // (a0, a1), (b0, b1), (a2, b2), (b2, a2)
  float2 cp[3 * 4];
  float2 ga[ND];    // (g[3d-1], g[3d])
  float2 gc[ND];    // (g[3d+1], g[3d-2])
};

template <int K, int Q, int MINB>
__global__ void __launch_bounds__(PQ_THREADS, MINB)
k_pqmf_synthesis_n3p(const float* __restrict__ z, float* __restrict__ y, int L, int tiles_per_row, TapsSynN3P<K> taps) {
  constexpr int N = 3;
  using R = SynRows<N, K>;
  constexpr int DMIN = R::DMIN, HALO = R::HALO;
  static_assert(R::jb(DMIN) == -1 && Q % 2 == 0, "pair tables assume taps 3d-1 .. 3d+1 and an even Q");
  static_assert(PQ_THREADS % (2 * Q) == 0, "phase-1 store offsets");
  constexpr int TILE_N = PQ_THREADS * Q;
  constexpr int ROWS = TILE_N + HALO;
  constexpr int PASSES = (ROWS + PQ_THREADS - 1) / PQ_THREADS;
  // phase-1 passes whose loads are issued before any arithmetic: all of them up to Q = 8 (measured: chunks of 4 passes
  // save 6 registers and cost 4 %: 0.327 against 0.314 ms -- the phase waits on those loads), chunks of 4 beyond
  constexpr int CH = Q <= 8 ? PASSES : 4;
  // Plane 0 (16-byte units): row r at unit r + r/Q.  Lane t of phase 2 reads rows Q t + i, i.e. units (Q + 1) t + const
  // with Q + 1 odd: the 8 lanes of a quarter warp hit 8 distinct 16-byte bank groups; phase 1 (lane = consecutive rows)
  // is conflict free as well.  Plane 1 (8-byte units) is served per half warp, 16 lanes over 16 8-byte bank pairs.  Its
  // pad may only change between the 16-row groups a half warp of phase 1 writes (m = r/16 for Q = 8), and phase 2 reads
  // rows 8 (t + a) + i' (a = i/8): sixteen lanes cover eight or nine consecutive m, both rows of an m land 8 units
  // apart, so the pads must differ mod 8 over eight consecutive m and by a multiple of 16 between m and m + 8:
  // u1(r) = r + m + 8 (m/8).  (ncu source pages: with the plane-0 padding r + r/8 every 64-bit store took 4 wavefronts
  // instead of 2, run r4y; with r + m alone the loads of every other row group did, run r4z.)  Phase-2 offsets are
  // compile-time constants relative to one base per row group a = i/Q.  Dynamic shared memory: Q = 16 needs 52 KB.
  constexpr int UNITS = SynN3PSmem<K, Q>::UNITS;
  using SM = SynN3PSmem<K, Q>;
  extern __shared__ __align__(16) float vs[];
  float* const plane1 = vs + UNITS * 4;

  int b, tile;
  row_and_tile(tiles_per_row, b, tile);
  const int n_tile = tile * TILE_N;
  const float* zb = z + (size_t)b * N * L;

  // ---- phase 1: modulate rows n_tile + DMIN + [0, ROWS) ----
  {
    const bool interior = n_tile + DMIN >= 0 && n_tile + DMIN + ROWS <= L;
    // interior tile (all but the first and last of a sound): one pointer per band, constant offsets, no bounds logic
    const float* p0 = zb + (n_tile + DMIN + (int)threadIdx.x);
    const float* p1 = p0 + L;
    const float* p2_ = p1 + L;
    // this thread's rows all have the parity of threadIdx.x (PQ_THREADS is even): order of the (a2, b2) pair
    const bool odd = (threadIdx.x & 1) != 0;
    P2 cx[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float2 e = taps.cp[k * 4 + 2], o = taps.cp[k * 4 + 3];
      cx[k] = odd ? p2(o.x, o.y) : p2(e.x, e.y);
    }
    const int unit0 = (int)threadIdx.x + (int)threadIdx.x / Q;
#pragma unroll
    for (int c0 = 0; c0 < PASSES; c0 += CH) {
      float zk[CH][N];
      if (interior) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int i = c0 + c;
          if (i < PASSES && ((i + 1) * PQ_THREADS <= ROWS || (int)threadIdx.x + i * PQ_THREADS < ROWS)) {
            zk[c][0] = __ldg(p0 + i * PQ_THREADS);
            zk[c][1] = __ldg(p1 + i * PQ_THREADS);
            zk[c][2] = __ldg(p2_ + i * PQ_THREADS);
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int row = (int)threadIdx.x + (c0 + c) * PQ_THREADS;
          const int m = n_tile + DMIN + row;
          const bool in = row < ROWS && m >= 0 && m < L;
#pragma unroll
          for (int k = 0; k < N; ++k) zk[c][k] = in ? __ldg(zb + (size_t)k * L + m) : 0.0f;
        }
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int i = c0 + c;
        if (i < PASSES && ((i + 1) * PQ_THREADS <= ROWS || (int)threadIdx.x + i * PQ_THREADS < ROWS)) {
          // the stored pairs straight from packed multiply-adds (same k order per value as the scalar kernel)
          P2 va = p2(0.0f, 0.0f), vb = va, vx = va;
#pragma unroll
          for (int k = 0; k < N; ++k) {
            const P2 zz = p2(zk[c][k], zk[c][k]);
            const float2 ca = taps.cp[k * 4 + 0], cb = taps.cp[k * 4 + 1];
            va = p2_fma(p2(ca.x, ca.y), zz, va);
            vb = p2_fma(p2(cb.x, cb.y), zz, vb);
            vx = p2_fma(cx[k], zz, vx);
          }
          // row / Q = threadIdx.x / Q + i * (PQ_THREADS / Q): the pass offsets are compile-time constants
          *reinterpret_cast<ulonglong2*>(vs + (unit0 + i * (PQ_THREADS + PQ_THREADS / Q)) * 4) = make_ulonglong2(va.v, vb.v);
          *reinterpret_cast<unsigned long long*>(plane1 + SM::u1((int)threadIdx.x + i * PQ_THREADS) * 2) = vx.v;
        }
      }
    }
  }
  __syncthreads();

  // ---- phase 2: Q time steps per thread, rows streamed once, two multiply-adds per instruction ----
  const int n0 = n_tile + (int)threadIdx.x * Q;
  if (n0 >= L) return;
  P2 acc21[Q];      // (y[q][2], y[q][1])
  P2 acc00[Q / 2];  // (y[2 q2][0], y[2 q2 + 1][0])
#pragma unroll
  for (int q = 0; q < Q; ++q) acc21[q] = p2(0.0f, 0.0f);
#pragma unroll
  for (int q = 0; q < Q / 2; ++q) acc00[q] = p2(0.0f, 0.0f);
  const int ubase = (int)threadIdx.x * (Q + 1);  // plane-0 unit of row threadIdx.x * Q
  // plane-1 unit of row Q t + i = u1(Q (t + a)) + (i - Q a), a = i/Q: the pad is constant over the Q rows of a group
  constexpr int GROUPS = (Q + HALO + Q - 1) / Q;
  int ubase1[GROUPS];
#pragma unroll
  for (int a = 0; a < GROUPS; ++a) ubase1[a] = SM::u1(((int)threadIdx.x + a) * Q);
#pragma unroll
  for (int i = 0; i < Q + HALO; ++i) {
    const int unit = ubase + i + i / Q;
    const int unit1 = ubase1[i / Q] + (i % Q);
    const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(vs + unit * 4);
    P2 half[2];
    half[0].v = t.x;
    half[1].v = t.y;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int d = i - q;
      if (d >= 0 && d <= HALO) {
        const float2 g = taps.ga[d];
        acc21[q] = p2_fma(half[d & 1], p2(g.x, g.y), acc21[q]);
      }
    }
    P2 x;
    x.v = *reinterpret_cast<const unsigned long long*>(plane1 + unit1 * 2);
#pragma unroll
    for (int q2 = 0; q2 < Q / 2; ++q2) {
      const int d = i - 2 * q2;  // step 2 q2 meets this row at d, step 2 q2 + 1 at d - 1 (tap 0 in gc[0].y)
      if (d >= 0 && d <= HALO) {
        const float2 g = taps.gc[d];
        acc00[q2] = p2_fma(x, p2(g.x, g.y), acc00[q2]);
      }
    }
  }

  float flat[Q * N];
#pragma unroll
  for (int q2 = 0; q2 < Q / 2; ++q2) {
    p2_unpack(acc00[q2], flat[(2 * q2) * N], flat[(2 * q2 + 1) * N]);
    p2_unpack(acc21[2 * q2], flat[(2 * q2) * N + 2], flat[(2 * q2) * N + 1]);
    p2_unpack(acc21[2 * q2 + 1], flat[(2 * q2 + 1) * N + 2], flat[(2 * q2 + 1) * N + 1]);
  }
  float* yo = y + (size_t)b * L * N + (size_t)n0 * N;
  const bool st_vec = (((size_t)L * N) % 4 == 0) && ((Q * N) % 4 == 0) && ((reinterpret_cast<uintptr_t>(y) & 15u) == 0);
  if (st_vec && n0 + Q <= L) {
    store_run<Q * N>(yo, flat, (reinterpret_cast<uintptr_t>(yo) & 31u) == 0);
  } else {
#pragma unroll
    for (int q = 0; q < Q; ++q)
      if (n0 + q < L) {
#pragma unroll
        for (int p = 0; p < N; ++p) yo[q * N + p] = flat[q * N + p];
      }
  }
}

// ------------------------------------------------------------------------------------------------------------
// synthesis, N = 4 (the default of pqmf.PQMF, pqmf.py:10), cosine-modulated form in packed fp32: the four taps that meet
// (time step s, row s + d) are j = 4d .. 4d+3 on output phases 3, 2, 1, 0 and read elements 0 .. 3 of row half d & 1.
// A row half is stored REVERSED (v3 v2 v1 v0), so that both operand pairs of a 128-bit shared load line up with
// accumulator pairs in output order: (y[s][0], y[s][1]) += (g[4d+3], g[4d+2]) * (v3, v2) and
// (y[s][2], y[s][3]) += (g[4d+1], g[4d]) * (v1, v0).  The kernel is bound by the L1 / shared-memory data pipe (ncu:
// 90 % of its wavefront peak with Q consecutive steps per thread, which read BOTH halves of every row: 46 128-bit
// loads per 32 outputs).  A step only ever reads the half whose parity is that of d, so a thread that owns Q steps of
// ONE parity (s = s0 + par + 2q) reads one half per row: 30 loads per 32 outputs.  Lanes 2p / 2p+1 own the even / odd
// steps of the same 2Q-step block and store 16-byte runs that interleave to contiguous memory.  Two 16-byte planes
// (one per half), row r at unit r + 2 (r / 2Q): the eight lanes of a quarter warp (four pairs, 2Q rows apart, plus
// the odd lanes' one-row shift) hit eight distinct bank groups, and so do the eight consecutive rows a quarter warp
// stores in phase 1 (bank simulation: 0 conflicts).  Same rows, taps and accumulation order per output as
// k_pqmf_synthesis_small<4>: bit-identical results.
// ------------------------------------------------------------------------------------------------------------
template <int K>
struct TapsSynN4P {
  static constexpr int ND = SynRows<4, K>::HALO + 1;
  // modulation pairs per band k, c[k][r] as TapsSynSmall<4, K>::c: (c[res(0,3)], c[res(0,2)]), (c[res(0,1)], c[res(0,0)]),
  // then the same for half 1
  float2 cp[4 * 4];
  float2 ga[ND];  // (g[4d+3], g[4d+2])
  float2 gb[ND];  // (g[4d+1], g[4d])
};

template <int K, int Q>
__global__ void __launch_bounds__(PQ_THREADS)
k_pqmf_synthesis_n4p(const float* __restrict__ z, float* __restrict__ y, int L, int tiles_per_row, TapsSynN4P<K> taps) {
  constexpr int N = 4;
  using R = SynRows<N, K>;
  constexpr int DMIN = R::DMIN, HALO = R::HALO;
  constexpr int G = 2 * Q;  // steps (rows) per lane pair
  static_assert(R::jb(DMIN) == 0, "pair tables assume taps 4d .. 4d+3");
  static_assert(PQ_THREADS % G == 0 && G % 8 == 0, "phase-1 store offsets / bank layout");
  constexpr int TILE_N = PQ_THREADS * Q;
  constexpr int ROWS = TILE_N + HALO;
  constexpr int PASSES = (ROWS + PQ_THREADS - 1) / PQ_THREADS;
  constexpr int UNITS = ROWS + 2 * (ROWS / G) + 2;
  __shared__ __align__(16) float vs[2 * UNITS * 4];

  int b, tile;
  row_and_tile(tiles_per_row, b, tile);
  const int n_tile = tile * TILE_N;
  const float* zb = z + (size_t)b * N * L;

  // ---- phase 1: modulate rows n_tile + DMIN + [0, ROWS) ----
  float zk[PASSES][N];
  if (n_tile + DMIN >= 0 && n_tile + DMIN + ROWS <= L) {
    // interior tile: one pointer per band, constant offsets, no bounds logic
    const float* p0 = zb + (n_tile + DMIN + (int)threadIdx.x);
    const float* p1 = p0 + L;
    const float* p2_ = p1 + L;
    const float* p3 = p2_ + L;
#pragma unroll
    for (int i = 0; i < PASSES; ++i)
      if ((i + 1) * PQ_THREADS <= ROWS || (int)threadIdx.x + i * PQ_THREADS < ROWS) {
        zk[i][0] = __ldg(p0 + i * PQ_THREADS);
        zk[i][1] = __ldg(p1 + i * PQ_THREADS);
        zk[i][2] = __ldg(p2_ + i * PQ_THREADS);
        zk[i][3] = __ldg(p3 + i * PQ_THREADS);
      }
  } else {
#pragma unroll
    for (int i = 0; i < PASSES; ++i) {
      const int row = (int)threadIdx.x + i * PQ_THREADS;
      const int m = n_tile + DMIN + row;
      const bool in = row < ROWS && m >= 0 && m < L;
#pragma unroll
      for (int k = 0; k < N; ++k) zk[i][k] = in ? __ldg(zb + (size_t)k * L + m) : 0.0f;
    }
  }
  {
    const int unit0 = (int)threadIdx.x + 2 * ((int)threadIdx.x / G);
#pragma unroll
    for (int i = 0; i < PASSES; ++i) {
      if ((i + 1) * PQ_THREADS <= ROWS || (int)threadIdx.x + i * PQ_THREADS < ROWS) {
        P2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = p2(0.0f, 0.0f);
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const P2 zz = p2(zk[i][k], zk[i][k]);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float2 c = taps.cp[k * 4 + u];
            v[u] = p2_fma(p2(c.x, c.y), zz, v[u]);
          }
        }
        float* dst = vs + (unit0 + i * (PQ_THREADS + 2 * (PQ_THREADS / G))) * 4;
        *reinterpret_cast<ulonglong2*>(dst) = make_ulonglong2(v[0].v, v[1].v);
        *reinterpret_cast<ulonglong2*>(dst + UNITS * 4) = make_ulonglong2(v[2].v, v[3].v);
      }
    }
  }
  __syncthreads();

  // ---- phase 2: Q time steps of one parity per thread (s = s0 + par + 2q), one row half streamed once ----
  const int pr = (int)threadIdx.x >> 1, par = (int)threadIdx.x & 1;
  const int s0 = n_tile + pr * G + par;  // this thread's first step; its row window starts at tile row pr * G + par
  if (s0 >= L) return;
  P2 acc01[Q], acc23[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) acc01[q] = acc23[q] = p2(0.0f, 0.0f);
  // tile row pr*G + par + j lives at unit (G + 2) pr + par + j + 2 ((par + j) / G): constant offsets, except that the
  // odd lane is one pad step ahead on the rows where par + j crosses a multiple of G
  const float* base = vs + ((G + 2) * pr + par) * 4;
  const int bump = par * 2 * 4;
#pragma unroll
  for (int j = 0; j < 2 * (Q - 1) + HALO + 1; ++j) {
    // step q meets this row at d = j - 2q: the parity of d, i.e. the row half, is that of j for every q
    const float* row = base + (j & 1) * UNITS * 4 + (j + 2 * (j / G)) * 4 + ((j % G == G - 1) ? bump : 0);
    const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(row);
    P2 lo, hi;  // (v3, v2) and (v1, v0)
    lo.v = t.x;
    hi.v = t.y;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int d = j - 2 * q;
      if (d >= 0 && d <= HALO) {
        const float2 ga = taps.ga[d], gb = taps.gb[d];
        acc01[q] = p2_fma(lo, p2(ga.x, ga.y), acc01[q]);
        acc23[q] = p2_fma(hi, p2(gb.x, gb.y), acc23[q]);
      }
    }
  }

  float* yo = y + (size_t)b * L * N + (size_t)s0 * N;
  const bool st_vec = (reinterpret_cast<uintptr_t>(y) & 15u) == 0;  // L * N and the run starts are multiples of 4 floats
#pragma unroll
  for (int q = 0; q < Q; ++q)
    if (s0 + 2 * q < L) {
      float f[4];
      p2_unpack(acc01[q], f[0], f[1]);
      p2_unpack(acc23[q], f[2], f[3]);
      if (st_vec) {
        *reinterpret_cast<float4*>(yo + 2 * q * N) = make_float4(f[0], f[1], f[2], f[3]);
      } else {
#pragma unroll
        for (int p = 0; p < N; ++p) yo[2 * q * N + p] = f[p];
      }
    }
}

__global__ void k_pqmf_synthesis_generic(const float* __restrict__ z, const float* __restrict__ G,
                                         float* __restrict__ y, int B, int L, int N, int K) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t Tout = (size_t)L * N;
  if (idx >= (size_t)B * Tout) return;
  const int b = (int)(idx / Tout);
  const long long t = (long long)(idx % Tout);
  const int pad = (K - 1) / 2;
  const float gain = (float)N;
  float acc = 0.0f;
  for (int k = 0; k < N; ++k) {
    const float* zr = z + ((size_t)b * N + k) * L;
    for (int j = 0; j < K; ++j) {
      const long long m = t + j - pad;
      if (m >= 0 && m < (long long)Tout && (m % N) == 0) acc = fmaf(G[k * K + j], zr[m / N] * gain, acc);
    }
  }
  y[idx] = acc;
}

template <int N, int K, int Q>
int launch_analysis(const float* x, const float* H_host, const float* proto_host, const float* mod_host,
                    const float* row_scale, const float* mean_host, const float* std_host, float* out, int B, int T,
                    int L, cudaStream_t st, const PoolReq* pr = nullptr) {
  constexpr int TILE_N = PQ_THREADS * Q;
  const int tiles = (L + TILE_N - 1) / TILE_N;
  const dim3 grid = row_tile_grid(B, tiles);
  PoolArgs pool{nullptr, 0, 0, 0, 0, 0, 0};
  if (pr) {
    const long long S = (long long)N * L;
    IAS_REQUIRE(pr->P > 1 && S * pr->P < (1LL << 31) && TILE_N + 1 <= S / pr->P, IAS_ERR_UNSUPPORTED,
                "ias_pqmf_analysis_pooled: P=%d needs bins wider than the %d-step CTA tile (N*L=%lld)", pr->P, TILE_N, S);
    const size_t need = (size_t)B * N * tiles * 2 * sizeof(float);
    IAS_REQUIRE(pr->workspace && pr->workspace_bytes >= need, IAS_ERR_INVALID,
                "ias_pqmf_analysis_pooled: workspace %zu < %zu bytes", pr->workspace_bytes, need);
    pool.acc = static_cast<unsigned long long*>(pr->workspace);  // B*P*8 <= need: P <= S / (TILE_N + 1)
    IAS_REQUIRE((reinterpret_cast<uintptr_t>(pr->workspace) & 7u) == 0, IAS_ERR_INVALID,
                "ias_pqmf_analysis_pooled: workspace must be 8-byte aligned");
    IAS_CUDA(cudaMemsetAsync(pool.acc, 0, (size_t)B * pr->P * sizeof(unsigned long long), st));
    pool.P = pr->P;
    pool.S = (int)S;
    magic_for((unsigned)S, pool.mS, pool.shS);
    magic_for((unsigned)pr->P, pool.mP, pool.shP);
  }
  BandNorm<N> norm;
  norm.on = (mean_host && std_host) ? 1 : 0;
  for (int k = 0; k < N; ++k) {
    norm.mean[k] = norm.on ? mean_host[k] : 0.0f;
    norm.std[k] = norm.on ? std_host[k] : 1.0f;
  }
  {
  ProfScope prof_(K_PQMF_ANALYSIS, st);
  if (proto_host && mod_host) {
    TapsCM<N, K> taps;
    for (int i = 0; i < K; ++i) taps.g[i] = proto_host[i];
    auto tap = [&](int j) { return (j >= 0 && j < K) ? proto_host[j] : 0.0f; };
    for (int m = 0; m < (K + 1) / 2; ++m) {
      taps.ge[m] = make_float2(tap(2 * m), tap(2 * m + 1));
      taps.go[m] = make_float2(tap(2 * m - 1), tap(2 * m));
    }
    if (N >= FOLD_MIN_N) {  // the folded form needs only N: cos((2k+1)(2m+1)pi/(4N)) / sqrt(2)
      for (int i = 0; i < N * 2 * N; ++i) taps.c[i] = 0.0f;
      for (int k = 0; k < N; ++k)
        for (int m = 0; m < N; ++m)
          taps.c[k * N + m] = (float)(cos((2.0 * k + 1.0) * (2.0 * m + 1.0) * 3.14159265358979323846 / (4.0 * N)) *
                                      0.70710678118654752440);
    } else {
      for (int i = 0; i < N * 2 * N; ++i) taps.c[i] = mod_host[i];
    }
    if (pr)
      k_pqmf_analysis<N, K, Q, TapsCM<N, K>, true><<<grid, PQ_THREADS, 0, st>>>(x, row_scale, out, T, L, tiles, taps, norm, pool);
    else
      k_pqmf_analysis<N, K, Q, TapsCM<N, K>, false><<<grid, PQ_THREADS, 0, st>>>(x, row_scale, out, T, L, tiles, taps, norm, pool);
  } else {
    Taps<N, K> taps;  // tap-major ([j][k]) so the bands of one tap are contiguous constants (vectorised LDCU)
    for (int k = 0; k < N; ++k)
      for (int j = 0; j < K; ++j) taps.h[j * N + k] = H_host[k * K + j];
    if (pr)
      k_pqmf_analysis<N, K, Q, Taps<N, K>, true><<<grid, PQ_THREADS, 0, st>>>(x, row_scale, out, T, L, tiles, taps, norm, pool);
    else
      k_pqmf_analysis<N, K, Q, Taps<N, K>, false><<<grid, PQ_THREADS, 0, st>>>(x, row_scale, out, T, L, tiles, taps, norm, pool);
  }
  }
  IAS_LAUNCH_CHECK("k_pqmf_analysis");
  if (pr) {
    ProfScope prof2_(K_POOL_FINALIZE, st);
    k_pool_finalize_fixed<<<(B * pr->P + 255) / 256, 256, 0, st>>>(pool.acc, pr->feat, B, N, L, pr->P);
    IAS_LAUNCH_CHECK("k_pool_finalize");
  }
  return IAS_OK;
}

template <int N, int K, int Q>
int launch_synthesis(const float* z, const float* G_host, float* y, int B, int L, cudaStream_t st) {
  Taps<N, K> taps;
  for (int i = 0; i < N * K; ++i) taps.h[i] = G_host[i];
  constexpr int TILE_N = PQ_THREADS * Q;
  const int tiles = (L + TILE_N - 1) / TILE_N;
  {
    ProfScope prof_(K_PQMF_SYNTHESIS, st);
    k_pqmf_synthesis<N, K, Q><<<row_tile_grid(B, tiles), PQ_THREADS, 0, st>>>(z, y, L, tiles, taps);
  }
  IAS_LAUNCH_CHECK("k_pqmf_synthesis");
  return IAS_OK;
}

template <int N, int K>
int launch_synthesis_cm(const float* z, const float* proto_host, float* y, int B, int L, cudaStream_t st) {
  TapsSynCM<N, K> taps;
  for (int i = 0; i < K; ++i) taps.g[i] = proto_host[i];
  for (int k = 0; k < N; ++k)
    for (int m = 0; m < N; ++m)
      taps.c[k * N + m] = (float)(cos((2.0 * k + 1.0) * (2.0 * m + 1.0) * 3.14159265358979323846 / (4.0 * N)) *
                                  0.70710678118654752440 * N);
  constexpr int TILE_N = PQ_THREADS - (SynthGeom<N, K>::omax() - SynthGeom<N, K>::omin());
  const int tiles = (L + TILE_N - 1) / TILE_N;
  {
    ProfScope prof_(K_PQMF_SYNTHESIS, st);
    k_pqmf_synthesis_cm<N, K><<<row_tile_grid(B, tiles), PQ_THREADS, 0, st>>>(z, y, L, tiles, taps);
  }
  IAS_LAUNCH_CHECK("k_pqmf_synthesis_cm");
  return IAS_OK;
}

template <int N, int K>
int launch_synthesis_cm2(const float* z, const float* proto_host, float* y, int B, int L, cudaStream_t st) {
  TapsSynCM2<N, K> taps;
  for (int k = 0; k < N; ++k)
    for (int m = 0; m < N; ++m)
      taps.c[k * N + m] = (float)(cos((2.0 * k + 1.0) * (2.0 * m + 1.0) * 3.14159265358979323846 / (4.0 * N)) *
                                  0.70710678118654752440 * N);
  auto g = [&](int j) { return (j >= 0 && j < K) ? proto_host[j] : 0.0f; };
  for (int rho = 0; rho < TapsSynCM2<N, K>::NR; ++rho)
    for (int k = 0; k < N; k += 2)
      taps.gp[rho * (N / 2) + k / 2] = make_float2(g(N * rho + N - 1 - k), g(N * rho + N - 2 - k));
  constexpr int TILE_N = (PQ_THREADS - (SynthGeom<N, K>::omax() - SynthGeom<N, K>::omin())) / 4 * 4;
  const int tiles = (L + TILE_N - 1) / TILE_N;
  {
    ProfScope prof_(K_PQMF_SYNTHESIS, st);
    k_pqmf_synthesis_cm2<N, K><<<row_tile_grid(B, tiles), PQ_THREADS, 0, st>>>(z, y, L, tiles, taps);
  }
  IAS_LAUNCH_CHECK("k_pqmf_synthesis_cm2");
  return IAS_OK;
}

template <int N, int K, int Q>
int launch_synthesis_small(const float* z, const float* proto_host, float* y, int B, int L, cudaStream_t st) {
  TapsSynSmall<N, K> taps;
  for (int i = 0; i < K; ++i) taps.g[i] = proto_host[i];
  for (int k = 0; k < N; ++k)
    for (int r = 0; r < 2 * N; ++r)
      taps.c[k * 2 * N + r] = (float)(N * cos((2.0 * k + 1.0) * (3.14159265358979323846 / (2.0 * N)) * (r - (K - 2) / 2.0) -
                                              ((k & 1) ? -1.0 : 1.0) * 3.14159265358979323846 / 4.0));
  constexpr int TILE_N = PQ_THREADS * Q;
  const int tiles = (L + TILE_N - 1) / TILE_N;
  {
    ProfScope prof_(K_PQMF_SYNTHESIS, st);
    k_pqmf_synthesis_small<N, K, Q><<<row_tile_grid(B, tiles), PQ_THREADS, 0, st>>>(z, y, L, tiles, taps);
  }
  IAS_LAUNCH_CHECK("k_pqmf_synthesis_small");
  return IAS_OK;
}

template <int K, int Q, int MINB>
int launch_synthesis_n3p(const float* z, const float* proto_host, float* y, int B, int L, cudaStream_t st) {
  constexpr int N = 3;
  TapsSynN3P<K> taps;
  using R = SynRows<N, K>;
  for (int k = 0; k < N; ++k) {
    auto c = [&](int h, int e) {
      const int r = R::res(h, e);
      return (float)(N * cos((2.0 * k + 1.0) * (3.14159265358979323846 / (2.0 * N)) * (r - (K - 2) / 2.0) -
                             ((k & 1) ? -1.0 : 1.0) * 3.14159265358979323846 / 4.0));
    };
    taps.cp[k * 4 + 0] = make_float2(c(0, 0), c(0, 1));
    taps.cp[k * 4 + 1] = make_float2(c(1, 0), c(1, 1));
    taps.cp[k * 4 + 2] = make_float2(c(0, 2), c(1, 2));
    taps.cp[k * 4 + 3] = make_float2(c(1, 2), c(0, 2));
  }
  auto g = [&](int j) { return (j >= 0 && j < K) ? proto_host[j] : 0.0f; };
  for (int d = 0; d < TapsSynN3P<K>::ND; ++d) {
    taps.ga[d] = make_float2(g(3 * d - 1), g(3 * d));
    taps.gc[d] = make_float2(g(3 * d + 1), g(3 * d - 2));
  }
  constexpr int TILE_N = PQ_THREADS * Q;
  const int tiles = (L + TILE_N - 1) / TILE_N;
  constexpr size_t smem = SynN3PSmem<K, Q>::BYTES;
  if (smem > 48 * 1024) {
    static unsigned long long seen = 0;
    if (ias_first_use_on_device(seen))
      IAS_CUDA(cudaFuncSetAttribute(k_pqmf_synthesis_n3p<K, Q, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  {
    ProfScope prof_(K_PQMF_SYNTHESIS, st);
    k_pqmf_synthesis_n3p<K, Q, MINB><<<row_tile_grid(B, tiles), PQ_THREADS, smem, st>>>(z, y, L, tiles, taps);
  }
  IAS_LAUNCH_CHECK("k_pqmf_synthesis_n3p");
  return IAS_OK;
}

template <int K, int Q>
int launch_synthesis_n4p(const float* z, const float* proto_host, float* y, int B, int L, cudaStream_t st) {
  constexpr int N = 4;
  TapsSynN4P<K> taps;
  using R = SynRows<N, K>;
  for (int k = 0; k < N; ++k) {
    auto c = [&](int h, int e) {
      const int r = R::res(h, e);
      return (float)(N * cos((2.0 * k + 1.0) * (3.14159265358979323846 / (2.0 * N)) * (r - (K - 2) / 2.0) -
                             ((k & 1) ? -1.0 : 1.0) * 3.14159265358979323846 / 4.0));
    };
    for (int h = 0; h < 2; ++h) {
      taps.cp[k * 4 + 2 * h + 0] = make_float2(c(h, 3), c(h, 2));
      taps.cp[k * 4 + 2 * h + 1] = make_float2(c(h, 1), c(h, 0));
    }
  }
  auto g = [&](int j) { return (j >= 0 && j < K) ? proto_host[j] : 0.0f; };
  for (int d = 0; d < TapsSynN4P<K>::ND; ++d) {
    taps.ga[d] = make_float2(g(4 * d + 3), g(4 * d + 2));
    taps.gb[d] = make_float2(g(4 * d + 1), g(4 * d));
  }
  constexpr int TILE_N = PQ_THREADS * Q;
  const int tiles = (L + TILE_N - 1) / TILE_N;
  {
    ProfScope prof_(K_PQMF_SYNTHESIS, st);
    k_pqmf_synthesis_n4p<K, Q><<<row_tile_grid(B, tiles), PQ_THREADS, 0, st>>>(z, y, L, tiles, taps);
  }
  IAS_LAUNCH_CHECK("k_pqmf_synthesis_n4p");
  return IAS_OK;
}

}  // namespace
}  // namespace ias

using namespace ias;

extern "C" int ias_pqmf_out_len(int T, int N, int K) {
  if (T <= 0 || N <= 0 || K <= 0) return -1;
  const int pad = (K - 1) / 2;
  const int span = T + 2 * pad - K;
  return span < 0 ? 0 : span / N + 1;
}

namespace {
// The cosine-modulated kernels rebuild (part of) the modulation from PQMF.__init__'s formula (pqmf.py:18-30: centre
// (taps-1)/2, phase (-1)^k pi/4) instead of reading it from the caller, so a caller-supplied factorisation is only
// honoured when it reproduces the filter that is actually passed: F[k][j] == proto[j] * cos(theta_kj +- phase_k) to
// fp32 rounding (and, for the small-N analysis kernel that does read mod_host, F[k][j] == proto[j] * mod[k][j % 2N]).
// Anything else -- e.g. the textbook taps/2 centring -- silently falls back to the direct form, which is valid for
// any taps.  The last accepted (F, proto, mod) per N and direction is cached, so steady-state calls cost a memcmp.
struct CmCheck {
  bool valid = false, ok = false;
  std::vector<float> F, proto, mod;
};
bool cm_design_matches(const float* F_host, const float* proto_host, const float* mod_host, int N, int K,
                       bool synthesis) {
  if (!F_host || !proto_host || N > 16 || N < 1) return false;
  static thread_local CmCheck cache[2][17];
  CmCheck& c = cache[synthesis ? 1 : 0][N];
  const size_t nF = (size_t)N * K, nM = mod_host ? (size_t)N * 2 * N : 0;
  if (c.valid && c.F.size() == nF && c.proto.size() == (size_t)K && c.mod.size() == nM &&
      memcmp(c.F.data(), F_host, nF * sizeof(float)) == 0 &&
      memcmp(c.proto.data(), proto_host, K * sizeof(float)) == 0 &&
      (nM == 0 || memcmp(c.mod.data(), mod_host, nM * sizeof(float)) == 0))
    return c.ok;
  const double PI = 3.14159265358979323846;
  float fmax = 0.0f;
  for (size_t i = 0; i < nF; ++i) fmax = fmaxf(fmax, fabsf(F_host[i]));
  const double tol = 2e-6 * (fmax > 0.0f ? fmax : 1.0f);
  bool ok = true;
  for (int k = 0; k < N && ok; ++k)
    for (int j = 0; j < K && ok; ++j) {
      const double sgn = ((j / (2 * N)) % 2 == 0) ? 1.0 : -1.0;  // the prototype carries the period-2N sign flip
      const double phase = ((k & 1) ? -1.0 : 1.0) * PI / 4.0;
      const double theta = (2.0 * k + 1.0) * (PI / (2.0 * N)) * (j - (K - 2) / 2.0);
      const double want = sgn * (double)proto_host[j] * cos(theta + (synthesis ? -phase : phase));
      if (fabs(want - (double)F_host[(size_t)k * K + j]) > tol) ok = false;
      if (ok && mod_host &&
          fabs((double)proto_host[j] * (double)mod_host[(size_t)k * 2 * N + j % (2 * N)] - (double)F_host[(size_t)k * K + j]) > tol)
        ok = false;
    }
  c.F.assign(F_host, F_host + nF);
  c.proto.assign(proto_host, proto_host + K);
  if (mod_host) c.mod.assign(mod_host, mod_host + nM); else c.mod.clear();
  c.valid = true;
  c.ok = ok;
  return ok;
}

int analysis_entry(const float* x, const float* H_dev, const float* H_host, const float* proto_host,
                   const float* mod_host, const float* row_scale, const float* mean_host, const float* std_host,
                   const float* norm_dev, float* out, int B, int T, int N, int K, ias_stream_t stream, const char* who,
                   const PoolReq* pr = nullptr) {
  IAS_REQUIRE(B > 0 && T > 0 && N > 0 && K > 0, IAS_ERR_INVALID, "%s: B=%d T=%d N=%d K=%d", who, B, T, N, K);
  IAS_REQUIRE(x && out, IAS_ERR_INVALID, "%s: NULL pointer", who);
  IAS_REQUIRE(H_dev || H_host, IAS_ERR_INVALID, "%s: no filter given", who);
  const int L = ias_pqmf_out_len(T, N, K);
  IAS_REQUIRE(L > 0, IAS_ERR_INVALID, "%s: empty output (T=%d K=%d)", who, T, K);
  cudaStream_t st = as_stream(stream);
  if (proto_host && mod_host && !cm_design_matches(H_host, proto_host, mod_host, N, K, false))
    proto_host = mod_host = nullptr;  // not the PQMF.__init__ design: direct form
  if (H_host && K == 63) {
#define IAS_PQ(NN, QQ) \
  case NN: return launch_analysis<NN, 63, QQ>(x, H_host, proto_host, mod_host, row_scale, mean_host, std_host, out, B, T, L, st, pr);
    switch (N) {
      IAS_PQ(2, 8)
      IAS_PQ(3, 4)  // measured (1024 x 4 s): Q=4 0.288 ms, Q=8 0.304 ms
      IAS_PQ(4, 8)
      IAS_PQ(8, 4)
      IAS_PQ(16, 2)
      default: break;
    }
#undef IAS_PQ
  }
  IAS_REQUIRE(!pr, IAS_ERR_UNSUPPORTED, "%s: N=%d K=%d has no specialised kernel (pooled epilogue unavailable)", who, N, K);
  IAS_REQUIRE(H_dev, IAS_ERR_UNSUPPORTED, "%s: N=%d K=%d has no specialised kernel and H_dev is NULL", who, N, K);
  IAS_REQUIRE(!(mean_host && std_host) || norm_dev, IAS_ERR_UNSUPPORTED,
              "%s: N=%d K=%d has no specialised kernel: the band normalisation needs norm_dev", who, N, K);
  const size_t total = (size_t)B * N * L;
  {
    ProfScope prof_(K_PQMF_ANALYSIS, st);
    k_pqmf_analysis_generic<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        x, H_dev, row_scale, (mean_host && std_host) ? norm_dev : nullptr, out, B, T, N, K, L);
  }
  IAS_LAUNCH_CHECK("k_pqmf_analysis_generic");
  return IAS_OK;
}
}  // namespace

extern "C" int ias_pqmf_analysis(const float* x, const float* H_dev, const float* H_host, const float* proto_host,
                                 const float* mod_host, const float* row_scale, float* out, int B, int T, int N,
                                 int K, ias_stream_t stream) {
  return analysis_entry(x, H_dev, H_host, proto_host, mod_host, row_scale, nullptr, nullptr, nullptr, out, B, T, N, K,
                        stream, "ias_pqmf_analysis");
}

extern "C" int ias_pqmf_analysis_image(const float* x, const float* H_dev, const float* H_host, const float* proto_host,
                                       const float* mod_host, const float* row_scale, const float* mean_host,
                                       const float* std_host, const float* norm_dev, float* out, int B, int T, int N,
                                       int K, ias_stream_t stream) {
  IAS_REQUIRE(mean_host && std_host, IAS_ERR_INVALID, "ias_pqmf_analysis_image: mean/std missing");
  for (int k = 0; k < N; ++k)
    IAS_REQUIRE(std_host[k] != 0.0f, IAS_ERR_INVALID, "ias_pqmf_analysis_image: std[%d] == 0", k);
  return analysis_entry(x, H_dev, H_host, proto_host, mod_host, row_scale, mean_host, std_host, norm_dev, out, B, T, N,
                        K, stream, "ias_pqmf_analysis_image");
}

extern "C" size_t ias_pqmf_pool_workspace_bytes(int B, int T, int N, int K) {
  const int L = ias_pqmf_out_len(T, N, K);
  if (B <= 0 || L <= 0) return 0;
  // tiles of the smallest CTA tile any specialised kernel uses (PQ_THREADS * 2 steps): an upper bound
  const size_t tiles = ((size_t)L + 2 * PQ_THREADS - 1) / (2 * PQ_THREADS);
  return (size_t)B * N * tiles * 2 * sizeof(float);
}

extern "C" int ias_pqmf_analysis_pooled(const float* x, const float* H_dev, const float* H_host, const float* proto_host,
                                        const float* mod_host, const float* row_scale, float* out, float* feat, int P,
                                        void* workspace, size_t workspace_bytes, int B, int T, int N, int K,
                                        ias_stream_t stream) {
  IAS_REQUIRE(feat && workspace, IAS_ERR_INVALID, "ias_pqmf_analysis_pooled: NULL feat / workspace");
  PoolReq pr{feat, P, workspace, workspace_bytes};
  return analysis_entry(x, H_dev, H_host, proto_host, mod_host, row_scale, nullptr, nullptr, nullptr, out, B, T, N, K,
                        stream, "ias_pqmf_analysis_pooled", &pr);
}

extern "C" int ias_pqmf_synthesis(const float* z, const float* G_dev, const float* G_host, const float* proto_host,
                                  float* y, int B, int L, int N, int K, ias_stream_t stream) {
  IAS_REQUIRE(B > 0 && L > 0 && N > 0 && K > 0, IAS_ERR_INVALID, "ias_pqmf_synthesis: B=%d L=%d N=%d K=%d", B, L, N, K);
  IAS_REQUIRE(z && y, IAS_ERR_INVALID, "ias_pqmf_synthesis: NULL pointer");
  IAS_REQUIRE(G_dev || G_host, IAS_ERR_INVALID, "ias_pqmf_synthesis: no filter given");
  cudaStream_t st = as_stream(stream);
  int q_env = 0;  // IAS_PQMF_SYNTH_Q: tuning override of the time steps per thread
  if (const char* e = getenv("IAS_PQMF_SYNTH_Q")) q_env = atoi(e);
  if (proto_host && !cm_design_matches(G_host, proto_host, nullptr, N, K, true)) proto_host = nullptr;
  if (proto_host && K == 63) {  // G is the designed filter: cosine-modulated form
    const char* cm2 = getenv("IAS_PQMF_SYNTH_CM2");  // tuning switch: 0 = one step per thread (bit-identical results)
    if (!cm2 || atoi(cm2) != 0) {
      if (N == 16) return launch_synthesis_cm2<16, 63>(z, proto_host, y, B, L, st);
      if (N == 8) return launch_synthesis_cm2<8, 63>(z, proto_host, y, B, L, st);
    }
    if (N == 16) return launch_synthesis_cm<16, 63>(z, proto_host, y, B, L, st);
    if (N == 8) return launch_synthesis_cm<8, 63>(z, proto_host, y, B, L, st);
    if (!getenv("IAS_PQMF_SYNTH_DIRECT")) {  // tuning switch: direct form for N <= 4
      if (N == 4) {
        const char* pk = getenv("IAS_PQMF_SYNTH_PACKED");  // tuning switch: 0 = scalar FIR phase (bit-identical results)
        if (!pk || atoi(pk) != 0) {
          if (q_env == 4) return launch_synthesis_n4p<63, 4>(z, proto_host, y, B, L, st);
          return launch_synthesis_n4p<63, 8>(z, proto_host, y, B, L, st);
        }
        if (q_env == 8) return launch_synthesis_small<4, 63, 8>(z, proto_host, y, B, L, st);
        return launch_synthesis_small<4, 63, 4>(z, proto_host, y, B, L, st);  // measured: Q=4 0.411 ms, Q=8 0.471, direct 0.643
      }
      if (N == 3) {
        const char* pk = getenv("IAS_PQMF_SYNTH_PACKED");  // tuning switch: 0 = scalar FIR phase (bit-identical results)
        if (!pk || atoi(pk) != 0) {
          if (q_env == 4) return launch_synthesis_n3p<63, 4, 1>(z, proto_host, y, B, L, st);
          if (q_env == 16) return launch_synthesis_n3p<63, 16, 1>(z, proto_host, y, B, L, st);  // 94 registers, 52 KB: 0.373 ms
          if (q_env == 88) return launch_synthesis_n3p<63, 8, 8>(z, proto_host, y, B, L, st);  // 63 registers, 8 CTAs per SM: 0.333 ms
          return launch_synthesis_n3p<63, 8, IAS_SYN3_MINB>(z, proto_host, y, B, L, st);  // measured: 0.272 ms (scalar kernel 0.345)
        }
        if (q_env == 4) return launch_synthesis_small<3, 63, 4>(z, proto_host, y, B, L, st);
        return launch_synthesis_small<3, 63, 8>(z, proto_host, y, B, L, st);  // measured: Q=8 0.345 ms, Q=4 0.485, direct 0.494
      }
      if (N == 2) return launch_synthesis_small<2, 63, 4>(z, proto_host, y, B, L, st);
    }
  }
  if (G_host && K == 63) {
    int q = 0;  // IAS_PQMF_SYNTH_Q: tuning override of the time steps per thread
    if (const char* e = getenv("IAS_PQMF_SYNTH_Q")) q = atoi(e);
    switch (N) {
      case 2: return launch_synthesis<2, 63, 8>(z, G_host, y, B, L, st);
      case 3:
        if (q == 16) return launch_synthesis<3, 63, 16>(z, G_host, y, B, L, st);
        if (q == 8) return launch_synthesis<3, 63, 8>(z, G_host, y, B, L, st);
        return launch_synthesis<3, 63, 4>(z, G_host, y, B, L, st);  // measured: Q=4 0.495 ms, Q=8 0.527, Q=16 0.668
      case 4: return launch_synthesis<4, 63, 8>(z, G_host, y, B, L, st);
      case 8: return launch_synthesis<8, 63, 4>(z, G_host, y, B, L, st);
      case 16:
        if (q == 4) return launch_synthesis<16, 63, 4>(z, G_host, y, B, L, st);
        return launch_synthesis<16, 63, 2>(z, G_host, y, B, L, st);
      default: break;
    }
  }
  IAS_REQUIRE(G_dev, IAS_ERR_UNSUPPORTED, "ias_pqmf_synthesis: N=%d K=%d has no specialised kernel and G_dev is NULL",
              N, K);
  const size_t total = (size_t)B * L * N;
  {
    ProfScope prof_(K_PQMF_SYNTHESIS, st);
    k_pqmf_synthesis_generic<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(z, G_dev, y, B, L, N, K);
  }
  IAS_LAUNCH_CHECK("k_pqmf_synthesis_generic");
  return IAS_OK;
}
