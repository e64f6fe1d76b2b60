// NOT COMPILED.  Two rejected variants of the audio kernel, kept for the record (profiles/sweep_r3c.log,
// sweep_r3d.log, sweep_r3e.log, sweep_r3f.log; ncu capture summarised in profiles/r3_audio_experiments.md).  They were
// members of csrc/voice.cu (same helpers: P2, pitch_pair, widen_pos, SpVoice, cos_arg_p2, squaresaw_core_p2 ...) and
// produced bit-identical audio.
//   k_voice_audio_sx  software-pipelined kernel with the phase increments in shared memory instead of registers
//                     (80 / 64 registers -> 6 / 8 CTAs per SM): 0.72 / 0.76 / 0.80 / 0.85 ms at 4 / 5 / 6 / 8 CTAs per SM
//   k_voice_audio_s2  two tiles of skew: scan of tile t+1 interleaved with pass 2 of tile t and pass 1 of tile t+2, no
//                     serial section between tiles: 0.769 ms (4 CTAs) against 0.708 of k_voice_audio_sp
// (s2's merged_passes_xs additionally took `bool SCAN`, `double* sv1, *sv2` and ran one shuffle level of the warp scan
// after each of the first five sample pairs.)
// ------------------------------------------------------------------------------------------------------------
// k_voice_audio_sx: k_voice_audio_sp with the phase increments of the tile in flight parked in shared memory
// ------------------------------------------------------------------------------------------------------------
// The pipelined kernel keeps x1, x2 (32 registers), the source coordinates (16) and the noise / output samples (16) of
// a thread alive through the merged block, which leaves ptxas ~20 registers of a 128-register budget for the chains
// it interleaves.  Here the increments live in shared memory ([vco][k/4][thread] float4: a thread only ever touches its
// own 16-byte slots, so there is no bank conflict and no synchronisation), the noise is loaded four samples at a time
// where it is used (L1-prefetched a tile ahead) and every four output samples are stored as soon as they exist.
__device__ __forceinline__ float4 ld_keep(const float4* p) {  // a load ptxas may not sink below the barrier
  float4 v;
  asm volatile("ld.global.ca.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

template <int NT, int SPT, bool CLAMP, bool DBG>
__device__ __forceinline__ void merged_passes_xs(float4 (*xs)[SPT / 4][NT], int tid, const float* nzp, float* outp,
                                                 bool live, double& acc1, double& acc2, double& tot1, double& tot2,
                                                 float& lpeak, const SpVoice& V, float ft0, float fj, float fj1,
                                                 const float4 r1, const float4 r2, const float4 r3, float ft0n,
                                                 float fjn, float fj1n, const float4 r0n, const float4 r1n, float* dbg1,
                                                 float* dbg2, int t0, int T) {
#pragma unroll
  for (int q = 0; q < SPT / 4; ++q) {
    const float4 X1 = xs[0][q][tid], X2 = xs[1][q][tid];
    float4 N4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) N4 = __ldg(reinterpret_cast<const float4*>(nzp) + q);
    const float xa1[4] = {X1.x, X1.y, X1.z, X1.w}, xa2[4] = {X2.x, X2.y, X2.z, X2.w};
    const float na[4] = {N4.x, N4.y, N4.z, N4.w};
    float ya[4], xn1[4], xn2[4];
#pragma unroll
    for (int h = 0; h < 4; h += 2) {
      const int k = 4 * q + h;
      // ---- tile t, pass 2 (source coordinates recomputed: same fma as pass 1) ----
      const P2 sp = p2_fma(p2b(V.scale), p2_add(p2b(ft0), p2((float)k, (float)(k + 1))), p2b(0.0f));
      const P2 u = p2_sub(sp, p2b(fj));
      const P2 um = p2_sub(sp, p2b(fj1));
      const P2 r = p2(fmaxf(p2lo(um), 0.0f), fmaxf(p2hi(um), 0.0f));
      acc1 += widen_pos(xa1[h]);
      const float a10 = (float)acc1;
      acc1 += widen_pos(xa1[h + 1]);
      const float a11 = (float)acc1;
      acc2 += widen_pos(xa2[h]);
      const float a20 = (float)acc2;
      acc2 += widen_pos(xa2[h + 1]);
      const float a21 = (float)acc2;
      const P2 arg1 = p2_add(p2(a10, a11), p2b(V.phase1));
      const P2 arg2 = p2_add(p2(a20, a21), p2b(V.phase2));
      const P2 g1 = p2_fma(r, p2b(r2.x), p2_fma(u, p2b(r1.w), p2b(r1.z)));
      const P2 g2 = p2_fma(r, p2b(r2.w), p2_fma(u, p2b(r2.z), p2b(r2.y)));
      const P2 g3 = p2_fma(r, p2b(r3.z), p2_fma(u, p2b(r3.y), p2b(r3.x)));
      const P2 yy = p2_fma(cos_arg_p2(arg1), g1,
                           p2_fma(squaresaw_core_p2(arg2, V.pk, V.shape), g2, p2_mul(p2(na[h], na[h + 1]), g3)));
      ya[h] = p2lo(yy);
      ya[h + 1] = p2hi(yy);
      lpeak = fmaxf(lpeak, fmaxf(fabsf(ya[h]), fabsf(ya[h + 1])));
      if (DBG) {
        if ((t0 + k) < T) {
          dbg1[t0 + k] = p2lo(arg1);
          dbg2[t0 + k] = p2lo(arg2);
        }
        if ((t0 + k + 1) < T) {
          dbg1[t0 + k + 1] = p2hi(arg1);
          dbg2[t0 + k + 1] = p2hi(arg2);
        }
      }
      // ---- tile t+1, pass 1 ----
      P2 i1, i2, spn;
      pitch_pair<CLAMP>(k, ft0n, V.scale, fjn, fj1n, r0n, r1n, V.midi1, V.depth1, V.midi2, V.depth2, V.sr, V.rsr, i1, i2,
                        spn);
      xn1[h] = p2lo(i1);
      xn1[h + 1] = p2hi(i1);
      xn2[h] = p2lo(i2);
      xn2[h + 1] = p2hi(i2);
      tot1 += widen_pos(xn1[h]);
      tot1 += widen_pos(xn1[h + 1]);
      tot2 += widen_pos(xn2[h]);
      tot2 += widen_pos(xn2[h + 1]);
    }
    xs[0][q][tid] = make_float4(xn1[0], xn1[1], xn1[2], xn1[3]);
    xs[1][q][tid] = make_float4(xn2[0], xn2[1], xn2[2], xn2[3]);
    if (live) reinterpret_cast<float4*>(outp)[q] = make_float4(ya[0], ya[1], ya[2], ya[3]);
  }
}

// Requires T % SPT == 0 and 16-byte aligned rows, as k_voice_audio_sp.
template <int NT, int SPT, int MINB, bool DBG>
__global__ void __launch_bounds__(NT, MINB) k_voice_audio_sx(AudioArgs A) {
  constexpr int TILE = NT * SPT;
  constexpr int NW = NT / 32;
  __shared__ float4 xs[2][SPT / 4][NT];  // phase increments of the tile in flight: [vco][quad][thread]
  __shared__ double s_wsum[2][2][NW];    // [buffer][vco][warp]
  __shared__ float s_peak[NW];
  __shared__ int s_slot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = A.T, C = A.C;
  SpVoice V;
  V.scale = A.scale;
  V.sr = A.sr;
  V.rsr = A.rsr;
  const float fTm1 = (float)(T - 1);

  for (;;) {
    if (tid == 0) s_slot = atomicAdd(A.counter, 1);
    __syncthreads();
    const int slot = s_slot;
    if (slot >= A.B) break;
    const int b = A.order[slot];
    const float* vc = A.vconst + (size_t)b * VC_COUNT;
    V.midi1 = vc[VC_MIDI1]; V.depth1 = vc[VC_DEPTH1]; V.phase1 = vc[VC_PHASE1];
    V.midi2 = vc[VC_MIDI2]; V.depth2 = vc[VC_DEPTH2]; V.phase2 = vc[VC_PHASE2];
    V.pk = vc[VC_PK]; V.shape = vc[VC_SHAPE];
    const bool noclamp = vc[VC_NOCLAMP] != 0.0f;
    const float4* rec = A.rec + (size_t)b * C * (REC_FLOATS / 4);
    const float* nz = A.noise + (size_t)(b % A.noise_rows) * T;
    float* out = A.audio + (size_t)b * T;
    float* dbg1 = DBG ? A.phase_dbg + ((size_t)b * 2 + 0) * T : nullptr;
    float* dbg2 = DBG ? A.phase_dbg + ((size_t)b * 2 + 1) * T : nullptr;
    const int ntiles = A.ntiles[b];

    double carry1 = 0, carry2 = 0;
    float tpeak = 0.0f;
    // ---- prologue: pass 1 of tile 0 ----
    float ft0 = (float)(tid * SPT);
    int j = min((int)mul(V.scale, fminf(ft0, fTm1)), C - 1);
    float fj = (float)j, fj1 = add(fj, 1.0f);
    const float4* rj = rec + (size_t)j * 4;
    double tot1 = 0, tot2 = 0;
    if (ntiles > 0) {
      const float4 r0 = __ldg(rj + 0);
      const float4 r1 = __ldg(rj + 1);
#pragma unroll
      for (int q = 0; q < SPT / 4; ++q) {
        float xn1[4], xn2[4];
#pragma unroll
        for (int h = 0; h < 4; h += 2) {
          P2 i1, i2, sp;
          if (noclamp)
            pitch_pair<false>(4 * q + h, ft0, V.scale, fj, fj1, r0, r1, V.midi1, V.depth1, V.midi2, V.depth2, V.sr, V.rsr,
                              i1, i2, sp);
          else
            pitch_pair<true>(4 * q + h, ft0, V.scale, fj, fj1, r0, r1, V.midi1, V.depth1, V.midi2, V.depth2, V.sr, V.rsr,
                             i1, i2, sp);
          xn1[h] = p2lo(i1); xn1[h + 1] = p2hi(i1);
          xn2[h] = p2lo(i2); xn2[h + 1] = p2hi(i2);
          tot1 += widen_pos(xn1[h]);
          tot1 += widen_pos(xn1[h + 1]);
          tot2 += widen_pos(xn2[h]);
          tot2 += widen_pos(xn2[h + 1]);
        }
        xs[0][q][tid] = make_float4(xn1[0], xn1[1], xn1[2], xn1[3]);
        xs[1][q][tid] = make_float4(xn2[0], xn2[1], xn2[2], xn2[3]);
      }
    }
    for (int tile = 0; tile < ntiles; ++tile) {
      const int t0 = tile * TILE + tid * SPT;
      const int buf = tile & 1;
      // records of tile t (gains) and t+1 (pitch points): issued before the scan so their latency hides under it
      const float4 r1 = ld_keep(rj + 1);
      const float4 r2 = ld_keep(rj + 2);
      const float4 r3 = ld_keep(rj + 3);
      const float ft0n = add(ft0, (float)TILE);
      const int jn = min((int)mul(V.scale, fminf(ft0n, fTm1)), C - 1);
      const float fjn = (float)jn, fj1n = add(fjn, 1.0f);
      const float4* rjn = rec + (size_t)jn * 4;
      const float4 r0n = ld_keep(rjn + 0);
      const float4 r1n = ld_keep(rjn + 1);
      if (tile + 2 < ntiles) {
        const int jnn = min((int)mul(V.scale, fminf(add(ft0n, (float)TILE), fTm1)), C - 1);
        prefetch_l1(rec + (size_t)jnn * 4);
      }
      if (t0 + TILE < T) prefetch_l1(nz + t0 + TILE);
      // ---- block scan of tile t ----
      const double inc1 = warp_incl_scan(tot1, lane);
      const double inc2 = warp_incl_scan(tot2, lane);
      if (lane == 31) {
        s_wsum[buf][0][warp] = inc1;
        s_wsum[buf][1][warp] = inc2;
      }
      __syncthreads();
      double acc1 = carry1 + (inc1 - tot1), acc2 = carry2 + (inc2 - tot2);
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        const double w1 = s_wsum[buf][0][w], w2 = s_wsum[buf][1][w];
        if (w < warp) {
          acc1 += w1;
          acc2 += w2;
        }
        carry1 += w1;
        carry2 += w2;
      }
      // ---- pass 2 of tile t merged with pass 1 of tile t+1 ----
      float lpeak = 0.0f;
      tot1 = 0;
      tot2 = 0;
      const bool live = t0 < T;
      if (noclamp)
        merged_passes_xs<NT, SPT, false, DBG>(xs, tid, nz + t0, out + t0, live, acc1, acc2, tot1, tot2, lpeak, V, ft0, fj,
                                              fj1, r1, r2, r3, ft0n, fjn, fj1n, r0n, r1n, dbg1, dbg2, t0, T);
      else
        merged_passes_xs<NT, SPT, true, DBG>(xs, tid, nz + t0, out + t0, live, acc1, acc2, tot1, tot2, lpeak, V, ft0, fj,
                                             fj1, r1, r2, r3, ft0n, fjn, fj1n, r0n, r1n, dbg1, dbg2, t0, T);
      if (live) tpeak = fmaxf(tpeak, lpeak);
      ft0 = ft0n; fj = fjn; fj1 = fj1n; rj = rjn;
    }

    // ---- silent tail ----
    if ((long long)ntiles * TILE < T) {
      float4* o4 = reinterpret_cast<float4*>(out);
      for (int i = ntiles * (TILE / 4) + tid; i < T / 4; i += NT) o4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // ---- per-voice peak, normalize_if_clipping (as in k_voice_audio) ----
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) tpeak = fmaxf(tpeak, __shfl_xor_sync(0xffffffffu, tpeak, d));
    if (lane == 0) s_peak[warp] = tpeak;
    __syncthreads();  // also orders this CTA's global stores before the re-read below
    float pkv = s_peak[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) pkv = fmaxf(pkv, s_peak[w]);
    if (tid == 0 && A.peak) A.peak[b] = pkv;
    if (A.normalize && pkv > 1.0f) {
      const float rp = vm::div(1.0f, pkv);
      const int live = min(T, ntiles * TILE);
      float4* o4 = reinterpret_cast<float4*>(out);
      const int n4 = live / 4;
      constexpr int U = 4;
      for (int base = 0; base < n4; base += U * NT) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = base + u * NT + tid;
          if (i < n4) v[u] = o4[i];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = base + u * NT + tid;
          if (i < n4) {
            v[u].x = div_const(v[u].x, pkv, rp); v[u].y = div_const(v[u].y, pkv, rp);
            v[u].z = div_const(v[u].z, pkv, rp); v[u].w = div_const(v[u].w, pkv, rp);
            o4[i] = v[u];
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// k_voice_audio_s2: two tiles of skew -- no serial section between the tiles
// ------------------------------------------------------------------------------------------------------------
// Occupancy scaling of the kernels above (profiles/sweep_r3e.log; every voice rendered in full, B = 3552): 1 / 2 / 3 /
// 4 resident CTAs per SM take 5.10 / 3.25 / 2.81 / 2.55 ms, i.e. a lone warp needs ~4800 cycles per tile of which the
// FMA pipe is busy 1313; ~3000 of them are the serial part between two tiles -- the fp64 warp scan of the thread
// totals (five dependent shuffle + DADD levels), the CTA barrier, the carry, and the record / noise loads -- during
// which the warp issues next to nothing, and four warps per scheduler do not cover it.
// Here the block of one iteration evaluates pass 2 of tile t, pass 1 of tile t+2 AND the scan of tile t+1 (whose
// thread totals the previous iteration produced): three independent instruction streams in one basic block, so the
// scan's latency chain and the loads hide under the arithmetic.  What is left between two blocks is the barrier
// that publishes the warp totals and a four-deep carry sum.  The increments of the two tiles in flight live in
// shared memory (2 x 16 KB per CTA, [buffer][vco][k/4][thread] float4, a thread touches only its own slots).
template <int NT, int SPT, bool CLAMP, bool DBG>
__device__ __forceinline__ void s2_pass1_tile(float4 (*xs)[SPT / 4][NT], int tid, const SpVoice& V, float ft0, float fj,
                                              float fj1, const float4 r0, const float4 r1, double& tot1, double& tot2) {
#pragma unroll
  for (int q = 0; q < SPT / 4; ++q) {
    float xn1[4], xn2[4];
#pragma unroll
    for (int h = 0; h < 4; h += 2) {
      P2 i1, i2, sp;
      pitch_pair<CLAMP>(4 * q + h, ft0, V.scale, fj, fj1, r0, r1, V.midi1, V.depth1, V.midi2, V.depth2, V.sr, V.rsr, i1,
                        i2, sp);
      xn1[h] = p2lo(i1); xn1[h + 1] = p2hi(i1);
      xn2[h] = p2lo(i2); xn2[h + 1] = p2hi(i2);
      tot1 += widen_pos(xn1[h]);
      tot1 += widen_pos(xn1[h + 1]);
      tot2 += widen_pos(xn2[h]);
      tot2 += widen_pos(xn2[h + 1]);
    }
    xs[0][q][tid] = make_float4(xn1[0], xn1[1], xn1[2], xn1[3]);
    xs[1][q][tid] = make_float4(xn2[0], xn2[1], xn2[2], xn2[3]);
  }
}

struct S2Tile {  // where a thread's SPT samples of one tile sit on the control grid
  float ft0, fj, fj1;
  const float4* rj;
};
__device__ __forceinline__ S2Tile s2_tile(float ft0, float scale, float fTm1, int C, const float4* rec) {
  S2Tile t;
  t.ft0 = ft0;
  const int j = min((int)mul(scale, fminf(ft0, fTm1)), C - 1);
  t.fj = (float)j;
  t.fj1 = add(t.fj, 1.0f);
  t.rj = rec + (size_t)j * 4;
  return t;
}

template <int NT, int SPT, bool CLAMP, bool DBG>
__device__ __forceinline__ float s2_render_voice(const AudioArgs& A, const SpVoice& V, int b, int ntiles,
                                                 float4 (*xs)[2][SPT / 4][NT], double (*s_wsum)[2][NT / 32]) {
  constexpr int TILE = NT * SPT;
  constexpr int NW = NT / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = A.T, C = A.C;
  const float fTm1 = (float)(T - 1);
  const float4* rec = A.rec + (size_t)b * C * (REC_FLOATS / 4);
  const float* nz = A.noise + (size_t)(b % A.noise_rows) * T;
  float* out = A.audio + (size_t)b * T;
  float* dbg1 = DBG ? A.phase_dbg + ((size_t)b * 2 + 0) * T : nullptr;
  float* dbg2 = DBG ? A.phase_dbg + ((size_t)b * 2 + 1) * T : nullptr;

  // ---- prologue: pass 1 of tiles 0 and 1, scan + carry of tile 0 ----
  S2Tile cur = s2_tile((float)(tid * SPT), V.scale, fTm1, C, rec);
  S2Tile nxt = s2_tile(add(cur.ft0, (float)TILE), V.scale, fTm1, C, rec);
  double tot1 = 0, tot2 = 0, totn1 = 0, totn2 = 0;
  s2_pass1_tile<NT, SPT, CLAMP, DBG>(xs[0], tid, V, cur.ft0, cur.fj, cur.fj1, __ldg(cur.rj), __ldg(cur.rj + 1), tot1, tot2);
  s2_pass1_tile<NT, SPT, CLAMP, DBG>(xs[1], tid, V, nxt.ft0, nxt.fj, nxt.fj1, __ldg(nxt.rj), __ldg(nxt.rj + 1), totn1,
                                     totn2);
  double carry1 = 0, carry2 = 0, acc1, acc2;
  {
    const double inc1 = warp_incl_scan(tot1, lane), inc2 = warp_incl_scan(tot2, lane);
    if (lane == 31) {
      s_wsum[1][0][warp] = inc1;  // parity of "tile -1 + 1": the loop's first write goes to buffer (0 + 1) & 1 = 1 ...
      s_wsum[1][1][warp] = inc2;
    }
    __syncthreads();
    acc1 = inc1 - tot1;
    acc2 = inc2 - tot2;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const double w1 = s_wsum[1][0][w], w2 = s_wsum[1][1][w];
      if (w < warp) {
        acc1 += w1;
        acc2 += w2;
      }
      carry1 += w1;
      carry2 += w2;
    }
    __syncthreads();  // ... so buffer 1 must be free again before iteration 0 writes tile 1's totals into it
  }
  float tpeak = 0.0f;
  for (int tile = 0; tile < ntiles; ++tile) {
    const int t0 = tile * TILE + tid * SPT;
    const int pb = (tile + 1) & 1;  // s_wsum buffer of tile t+1
    float4(*xc)[SPT / 4][NT] = xs[tile & 1];
    // records: gains of tile t, pitch points of tile t+2
    const float4 r1 = __ldg(cur.rj + 1);
    const float4 r2 = __ldg(cur.rj + 2);
    const float4 r3 = __ldg(cur.rj + 3);
    const S2Tile nn = s2_tile(add(nxt.ft0, (float)TILE), V.scale, fTm1, C, rec);
    const float4 r0n = __ldg(nn.rj + 0);
    const float4 r1n = __ldg(nn.rj + 1);
    if (tile + 3 < ntiles) {
      const int jp = min((int)mul(V.scale, fminf(add(nn.ft0, (float)TILE), fTm1)), C - 1);
      prefetch_l1(rec + (size_t)jp * 4);
    }
    if (t0 + TILE < T) prefetch_l1(nz + t0 + TILE);
    // ---- scan of tile t+1 (independent of everything below until the barrier) ----
    const double inc1 = warp_incl_scan(totn1, lane), inc2 = warp_incl_scan(totn2, lane);
    if (lane == 31) {
      s_wsum[pb][0][warp] = inc1;
      s_wsum[pb][1][warp] = inc2;
    }
    const double ex1 = inc1 - totn1, ex2 = inc2 - totn2;
    // ---- pass 2 of tile t merged with pass 1 of tile t+2 ----
    float lpeak = 0.0f;
    totn1 = 0;
    totn2 = 0;
    const bool live = t0 < T;
    merged_passes_xs<NT, SPT, CLAMP, DBG>(xc, tid, nz + t0, out + t0, live, acc1, acc2, totn1, totn2, lpeak, V, cur.ft0,
                                          cur.fj, cur.fj1, r1, r2, r3, nn.ft0, nn.fj, nn.fj1, r0n, r1n, dbg1, dbg2, t0, T);
    if (live) tpeak = fmaxf(tpeak, lpeak);
    __syncthreads();
    // ---- carry of tile t+1 ----
    acc1 = carry1 + ex1;
    acc2 = carry2 + ex2;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const double w1 = s_wsum[pb][0][w], w2 = s_wsum[pb][1][w];
      if (w < warp) {
        acc1 += w1;
        acc2 += w2;
      }
      carry1 += w1;
      carry2 += w2;
    }
    cur = nxt;
    nxt = nn;
  }
  return tpeak;
}

// Requires T % SPT == 0 and 16-byte aligned rows, as k_voice_audio_sp.
template <int NT, int SPT, int MINB, bool DBG>
__global__ void __launch_bounds__(NT, MINB) k_voice_audio_s2(AudioArgs A) {
  constexpr int TILE = NT * SPT;
  constexpr int NW = NT / 32;
  __shared__ float4 xs[2][2][SPT / 4][NT];  // phase increments of the two tiles in flight
  __shared__ double s_wsum[2][2][NW];       // [buffer][vco][warp]
  __shared__ float s_peak[NW];
  __shared__ int s_slot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = A.T;
  SpVoice V;
  V.scale = A.scale;
  V.sr = A.sr;
  V.rsr = A.rsr;

  for (;;) {
    if (tid == 0) s_slot = atomicAdd(A.counter, 1);
    __syncthreads();
    const int slot = s_slot;
    if (slot >= A.B) break;
    const int b = A.order[slot];
    const float* vc = A.vconst + (size_t)b * VC_COUNT;
    V.midi1 = vc[VC_MIDI1]; V.depth1 = vc[VC_DEPTH1]; V.phase1 = vc[VC_PHASE1];
    V.midi2 = vc[VC_MIDI2]; V.depth2 = vc[VC_DEPTH2]; V.phase2 = vc[VC_PHASE2];
    V.pk = vc[VC_PK]; V.shape = vc[VC_SHAPE];
    const bool noclamp = vc[VC_NOCLAMP] != 0.0f;
    float* out = A.audio + (size_t)b * T;
    const int ntiles = A.ntiles[b];
    float tpeak = 0.0f;
    if (ntiles > 0) {
      if (noclamp)
        tpeak = s2_render_voice<NT, SPT, false, DBG>(A, V, b, ntiles, xs, s_wsum);
      else
        tpeak = s2_render_voice<NT, SPT, true, DBG>(A, V, b, ntiles, xs, s_wsum);
    }
    // ---- silent tail ----
    if ((long long)ntiles * TILE < T) {
      float4* o4 = reinterpret_cast<float4*>(out);
      for (int i = ntiles * (TILE / 4) + tid; i < T / 4; i += NT) o4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // ---- per-voice peak, normalize_if_clipping (as in k_voice_audio) ----
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) tpeak = fmaxf(tpeak, __shfl_xor_sync(0xffffffffu, tpeak, d));
    if (lane == 0) s_peak[warp] = tpeak;
    __syncthreads();  // also orders this CTA's global stores before the re-read below
    float pkv = s_peak[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) pkv = fmaxf(pkv, s_peak[w]);
    if (tid == 0 && A.peak) A.peak[b] = pkv;
    if (A.normalize && pkv > 1.0f) {
      const float rp = vm::div(1.0f, pkv);
      const int live = min(T, ntiles * TILE);
      float4* o4 = reinterpret_cast<float4*>(out);
      const int n4 = live / 4;
      constexpr int U = 4;
      for (int base = 0; base < n4; base += U * NT) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = base + u * NT + tid;
          if (i < n4) v[u] = o4[i];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = base + u * NT + tid;
          if (i < n4) {
            v[u].x = div_const(v[u].x, pkv, rp); v[u].y = div_const(v[u].y, pkv, rp);
            v[u].z = div_const(v[u].z, pkv, rp); v[u].w = div_const(v[u].w, pkv, rp);
            o4[i] = v[u];
          }
        }
      }
    }
  }
}

