// Voice render for sm_100a: seeded parameters -> control-rate signals -> audio.
//
// Replaces torchsynth.synth.Voice as the reference calls it (vicreg_audio_params.py:86-94,114;
// audio_to_params.py:196-203,215,238-257).  Three kernels:
//   k_seed_params   one thread per sound: torch-CPU-compatible MT19937 (first 78 outputs of seed = sound id)
//   k_voice_control one CTA per voice: from_0to1, 6 ADSR, 2 LFO (fp64 phase scan), modulation matrix -> ctrl[B][5][C]
//   k_voice_audio   one CTA per voice, 8 samples per thread per tile: linear upsample of the 5 control signals,
//                   2 VCO pitch paths, fp64 block scan of the phase increments, oscillators, VCAs, noise, mix,
//                   running peak, optional normalize_if_clipping pass.
// HBM traffic per voice: 4T audio written (+4T noise read when the table is not L2 resident, +8T on the voices
// that clip when normalize != 0); ctrl/scratch are 11*C floats (L2 resident).
#include "ias_common.cuh"
#include "voice_math.cuh"

#include <algorithm>
#include <string>
#include <vector>

namespace ias {
namespace {

using namespace vm;

// ------------------------------------------------------------------------------------------------------------
// parameter names / orders (host)
// ------------------------------------------------------------------------------------------------------------
struct NameTable {
  std::vector<std::string> reg;      // "module/param", registration order
  std::vector<std::string> dotted;   // "module.torchparameters.param"
  int sorted_index[NROWS];           // registration row -> position in sorted(named_parameters())
  uint8_t reg_of_sorted[NROWS];      // inverse
};

const NameTable& names() {
  static const NameTable t = [] {
    NameTable n;
    const char* adsr[5] = {"attack", "decay", "sustain", "release", "alpha"};
    const char* lfo[8] = {"frequency", "mod_depth", "initial_phase", "sin", "tri", "saw", "rsaw", "sqr"};
    const char* vco[4] = {"tuning", "mod_depth", "initial_phase", "shape"};
    const char* mod_in[4] = {"adsr_1", "adsr_2", "lfo_1", "lfo_2"};
    const char* mod_out[5] = {"vco_1_pitch", "vco_1_amp", "vco_2_pitch", "vco_2_amp", "noise_amp"};
    auto push = [&](const std::string& m, const std::string& p) {
      n.reg.push_back(m + "/" + p);
      n.dotted.push_back(m + ".torchparameters." + p);
    };
    push("keyboard", "midi_f0");
    push("keyboard", "duration");
    for (const char* m : {"adsr_1", "adsr_2"}) for (auto p : adsr) push(m, p);
    for (const char* m : {"lfo_1", "lfo_2"}) for (auto p : lfo) push(m, p);
    for (const char* m : {"lfo_1_amp_adsr", "lfo_2_amp_adsr", "lfo_1_rate_adsr", "lfo_2_rate_adsr"})
      for (auto p : adsr) push(m, p);
    for (auto i : mod_in) for (auto o : mod_out) push("mod_matrix", std::string(i) + "->" + o);
    for (int i = 0; i < 3; ++i) push("vco_1", vco[i]);
    for (int i = 0; i < 4; ++i) push("vco_2", vco[i]);
    for (const char* p : {"vco_1", "vco_2", "noise"}) push("mixer", p);
    std::vector<int> order(NROWS);
    for (int i = 0; i < NROWS; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return n.dotted[a] < n.dotted[b]; });
    for (int s = 0; s < NROWS; ++s) {
      n.sorted_index[order[s]] = s;
      n.reg_of_sorted[s] = (uint8_t)order[s];
    }
    return n;
  }();
  return t;
}

// ------------------------------------------------------------------------------------------------------------
// k_seed_params
// ------------------------------------------------------------------------------------------------------------
struct SeedTable {
  uint8_t reg_of_sorted[NROWS];
  uint8_t frozen[NROWS];
};

constexpr int SEED_THREADS = 64;

__global__ void __launch_bounds__(SEED_THREADS) k_seed_params(long long first_id,
                                                              const long long* __restrict__ batch_idx_dev, int B,
                                                              SeedTable tab, float* __restrict__ params01,
                                                              uint8_t* __restrict__ is_train) {
  // MT19937 state words mt[0..78] of this CTA's sounds, [word][thread]: conflict-free, and no per-thread local array
  __shared__ uint32_t s_lo[NROWS + 1][SEED_THREADS];
  const int tid = threadIdx.x;
  const int i = blockIdx.x * SEED_THREADS + tid;
  if (i >= B) return;  // no barrier below: every thread only touches its own column
  if (batch_idx_dev) first_id = *batch_idx_dev * (long long)B;  // batch number read on the device (graph replay)
  const unsigned long long id = (unsigned long long)(first_id + i);
  // init_genrand(low 32 bits of the seed); only mt[0..78] and mt[397..474] feed outputs 0..77
  uint32_t s = (uint32_t)id;
  s_lo[0][tid] = s;
#pragma unroll 6
  for (int j = 1; j <= NROWS; ++j) {
    s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)j;
    s_lo[j][tid] = s;
  }
#pragma unroll 6
  for (int j = NROWS + 1; j < 397; ++j) s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)j;
  for (int k = 0; k < NROWS; ++k) {
    s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)(397 + k);  // mt[397 + k]
    const uint32_t y = (s_lo[k][tid] & 0x80000000u) | (s_lo[k + 1][tid] & 0x7fffffffu);
    uint32_t v = s ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    v ^= v >> 11;
    v ^= (v << 7) & 0x9d2c5680u;
    v ^= (v << 15) & 0xefc60000u;
    v ^= v >> 18;
    const int row = tab.reg_of_sorted[k];
    if (!tab.frozen[row]) params01[(size_t)row * B + i] = (float)(v & 0xffffffu) * (1.0f / 16777216.0f);
  }
  if (is_train) is_train[i] = ((id / 32ull) % 10ull) != 9ull;
}

// ------------------------------------------------------------------------------------------------------------
// block-wide exclusive scan of one double per thread (NT threads), total returned to every thread
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double shfl_up_d(double v, int d) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_up_sync(0xffffffffu, lo, d);
  hi = __shfl_up_sync(0xffffffffu, hi, d);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    double o = shfl_up_d(v, d);
    if (lane >= d) v += o;
  }
  return v;
}

[[maybe_unused]] __device__ __forceinline__ float warp_incl_scan(float v, int lane) {  // (ablation builds only)
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    float o = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += o;
  }
  return v;
}

// ------------------------------------------------------------------------------------------------------------
// k_voice_control
// ------------------------------------------------------------------------------------------------------------
constexpr int CTRL_THREADS = 256;
// CTAs per SM of the control kernel: 4 (64 registers, 40 B of spills) turn 1024 voices into 2 rounds of 592 instead
// of 3 rounds of 444.
#ifndef IAS_CTRL_MINB
#define IAS_CTRL_MINB 4
#endif
constexpr int REC_FLOATS = 16;  // floats per control-interval record (4 x 16-byte loads in the audio stage)

struct ControlShared {
  float P[NROWS];
  Lfo lfo[2];
  ModMatrix mm;
  double wsum[2][CTRL_THREADS / 32];
  int last_nz[CTRL_THREADS / 32];
  float range[CTRL_THREADS / 32][4];
};

// ------------------------------------------------------------------------------------------------------------
// k_voice_adsr: the six envelopes of every voice, one CTA per (envelope, voice): env[b][e][j], e in the order
// adsr_1, adsr_2, lfo_1_amp_adsr, lfo_2_amp_adsr, lfo_1_rate_adsr, lfo_2_rate_adsr.  This is where the SLEEF-exact
// pow lives (~300 dependent instructions, ~3.5 live per control point and voice); as its own kernel it needs < 64
// registers, so 2.5x more warps are resident to hide the dependency chains than inside the control kernel.
// ------------------------------------------------------------------------------------------------------------
constexpr int ADSR_THREADS = 128;

__global__ void __launch_bounds__(ADSR_THREADS, 8)
k_voice_adsr(const float* __restrict__ params01, int B, int C, float cr, float eps, RangeTable ranges,
             float* __restrict__ env) {
  __shared__ float s_v[6];  // attack, decay, sustain, release, alpha, keyboard duration (through from_0to1)
  __shared__ Adsr s_adsr;
  const int b = blockIdx.x, e = blockIdx.y;  // voices on x: no 65535 limit on the batch
  const int tid = threadIdx.x;
  const int base[6] = {ADSR1, ADSR2, LFO1_AMP, LFO2_AMP, LFO1_RATE, LFO2_RATE};
  if (tid < 6) {
    const int row = tid < 5 ? base[e] + tid : (int)KEY_DURATION;
    s_v[tid] = from_0to1(params01[(size_t)row * B + b], ranges.r[row]);
  }
  __syncthreads();
  if (tid < 3) {
    // the three ramps of the envelope on three lanes of ONE warp running the SAME code (duration / start / direction
    // chosen with selects): ramp_setup -- two SLEEF pows and the saturation search, ~800 instructions -- executes
    // once for the three lanes instead of three times in a row, as it did behind adsr_setup_part's if / else chain
    const float note_on = s_v[5], alpha = s_v[4];
    const float new_attack = fminf(s_v[0], note_on);
    const float new_decay = fminf(fmaxf(sub(note_on, s_v[0]), 0.0f), s_v[1]);
    const float dur = mul(tid == 0 ? new_attack : (tid == 1 ? new_decay : s_v[3]), cr);
    const float start = tid == 0 ? 0.0f : mul(tid == 1 ? new_attack : note_on, cr);
    const Ramp r = ramp_setup(dur, start, tid != 0, alpha, eps, C);
    Ramp* dst = tid == 0 ? &s_adsr.a : (tid == 1 ? &s_adsr.d : &s_adsr.r);
    *dst = r;
    if (tid == 0) {
      s_adsr.sustain = s_v[2];
      s_adsr.one_minus_sustain = sub(1.0f, s_v[2]);
      s_adsr.alpha = alpha;
    }
  }
  __syncthreads();
  float* out = env + ((size_t)b * 6 + e) * C;
  for (int j = tid; j < C; j += ADSR_THREADS) out[j] = adsr_eval(s_adsr, (float)j, eps);
}

// Per-interval record j of one voice, read by the audio-stage threads whose samples start in control interval j:
//   [0..2]  vco_1_pitch at points j, j+1, j+2 (clamped to C-1)      exact values: they feed the phase
//   [3..5]  vco_2_pitch at points j, j+1, j+2
//   [6..8]  AmpLine of vco_1_amp  * mixer level 1
//   [9..11] AmpLine of vco_2_amp  * mixer level 2 * (1 - shape/2)
//   [12..14] AmpLine of noise_amp * mixer level 3
// `env` holds the six envelopes (k_voice_adsr).  SMEM: the two LFO phase rows and the five output rows of the voice
// live in shared memory (7*C floats: 4 s clips), so the dependent phases never wait on global memory; otherwise
// (30 s clips: 370 KB) the phase rows are built in place in `env` and the output rows in `ctrl_ws`.
// ctrl_out (may be null) receives the five signals [B][5][C] for inspection.
template <bool SMEM>
__global__ void __launch_bounds__(CTRL_THREADS, IAS_CTRL_MINB)
k_voice_control(const float* __restrict__ params01, int B, int C, float cr, float eps, RangeTable ranges,
                const float* __restrict__ ctrl_in, float* __restrict__ ctrl_out, float* __restrict__ ctrl_ws,
                float* __restrict__ env_all, float* __restrict__ vconst, float4* __restrict__ rec) {
  extern __shared__ __align__(16) float s_rows[];
  __shared__ ControlShared sh;
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;

  if (tid < NROWS) sh.P[tid] = from_0to1(params01[(size_t)tid * B + b], ranges.r[tid]);
  __syncthreads();
  // per-voice setup, one item per thread of different warps so the serial pow / division chains overlap
  if (tid < 10) {  // the five mode weights of each LFO: w ** e
    sh.lfo[tid / 5].w[tid % 5] = pow_sleef(sh.P[(tid < 5 ? LFO1 : LFO2) + 3 + tid % 5], 2.718281828f);
  } else if (tid == 32) {
    sh.mm = modmatrix_setup(&sh.P[MODM]);
  } else if (tid == 64) {
    voice_constants(sh.P, vconst + (size_t)b * VC_COUNT);
  }
  __syncthreads();
  if (tid < 2) lfo_finish(sh.lfo[tid], &sh.P[tid == 0 ? LFO1 : LFO2]);
  __syncthreads();

  float* env = env_all + (size_t)b * 6 * C;  // [6][C]: adsr_1, adsr_2, lfo_1 amp, lfo_2 amp, lfo_1 rate, lfo_2 rate
  float* ph0 = SMEM ? s_rows : env + 4 * C;
  float* ph1 = SMEM ? s_rows + C : env + 5 * C;
  float* out = SMEM ? s_rows + 2 * C : ctrl_ws + (size_t)b * IAS_VOICE_NCONTROL * C;
  // Phase A: LFO phase increments from the rate envelopes
  const float rcr = rcp(cr);
  constexpr int UJ = 4;  // control points per thread in flight: the envelope loads of all of them are issued first
  for (int jb = tid; jb < C; jb += UJ * CTRL_THREADS) {
    float e4[UJ], e5[UJ];
#pragma unroll
    for (int u = 0; u < UJ; ++u) {
      const int j = jb + u * CTRL_THREADS;
      if (j < C) {
        // (the read-only path only when the rows are not rewritten in place by this kernel)
        e4[u] = SMEM ? __ldg(env + 4 * C + j) : env[4 * C + j];
        e5[u] = SMEM ? __ldg(env + 5 * C + j) : env[5 * C + j];
      }
    }
#pragma unroll
    for (int u = 0; u < UJ; ++u) {
      const int j = jb + u * CTRL_THREADS;
      if (j < C) {
        const float i0 = lfo_increment(sh.lfo[0], e4[u], cr, rcr), i1 = lfo_increment(sh.lfo[1], e5[u], cr, rcr);
        ph0[j] = i0;
        ph1[j] = i1;
      }
    }
  }
  __syncthreads();

  // Phase B: inclusive scan of the two increment rows, fp64 accumulate, fp32 round per element (torch CPU cumsum)
  const int chunk = (C + CTRL_THREADS - 1) / CTRL_THREADS;
  const int j0 = min(tid * chunk, C), j1 = min(j0 + chunk, C);
  double tot[2] = {0.0, 0.0};
  for (int j = j0; j < j1; ++j) {
    tot[0] += (double)ph0[j];
    tot[1] += (double)ph1[j];
  }
  double pre[2];
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    double inc = warp_incl_scan(tot[l], lane);
    if (lane == 31) sh.wsum[l][warp] = inc;
    pre[l] = inc - tot[l];
  }
  __syncthreads();
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    double acc = pre[l];
    for (int w = 0; w < warp; ++w) acc += sh.wsum[l][w];
    const float phase0 = sh.lfo[l].initial_phase;
    float* ph = l ? ph1 : ph0;
    for (int j = j0; j < j1; ++j) {
      acc += (double)ph[j];
      ph[j] = add((float)acc, phase0);
    }
  }
  __syncthreads();

  // Phase C: LFO shapes, VCAs, modulation matrix
  float* user = ctrl_out ? ctrl_out + (size_t)b * IAS_VOICE_NCONTROL * C : nullptr;
  for (int jb = tid; jb < C; jb += UJ * CTRL_THREADS) {
    float e0[UJ], e1[UJ], e2[UJ], e3[UJ];
#pragma unroll
    for (int u = 0; u < UJ; ++u) {
      const int j = jb + u * CTRL_THREADS;
      if (j < C) {
        e0[u] = SMEM ? __ldg(env + 0 * C + j) : env[0 * C + j];
        e1[u] = SMEM ? __ldg(env + 1 * C + j) : env[1 * C + j];
        e2[u] = SMEM ? __ldg(env + 2 * C + j) : env[2 * C + j];
        e3[u] = SMEM ? __ldg(env + 3 * C + j) : env[3 * C + j];
      }
    }
#pragma unroll
    for (int u = 0; u < UJ; ++u) {
      const int j = jb + u * CTRL_THREADS;
      if (j < C) {
        const float l1 = mul(lfo_shapes_mix(sh.lfo[0], ph0[j]), e2[u]);
        const float l2 = mul(lfo_shapes_mix(sh.lfo[1], ph1[j]), e3[u]);
#pragma unroll
        for (int o = 0; o < 5; ++o) {
          const float v = modmatrix_out(sh.mm, o, e0[u], e1[u], l1, l2);
          out[o * C + j] = v;
          if (user) user[o * C + j] = v;
        }
      }
    }
  }
  __syncthreads();  // this CTA's output rows are re-read below

  // Phase D: per-interval records for the audio stage (from the caller's signals when the parity hook supplies them)
  const float* src = ctrl_in ? ctrl_in + (size_t)b * IAS_VOICE_NCONTROL * C : out;
  const float shape = sh.P[VCO2 + 3];
  const float lev1 = sh.P[MIX + 0];
  const float lev2 = mul(sh.P[MIX + 1], sub(1.0f, mul(shape, 0.5f)));
  const float lev3 = sh.P[MIX + 2];
  float4* rv = rec + (size_t)b * C * (REC_FLOATS / 4);
  int last_nz = -1;  // last control point where any amplitude signal (vco_1_amp, vco_2_amp, noise_amp) is non-zero
  float lo1 = 1e30f, hi1 = -1e30f, lo2 = 1e30f, hi2 = -1e30f;  // range of the two pitch signals
  for (int j = tid; j < C; j += CTRL_THREADS) {
    const int ja = min(j + 1, C - 1), jb = min(j + 2, C - 1);
    const float* s0 = src + 0 * C;
    const float* s1 = src + 1 * C;
    const float* s2 = src + 2 * C;
    const float* s3 = src + 3 * C;
    const float* s4 = src + 4 * C;
    const AmpLine g1 = amp_line(s1[j], s1[ja], s1[jb], lev1);
    const AmpLine g2 = amp_line(s3[j], s3[ja], s3[jb], lev2);
    const AmpLine g3 = amp_line(s4[j], s4[ja], s4[jb], lev3);
    rv[j * 4 + 0] = make_float4(s0[j], s0[ja], s0[jb], s2[j]);
    rv[j * 4 + 1] = make_float4(s2[ja], s2[jb], g1.c0, g1.d0);
    rv[j * 4 + 2] = make_float4(g1.dd, g2.c0, g2.d0, g2.dd);
    rv[j * 4 + 3] = make_float4(g3.c0, g3.d0, g3.dd, 0.0f);
    if (s1[j] != 0.0f || s3[j] != 0.0f || s4[j] != 0.0f) last_nz = j;
    lo1 = fminf(lo1, s0[j]);
    hi1 = fmaxf(hi1, s0[j]);
    lo2 = fminf(lo2, s2[j]);
    hi2 = fmaxf(hi2, s2[j]);
  }
  // Envelopes end in exact zeros (pow(0, alpha) == 0 after the release): tell the audio stage where the silent tail
  // starts so it can write zeros instead of rendering oscillators that are multiplied by 0.
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    last_nz = max(last_nz, __shfl_xor_sync(0xffffffffu, last_nz, d));
    lo1 = fminf(lo1, __shfl_xor_sync(0xffffffffu, lo1, d));
    hi1 = fmaxf(hi1, __shfl_xor_sync(0xffffffffu, hi1, d));
    lo2 = fminf(lo2, __shfl_xor_sync(0xffffffffu, lo2, d));
    hi2 = fmaxf(hi2, __shfl_xor_sync(0xffffffffu, hi2, d));
  }
  if (lane == 0) {
    sh.last_nz[warp] = last_nz;
    sh.range[warp][0] = lo1;
    sh.range[warp][1] = hi1;
    sh.range[warp][2] = lo2;
    sh.range[warp][3] = hi2;
  }
  __syncthreads();
  if (tid == 0) {
    int m = -1;
    for (int w = 0; w < CTRL_THREADS / 32; ++w) {
      m = max(m, sh.last_nz[w]);
      lo1 = fminf(lo1, sh.range[w][0]);
      hi1 = fmaxf(hi1, sh.range[w][1]);
      lo2 = fminf(lo2, sh.range[w][2]);
      hi2 = fmaxf(hi2, sh.range[w][3]);
    }
    vconst[(size_t)b * VC_COUNT + VC_SILENT_FROM] = (float)(m + 1);
    // Interpolated values stay within the range of the control points (up to an ulp), so with a 0.5-semitone margin
    // clamp(midi + depth*mod, 0, 127) provably never engages.  NaNs fail the comparisons and keep the clamp.
    const float midi1 = add(sh.P[KEY_MIDI_F0], sh.P[VCO1 + 0]), d1 = sh.P[VCO1 + 1];
    const float midi2 = add(sh.P[KEY_MIDI_F0], sh.P[VCO2 + 0]), d2 = sh.P[VCO2 + 1];
    const float a1 = midi1 + fminf(d1 * lo1, d1 * hi1), b1 = midi1 + fmaxf(d1 * lo1, d1 * hi1);
    const float a2 = midi2 + fminf(d2 * lo2, d2 * hi2), b2 = midi2 + fmaxf(d2 * lo2, d2 * hi2);
    const bool inside = a1 > 0.5f && b1 < 126.5f && a2 > 0.5f && b2 < 126.5f;
    vconst[(size_t)b * VC_COUNT + VC_NOCLAMP] = inside ? 1.0f : 0.0f;
  }
}

// ------------------------------------------------------------------------------------------------------------
// k_voice_schedule: live tile count per voice and the order the audio CTAs pull voices in (longest first), so that
// the SMs finish together although voices differ in length (note duration + release) by two orders of magnitude.
// One CTA; counting sort on the live tile count.
// ------------------------------------------------------------------------------------------------------------
constexpr int SCHED_THREADS = 1024;
constexpr int SCHED_BINS = 2048;

__global__ void __launch_bounds__(SCHED_THREADS)
k_voice_schedule(const float* __restrict__ vconst, int B, int T, float scale, int tile, int render_all,
                 int* __restrict__ ntiles, int* __restrict__ order, int* __restrict__ counter) {
  __shared__ int s_bin[SCHED_BINS];
  __shared__ int s_wtot[SCHED_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int all = (T + tile - 1) / tile;
  for (int i = tid; i < SCHED_BINS; i += SCHED_THREADS) s_bin[i] = 0;
  if (tid == 0) *counter = 0;
  __syncthreads();
  // Samples whose control interval [i0, i0+1] lies wholly in the silent tail are exactly 0 (every VCA gain is 0):
  // only the tiles before the first tile that starts inside the tail are rendered.
  auto starts_silent = [&](int t, int silent_from) { return (int)mul(scale, (float)(t * tile)) >= silent_from; };
  auto key_of = [&](int nt) { return (SCHED_BINS - 1) - (int)(((long long)nt * (SCHED_BINS - 1)) / all); };  // long first
  for (int v = tid; v < B; v += SCHED_THREADS) {
    int nt = all;
    if (!render_all) {
      const int silent_from = (int)fminf(vconst[(size_t)v * VC_COUNT + VC_SILENT_FROM], 1e9f);
      int t = min(all, (int)((float)silent_from / (scale * (float)tile)));
      while (t > 0 && starts_silent(t - 1, silent_from)) --t;
      while (t < all && !starts_silent(t, silent_from)) ++t;
      nt = t;
    }
    ntiles[v] = nt;
    atomicAdd(&s_bin[key_of(nt)], 1);
  }
  __syncthreads();
  // exclusive scan of the bins (2 per thread)
  const int c0 = s_bin[2 * tid], c1 = s_bin[2 * tid + 1];
  int inc = c0 + c1;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_wtot[warp] = inc;
  __syncthreads();
  int base = inc - (c0 + c1);
  for (int w = 0; w < warp; ++w) base += s_wtot[w];
  s_bin[2 * tid] = base;
  s_bin[2 * tid + 1] = base + c0;
  __syncthreads();
  for (int v = tid; v < B; v += SCHED_THREADS) order[atomicAdd(&s_bin[key_of(ntiles[v])], 1)] = v;
}

// ------------------------------------------------------------------------------------------------------------
// k_voice_audio
// ------------------------------------------------------------------------------------------------------------
#ifndef IAS_AUDIO_PREFETCH
#define IAS_AUDIO_PREFETCH 1
#endif

struct AudioArgs {
  const float4* rec;    // [B][C][4] per-interval records (k_voice_control)
  const float* vconst;  // [B][16]
  const float* noise;   // [R][T]
  const int* ntiles;    // [B] live tiles of each voice (k_voice_schedule)
  const int* order;     // [B] voice ids, longest first
  int* counter;         // work queue head
  float* audio;         // [B][T]
  float* peak;          // [B] or null
  float* phase_dbg;     // [B][2][T] or null
  int B, T, C, noise_rows;
  float scale;          // float(C-1)/float(T-1)
  float sr, rsr;
  int normalize;  // 0: none; 1: normalize_if_clipping in the kernel; 2: deferred -- audio stays un-normalised and
                  // peak[b] receives the factor the consumer applies (1/peak of a clipping row, else 1)
};

// ---- packed fp32 (sm_100: FFMA2 / FADD2 / FMUL2 process two fp32 lanes per instruction) --------------------------
// The audio kernel is bound by instruction issue, and 2/3 of its instructions are fp32 arithmetic: evaluating two
// adjacent samples per packed instruction halves those issue slots.  Each lane is an IEEE round-to-nearest fp32
// operation, so results are bit-identical to the scalar helpers in voice_math.cuh.  One caveat, measured with
// cuobjdump: ptxas contracts mul.rn.f32x2 feeding add/sub.rn.f32x2 into FFMA2 even with --fmad=false, so wherever the
// reference rounds a product before adding to it, one of the two operations is issued in scalar form (which ptxas
// never contracts).  tests/test_gpu_voice.py compares every phase argument bit for bit against the torch CPU path.
struct P2 {
  unsigned long long v;
};
__device__ __forceinline__ P2 p2(float lo, float hi) {
  P2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ P2 p2b(float a) { return p2(a, a); }  // broadcast (a scalar / immediate operand in SASS)
#pragma nv_diag_suppress 550  // the unused half of an unpacked pair
__device__ __forceinline__ float p2lo(P2 a) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
  (void)y;
  return x;
}
__device__ __forceinline__ float p2hi(P2 a) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
  (void)x;
  return y;
}
#pragma nv_diag_default 550
__device__ __forceinline__ P2 p2_fma(P2 a, P2 b, P2 c) {
  P2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
__device__ __forceinline__ P2 p2_add(P2 a, P2 b) {
  P2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ P2 p2_sub(P2 a, P2 b) {
  P2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ P2 p2_mul(P2 a, P2 b) {
  P2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}

// vm::vco_increment for two samples: 2*pi*hz(clamp(midi + depth*mod, 0, 127)) / sample_rate, same op order.
template <bool CLAMP>
__device__ __forceinline__ P2 vco_increment_p2(float midi, float depth, P2 mod, float sr, float rsr) {
#if IAS_ABL_NOPASS1
  return p2_fma(p2b(depth * rsr), mod, p2b(midi * rsr));
#endif
  // depth * mod must keep its own rounding (see the note on contraction above): issued as fma(depth, mod, +0.0),
  // which ptxas can neither fold into a multiply (a -0 product would change sign) nor contract with the add
  P2 m = p2_add(p2b(midi), p2_fma(p2b(depth), mod, p2b(0.0f)));
  if (CLAMP)  // voices whose range is provably inside (VC_NOCLAMP) skip the four min/max instructions
    m = p2(fminf(fmaxf(p2lo(m), 0.0f), 127.0f), fminf(fmaxf(p2hi(m), 0.0f), 127.0f));
  const P2 a = p2_add(m, p2b(-69.0f));
  // a / 12: Markstein step with RN(1/12)  (div_const)
  const P2 q = p2_mul(a, p2b(0.0833333358168601989746f));
  const P2 d = p2_fma(p2_fma(q, p2b(-12.0f), a), p2b(0.0833333358168601989746f), q);
  // exp2_fast
  const float magic = 12582912.0f;
  const P2 t = p2_add(d, p2b(magic));
  const P2 sfrac = p2_sub(d, p2_add(t, p2b(-magic)));
  P2 u = p2b(+0.1535920892e-3f);
  u = p2_fma(u, sfrac, p2b(+0.1339262701e-2f));
  u = p2_fma(u, sfrac, p2b(+0.9618384764e-2f));
  u = p2_fma(u, sfrac, p2b(+0.5550347269e-1f));
  u = p2_fma(u, sfrac, p2b(+0.2402264476e+0f));
  u = p2_fma(u, sfrac, p2b(+0.6931471825e+0f));
  u = p2_fma(u, sfrac, p2b(1.0f));
  // exponent add: (bits(t) - 0x4B400000) << 23 == bits(t) << 23 (the magic's low 9 bits are zero)
  const float e0 = i2f(f2i(p2lo(u)) + (f2i(p2lo(t)) << 23));
  const float e1 = i2f(f2i(p2hi(u)) + (f2i(p2hi(t)) << 23));
  const P2 w = p2_mul(p2b(IAS_TWO_PI_F), p2_mul(p2b(440.0f), p2(e0, e1)));
  // w / sr: Markstein step with RN(1/sr)
  const P2 qq = p2_mul(w, p2b(rsr));
  return p2_fma(p2_fma(qq, p2b(-sr), w), p2b(rsr), qq);
}

// vm::reduce_half_turns for two arguments: a = (n + f) * pi, f in [-0.5, 0.5]; t carries n in its low mantissa bits.
[[maybe_unused]] __device__ __forceinline__ void reduce_half_turns_p2(P2 a, P2& f, P2& t) {
  const float C1 = 0.318309873342514038f;
  const float C2 = (float)(0.31830988618379067154 - (double)0.318309873342514038f);
  const float magic = 12582912.0f;
  t = p2_fma(a, p2b(C1), p2b(magic));
  const P2 nneg = p2_sub(p2b(magic), t);  // -rint(a / pi)
  f = p2_fma(a, p2b(C2), p2_fma(a, p2b(C1), nneg));
}
// ---- ablation switches: DIAGNOSTIC variant builds only (tools/gpu_ablate.sh); they change the results on purpose to
// show what each part of the tile loop costs.  Never defined in the shipped library.
#ifndef IAS_ABL_NOXU     // SFU cosines / ex2 / rcp of the oscillators replaced by one FMA each
#define IAS_ABL_NOXU 0
#endif
#ifndef IAS_ABL_NOPASS1  // the bit-exact pitch -> increment chain replaced by one FMA
#define IAS_ABL_NOPASS1 0
#endif
#ifndef IAS_ABL_NOF64    // phase accumulated in fp32: no fp32<->fp64 conversions, no DADD
#define IAS_ABL_NOF64 0
#endif
#ifndef IAS_ABL_NOBAR    // no CTA barrier / cross-warp carry (each warp scans alone)
#define IAS_ABL_NOBAR 0
#endif
#if IAS_ABL_NOXU
#define IAS_COS(x) fmaf((x), (x), -1.0f)
#else
#define IAS_COS(x) __cosf(x)
#endif

// Oscillator arguments are reduced in RADIANS with a two-constant Cody-Waite step: n = rint(a / P) from one FMA against
// the 1.5 * 2^23 magic constant, then y = fma(-n, P_hi, a) (exact product, ONE rounding relative to the small result,
// so accuracy near the zeros of sin is kept) and y = fma(-n, P_lo, y) (P_lo = P - P_hi, |n P_lo| <= 0.07 for 30 s
// clips, its own error <= 5e-9).  The SFU cosine then takes y as it is -- reducing to turns costs one more multiply by
// 2 pi that cos.approx undoes again (FMUL.RZ by 1/(2 pi) + MUFU.COS).  Amplitude path: absolute error ~5e-7.
#ifndef IAS_AUDIO_RADIANS
#define IAS_AUDIO_RADIANS 1
#endif
// cos(a) for the sine VCO: P = 2 pi, y in [-pi, pi].
__device__ __forceinline__ P2 cos_arg_p2(P2 a) {
  const float magic = 12582912.0f;
#if IAS_AUDIO_RADIANS
  const float C1 = 0.159154936671257019f;                        // (float)(1/(2 pi))
  const float P_HI = 6.2831854820251465f, P_LO = -1.7484555314695172e-07f;
  const P2 t = p2_fma(a, p2b(C1), p2b(magic));
  const P2 nneg = p2_sub(p2b(magic), t);                          // -rint(a / (2 pi)), exact
  const P2 y = p2_fma(nneg, p2b(P_LO), p2_fma(nneg, p2b(P_HI), a));
  return p2(IAS_COS(p2lo(y)), IAS_COS(p2hi(y)));
#else
  const float C1 = 0.159154936671257019f;                                              // (float)(1/(2 pi))
  const float C2 = (float)(0.15915494309189533577 - (double)0.159154936671257019f);    // 1/(2 pi) - C1
  const P2 t = p2_fma(a, p2b(C1), p2b(magic));
  const P2 nneg = p2_sub(p2b(magic), t);
  const P2 fr = p2_mul(p2_fma(a, p2b(C2), p2_fma(a, p2b(C1), nneg)), p2b(IAS_TWO_PI_F));
  return p2(__cosf(p2lo(fr)), __cosf(p2hi(fr)));
#endif
}
// vm::squaresaw_core for two arguments: tanh(pk * sin(a)) * (1 + shape * cos(a)).  With a = n pi + y and
// sigma = (-1)^n: sin(a) = sigma * sin(y), cos(a) = sigma * cos(y), and tanh is odd, so the product equals
// tanh(pk * sin(y)) * (sigma + shape * cos(y)) -- the parity enters once, as the float +-1 (one shift-add).
__device__ __forceinline__ float parity_sign(float t) { return i2f((f2i(t) << 31) + 0x3f800000); }
__device__ __forceinline__ P2 squaresaw_core_p2(P2 a, float pk, float shape) {
#if IAS_AUDIO_RADIANS
  // P = pi, y in [-pi/2, pi/2] (may overshoot by an ulp of a/pi); sin(y) = y * Q(y^2), the degree-9 polynomial of
  // vm::sinpi_poly rescaled to radians (relative error <= 2.5e-7 on |y| <= 1.77: the zero crossings SquareSawVCO
  // amplifies through tanh stay accurate)
  const float magic = 12582912.0f;
  const float C1 = 0.318309873342514038f;                        // (float)(1/pi)
  const float P_HI = 3.1415927410125732f, P_LO = -8.7422776573475858e-08f;
  const P2 t = p2_fma(a, p2b(C1), p2b(magic));
  const P2 nneg = p2_sub(p2b(magic), t);
  const P2 y = p2_fma(nneg, p2b(P_LO), p2_fma(nneg, p2b(P_HI), a));
  const P2 u = p2_mul(y, y);
  // (folding pk into per-voice coefficients saves this multiply but costs five registers: measured slower, 0.725 vs
  // 0.719 ms, more spills)
  P2 p = p2b(2.5610734156e-06f);
  p = p2_fma(p, u, p2b(-1.9786701887e-04f));
  p = p2_fma(p, u, p2b(8.3326986060e-03f));
  p = p2_fma(p, u, p2b(-1.6666640341e-01f));
  p = p2_fma(p, u, p2b(9.9999994040e-01f));
  const P2 sc = p2_mul(p2b(pk), p2_mul(p, y));  // pk * sin(y)
  const P2 fr = y;
#else
  P2 f, t;
  reduce_half_turns_p2(a, f, t);
  const P2 u = p2_mul(f, f);
  P2 p = p2b(0.07634329050779343f);
  p = p2_fma(p, u, p2b(-0.59761643409729f));
  p = p2_fma(p, u, p2b(2.5499696731567383f));
  p = p2_fma(p, u, p2b(-5.1677045822143555f));
  p = p2_fma(p, u, p2b(3.141592502593994f));
  const P2 sc = p2_mul(p2b(pk), p2_mul(p, f));  // pk * sin(pi f)
  const P2 fr = p2_mul(f, p2b(IAS_PI_F));
#endif
  float e0, e1, r0, r1;
#if IAS_ABL_NOXU
  e0 = fmaf(p2lo(sc), 0.5f, 1.0f); e1 = fmaf(p2hi(sc), 0.5f, 1.0f);
#else
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(p2lo(sc)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(p2hi(sc)));
#endif
  const P2 ep = p2_add(p2(e0, e1), p2b(1.0f));
#if IAS_ABL_NOXU
  r0 = fmaf(p2lo(ep), -0.25f, 1.0f); r1 = fmaf(p2hi(ep), -0.25f, 1.0f);
#else
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(p2lo(ep)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(p2hi(ep)));
#endif
  const P2 th = p2_fma(p2b(-2.0f), p2(r0, r1), p2b(1.0f));
  const P2 c = p2(IAS_COS(p2lo(fr)), IAS_COS(p2hi(fr)));
  return p2_mul(th, p2_fma(p2b(shape), c, p2(parity_sign(p2lo(t)), parity_sign(p2hi(t)))));
}

// Pass 1 of a tile for one thread: the phase increments of both VCOs for its SPT samples (two samples per packed
// instruction), bit for bit the reference's fp32 op sequence.  srcs[] keeps the fp32 source coordinates for pass 2.
template <int SPT, bool VEC, bool CLAMP>
__device__ __forceinline__ void pitch_pass(float (&x1)[SPT], float (&x2)[SPT], float (&srcs)[SPT], const float4 r0,
                                           const float4 r1, float ft0, float scale, float fj, float fj1, float midi1,
                                           float depth1, float midi2, float depth2, float sr, float rsr, int t0, int T) {
#pragma unroll
  for (int k = 0; k < SPT; k += 2) {
    const P2 fi = p2_add(p2b(ft0), p2((float)k, (float)(k + 1)));
    // the reference's fp32 source coordinate, rounded before the subtraction: fma(scale, i, +0.0), as above
    const P2 sp = p2_fma(p2b(scale), fi, p2b(0.0f));
    const float s0 = p2lo(sp), s1 = p2hi(sp);
    srcs[k] = s0;
    srcs[k + 1] = s1;
    const bool d0 = s0 >= fj1, d1 = s1 >= fj1;
    const P2 l1 = p2_sub(p2(s0, s1), p2(d0 ? fj1 : fj, d1 ? fj1 : fj));  // in [0,1): j = floor(src) of sample 0
    const P2 l0 = p2_sub(p2b(1.0f), l1);
    // upsample_mix: fma(l0, x[i0], l1 * x[i1])
    const P2 m1 = p2_fma(l0, p2(d0 ? r0.y : r0.x, d1 ? r0.y : r0.x), p2_mul(l1, p2(d0 ? r0.z : r0.y, d1 ? r0.z : r0.y)));
    const P2 m2 = p2_fma(l0, p2(d0 ? r1.x : r0.w, d1 ? r1.x : r0.w), p2_mul(l1, p2(d0 ? r1.y : r1.x, d1 ? r1.y : r1.x)));
    const P2 i1 = vco_increment_p2<CLAMP>(midi1, depth1, m1, sr, rsr);
    const P2 i2 = vco_increment_p2<CLAMP>(midi2, depth2, m2, sr, rsr);
    // VEC: T % SPT == 0, so a thread is wholly live or wholly past the end; dead threads sit after every live
    // one in the last tile, and an inclusive scan never feeds later totals into earlier lanes -> no masking.
    const bool live0 = VEC || (t0 + k) < T, live1 = VEC || (t0 + k + 1) < T;
    x1[k] = live0 ? p2lo(i1) : 0.0f;
    x1[k + 1] = live1 ? p2hi(i1) : 0.0f;
    x2[k] = live0 ? p2lo(i2) : 0.0f;
    x2[k + 1] = live1 ? p2hi(i2) : 0.0f;
  }
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Persistent CTAs pull voices from the queue.  Per voice: tiles of NT*SPT samples, SPT consecutive samples per thread.
//   pass 1  pitch path of both VCOs, bit for bit the reference's fp32 op sequence -> phase increments x1, x2
//   scan    fp64 block scan of the increments (exact for 4 s clips), carry across tiles
//   pass 2  oscillators, VCA gains (AmpLine), noise, mix, running peak
template <int NT, int SPT, int MINB, bool VEC, bool DBG>
__global__ void __launch_bounds__(NT, MINB) k_voice_audio(AudioArgs A) {
  constexpr bool PREFETCH = IAS_AUDIO_PREFETCH != 0;
  constexpr int TILE = NT * SPT;
  constexpr int NW = NT / 32;
#if IAS_ABL_NOF64
  using acc_t = float;
#else
  using acc_t = double;
#endif
  __shared__ acc_t s_wsum[2][2][NW];  // [buffer][vco][warp]
  __shared__ float s_peak[NW];
  __shared__ int s_slot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = A.T, C = A.C;
  const float scale = A.scale;

  for (;;) {
    if (tid == 0) s_slot = atomicAdd(A.counter, 1);
    __syncthreads();
    const int slot = s_slot;
    if (slot >= A.B) break;
    const int b = A.order[slot];
    const float* vc = A.vconst + (size_t)b * VC_COUNT;
    const float midi1 = vc[VC_MIDI1], depth1 = vc[VC_DEPTH1], phase1 = vc[VC_PHASE1];
    const float midi2 = vc[VC_MIDI2], depth2 = vc[VC_DEPTH2], phase2 = vc[VC_PHASE2];
    const float pk = vc[VC_PK], shape = vc[VC_SHAPE];
    const bool noclamp = vc[VC_NOCLAMP] != 0.0f;
    const float4* rec = A.rec + (size_t)b * C * (REC_FLOATS / 4);
    const float* nz = A.noise + (size_t)(b % A.noise_rows) * T;
    float* out = A.audio + (size_t)b * T;
    const int ntiles = A.ntiles[b];

    acc_t carry1 = 0, carry2 = 0;
    float tpeak = 0.0f;
    float ft0 = (float)(tid * SPT);  // float(index of the thread's first sample); exact, advanced by TILE per tile
    for (int tile = 0; tile < ntiles; ++tile, ft0 += (float)TILE) {
      const int t0 = tile * TILE + tid * SPT;
      const int buf = tile & 1;
      // control interval of the thread's first sample (its SPT samples touch intervals j and j+1 only)
      const int j = min((int)mul(scale, fminf(ft0, (float)(T - 1))), C - 1);
      const float fj = (float)j;
      const float fj1 = add(fj, 1.0f);
      const float4* rj = rec + (size_t)j * 4;
      if (PREFETCH && tile + 1 < ntiles) {
        // the next tile's record (64 B) and noise (4*SPT B) of this thread: bring them into L1 now, so the loads at
        // the top of the next iteration do not expose an L2 / HBM round trip (they were ~10 % of the stall samples)
        const int jn = min((int)mul(scale, fminf(add(ft0, (float)TILE), (float)(T - 1))), C - 1);
        prefetch_l1(rec + (size_t)jn * 4);
        if (VEC && t0 + TILE < T) prefetch_l1(nz + t0 + TILE);
      }
      // ---- pass 1: phase increments of both VCOs ---------------------------------------------------------
      float x1[SPT], x2[SPT], srcs[SPT];
      const float4 r0 = __ldg(rj + 0);
      const float4 r1 = __ldg(rj + 1);
#ifdef IAS_AUDIO_COUNT_NOCLAMP_ONLY  // tools/issue_model.py: object for instruction counting only (never shipped)
      if (true)
#else
      if (noclamp)
#endif
        pitch_pass<SPT, VEC, false>(x1, x2, srcs, r0, r1, ft0, scale, fj, fj1, midi1, depth1, midi2, depth2, A.sr, A.rsr,
                                    t0, T);
      else
        pitch_pass<SPT, VEC, true>(x1, x2, srcs, r0, r1, ft0, scale, fj, fj1, midi1, depth1, midi2, depth2, A.sr, A.rsr,
                                   t0, T);
      // ---- block scan -------------------------------------------------------------------------------------
      acc_t tot1 = 0, tot2 = 0;
#pragma unroll
      for (int k = 0; k < SPT; ++k) {
        tot1 += (acc_t)x1[k];
        tot2 += (acc_t)x2[k];
      }
      const acc_t inc1 = warp_incl_scan(tot1, lane);
      const acc_t inc2 = warp_incl_scan(tot2, lane);
      if (lane == 31) {
        s_wsum[buf][0][warp] = inc1;
        s_wsum[buf][1][warp] = inc2;
      }
      // loads of pass 2 issued before the barrier so their latency overlaps it
      const float4 r2 = __ldg(rj + 2);
      const float4 r3 = __ldg(rj + 3);
      float nzv[SPT];
      if (VEC) {
        if (t0 < T) {
#pragma unroll
          for (int q = 0; q < SPT / 4; ++q) {
            const float4 n4 = __ldg(reinterpret_cast<const float4*>(nz + t0) + q);
            nzv[4 * q + 0] = n4.x; nzv[4 * q + 1] = n4.y; nzv[4 * q + 2] = n4.z; nzv[4 * q + 3] = n4.w;
          }
        } else {
#pragma unroll
          for (int k = 0; k < SPT; ++k) nzv[k] = 0.0f;
        }
      } else {
#pragma unroll
        for (int k = 0; k < SPT; ++k) nzv[k] = (t0 + k) < T ? __ldg(nz + t0 + k) : 0.0f;
      }
#if !IAS_ABL_NOBAR
      __syncthreads();
#endif
      acc_t acc1 = carry1 + (inc1 - tot1), acc2 = carry2 + (inc2 - tot2);
#pragma unroll
      for (int w = 0; w < (IAS_ABL_NOBAR ? 0 : NW); ++w) {
        const acc_t w1 = s_wsum[buf][0][w], w2 = s_wsum[buf][1][w];
        if (w < warp) {
          acc1 += w1;
          acc2 += w2;
        }
        carry1 += w1;
        carry2 += w2;
      }
      // ---- pass 2: oscillators, VCA gains, noise, mix --------------------------------------------------------
      float y[SPT];
      float lpeak = 0.0f;  // peak of this thread's samples of this tile (VEC: merged below only if the thread is live)
#pragma unroll
      for (int k = 0; k < SPT; k += 2) {
        const P2 sp = p2(srcs[k], srcs[k + 1]);
        const P2 u = p2_sub(sp, p2b(fj));
        const P2 um = p2_sub(sp, p2b(fj1));
        const P2 r = p2(fmaxf(p2lo(um), 0.0f), fmaxf(p2hi(um), 0.0f));
        acc1 += (acc_t)x1[k];
        const float a10 = (float)acc1;
        acc1 += (acc_t)x1[k + 1];
        const float a11 = (float)acc1;
        acc2 += (acc_t)x2[k];
        const float a20 = (float)acc2;
        acc2 += (acc_t)x2[k + 1];
        const float a21 = (float)acc2;
        const P2 arg1 = p2_add(p2(a10, a11), p2b(phase1));
        const P2 arg2 = p2_add(p2(a20, a21), p2b(phase2));
        const P2 g1 = p2_fma(r, p2b(r2.x), p2_fma(u, p2b(r1.w), p2b(r1.z)));
        const P2 g2 = p2_fma(r, p2b(r2.w), p2_fma(u, p2b(r2.z), p2b(r2.y)));
        const P2 g3 = p2_fma(r, p2b(r3.z), p2_fma(u, p2b(r3.y), p2b(r3.x)));
        const P2 yy = p2_fma(cos_arg_p2(arg1), g1,
                             p2_fma(squaresaw_core_p2(arg2, pk, shape), g2, p2_mul(p2(nzv[k], nzv[k + 1]), g3)));
        y[k] = p2lo(yy);
        y[k + 1] = p2hi(yy);
        if (VEC || (t0 + k) < T) lpeak = fmaxf(lpeak, fabsf(y[k]));
        if (VEC || (t0 + k + 1) < T) lpeak = fmaxf(lpeak, fabsf(y[k + 1]));
        if (DBG) {
          if ((t0 + k) < T) {
            A.phase_dbg[((size_t)b * 2 + 0) * T + t0 + k] = p2lo(arg1);
            A.phase_dbg[((size_t)b * 2 + 1) * T + t0 + k] = p2lo(arg2);
          }
          if ((t0 + k + 1) < T) {
            A.phase_dbg[((size_t)b * 2 + 0) * T + t0 + k + 1] = p2hi(arg1);
            A.phase_dbg[((size_t)b * 2 + 1) * T + t0 + k + 1] = p2hi(arg2);
          }
        }
      }
      // VEC: T % SPT == 0, so a thread is wholly live or wholly past the end of the clip.  Threads past the end of
      // the last tile still ran both passes (phase keeps accumulating, gains stay at the last control point): their
      // samples do not exist in the reference and must not reach max|mixed|.
      if (!VEC || t0 < T) tpeak = fmaxf(tpeak, lpeak);
      if (VEC) {
        if (t0 < T) {
#pragma unroll
          for (int q = 0; q < SPT / 4; ++q)
            reinterpret_cast<float4*>(out + t0)[q] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < SPT; ++k)
          if ((t0 + k) < T) out[t0 + k] = y[k];
      }
    }

    // ---- silent tail ---------------------------------------------------------------------------------------
    if ((long long)ntiles * TILE < T) {
      if (VEC) {
        float4* o4 = reinterpret_cast<float4*>(out);
        for (int i = ntiles * (TILE / 4) + tid; i < T / 4; i += NT) o4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        for (int i = ntiles * TILE + tid; i < T; i += NT) out[i] = 0.0f;
      }
    }

    // ---- per-voice peak, normalize_if_clipping -----------------------------------------------------------------
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) tpeak = fmaxf(tpeak, __shfl_xor_sync(0xffffffffu, tpeak, d));
    if (lane == 0) s_peak[warp] = tpeak;
    __syncthreads();  // also orders this CTA's global stores before the re-read below
    float pkv = s_peak[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) pkv = fmaxf(pkv, s_peak[w]);
    if (tid == 0 && A.peak) A.peak[b] = A.normalize == 2 ? (pkv > 1.0f ? vm::div(1.0f, pkv) : 1.0f) : pkv;
    if (A.normalize == 1 && pkv > 1.0f) {
      // x / peak, correctly rounded (Markstein step with r = RN(1/peak)); 4 independent 16-byte loads in flight per
      // thread so this second pass over the row runs at memory speed instead of one round trip per iteration
      const float rp = vm::div(1.0f, pkv);
      const int live = min(T, ntiles * TILE);  // the silent tail stays 0
      if (VEC) {
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
      } else {
        for (int i = tid; i < live; i += NT) out[i] = div_const(out[i], pkv, rp);
      }
    }
    // the barrier above separates this iteration's read of s_slot from the next iteration's write
  }
}

// ------------------------------------------------------------------------------------------------------------
// k_voice_audio_sp: the same arithmetic, software-pipelined across tiles
// ------------------------------------------------------------------------------------------------------------
// ncu of k_voice_audio (profiles/r2y): a warp spends 21 % of its time in pass 1 (FMA pipe only), 45 % in pass 2 (which
// needs 768 XU-pipe cycles per tile for 512 FMA-pipe cycles and is a chain of F2F -> reduction -> MUFU -> ... with
// little to issue while a result is in flight) and 20 % in the scan / barrier; with four warps per scheduler the two
// pipes are busy one after the other rather than together (FMA 5270 + XU 4096 cycles per round of four tiles, 9600
// measured).  Here pass 2 of tile t and pass 1 of tile t+1 are ONE straight-line block per thread: for every sample
// pair the oscillators of tile t are evaluated and the pitch chains of tile t+1 overwrite the increments just consumed
// (x1, x2, srcs are recycled in place), so the scheduler always has FMA-only chains to issue under the XU latencies.
// The increments stay fp32 in registers between the passes and are widened to fp64 where they are added (the
// classic kernel keeps 32 converted doubles = 64 registers alive across the barrier); the widening is two integer
// instructions instead of an XU-pipe F2F (increments are positive normal numbers).  Same operations on the same
// operands in the same order per sample: audio is bit-identical to k_voice_audio (tests/test_gpu_voice.py).
// Measured (profiles/r3_audio_experiments.md): 0.718 -> 0.706 ms, issue slots 54 -> 60 % busy.  The three switches
// below are the variants that were timed against it (F2F widening 0.726 ms, recomputed source coordinates 0.714,
// pitch chain first in source order 0.737); the last-tile threads past the end of the clip and the block's pass 1 of
// the tile after the last one compute on extrapolated control values -- finite or not, nothing reads them.
#ifndef IAS_SP_F2D_BITS
#define IAS_SP_F2D_BITS 1
#endif
#ifndef IAS_SP_PITCH_FIRST  // 1: source order pitch chain (t+1) before oscillators (t) within a sample pair
#define IAS_SP_PITCH_FIRST 0
#endif
#ifndef IAS_SP_RECOMPUTE_SRC  // 1: pass 2 recomputes the source coordinates instead of keeping 16 registers
#define IAS_SP_RECOMPUTE_SRC 0
#endif
__device__ __forceinline__ double widen_pos(float x) {
#if IAS_SP_F2D_BITS
  // fp32 -> fp64 of a positive normal number, exact: exponent re-biased (+896), mantissa moved up 29 bits
  const unsigned b = (unsigned)f2i(x);
  return __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29));
#else
  return (double)x;
#endif
}

// pitch chain of the sample pair (k, k+1) of the tile whose first sample of this thread is ft0 (interval fj)
template <bool CLAMP>
__device__ __forceinline__ void pitch_pair(int k, float ft0, float scale, float fj, float fj1, const float4 r0,
                                           const float4 r1, float midi1, float depth1, float midi2, float depth2,
                                           float sr, float rsr, P2& i1, P2& i2, P2& sp) {
  const P2 fi = p2_add(p2b(ft0), p2((float)k, (float)(k + 1)));
  sp = p2_fma(p2b(scale), fi, p2b(0.0f));
  const float s0 = p2lo(sp), s1 = p2hi(sp);
  const bool d0 = s0 >= fj1, d1 = s1 >= fj1;
  const P2 l1 = p2_sub(sp, p2(d0 ? fj1 : fj, d1 ? fj1 : fj));
  const P2 l0 = p2_sub(p2b(1.0f), l1);
  const P2 m1 = p2_fma(l0, p2(d0 ? r0.y : r0.x, d1 ? r0.y : r0.x), p2_mul(l1, p2(d0 ? r0.z : r0.y, d1 ? r0.z : r0.y)));
  const P2 m2 = p2_fma(l0, p2(d0 ? r1.x : r0.w, d1 ? r1.x : r0.w), p2_mul(l1, p2(d0 ? r1.y : r1.x, d1 ? r1.y : r1.x)));
  i1 = vco_increment_p2<CLAMP>(midi1, depth1, m1, sr, rsr);
  i2 = vco_increment_p2<CLAMP>(midi2, depth2, m2, sr, rsr);
}

struct SpVoice {  // per-voice constants of the merged block
  float midi1, depth1, phase1, midi2, depth2, phase2, pk, shape, scale, sr, rsr;
};

// pass 2 of the current tile + pass 1 of the next one, one straight-line block
template <int SPT, bool CLAMP, bool DBG>
__device__ __forceinline__ void merged_passes(float (&x1)[SPT], float (&x2)[SPT], float (&srcs)[SPT], float (&y)[SPT],
                                              const float (&nzv)[SPT], double& acc1, double& acc2, double& tot1,
                                              double& tot2, float& lpeak, const SpVoice& V, float ft0, float fj,
                                              float fj1, const float4 r1, const float4 r2, const float4 r3, float ft0n,
                                              float fjn, float fj1n, const float4 r0n, const float4 r1n,
                                              float* dbg1, float* dbg2, int t0, int T) {
#pragma unroll
  for (int k = 0; k < SPT; k += 2) {
#if IAS_SP_PITCH_FIRST
    P2 i1, i2, spn;
    pitch_pair<CLAMP>(k, ft0n, V.scale, fjn, fj1n, r0n, r1n, V.midi1, V.depth1, V.midi2, V.depth2, V.sr, V.rsr, i1, i2,
                      spn);
#endif
    // ---- tile t, pass 2 ----
#if IAS_SP_RECOMPUTE_SRC
    const P2 sp = p2_fma(p2b(V.scale), p2_add(p2b(ft0), p2((float)k, (float)(k + 1))), p2b(0.0f));
#else
    const P2 sp = p2(srcs[k], srcs[k + 1]);
#endif
    const P2 u = p2_sub(sp, p2b(fj));
    const P2 um = p2_sub(sp, p2b(fj1));
    const P2 r = p2(fmaxf(p2lo(um), 0.0f), fmaxf(p2hi(um), 0.0f));
    acc1 += widen_pos(x1[k]);
    const float a10 = (float)acc1;
    acc1 += widen_pos(x1[k + 1]);
    const float a11 = (float)acc1;
    acc2 += widen_pos(x2[k]);
    const float a20 = (float)acc2;
    acc2 += widen_pos(x2[k + 1]);
    const float a21 = (float)acc2;
    const P2 arg1 = p2_add(p2(a10, a11), p2b(V.phase1));
    const P2 arg2 = p2_add(p2(a20, a21), p2b(V.phase2));
    const P2 g1 = p2_fma(r, p2b(r2.x), p2_fma(u, p2b(r1.w), p2b(r1.z)));
    const P2 g2 = p2_fma(r, p2b(r2.w), p2_fma(u, p2b(r2.z), p2b(r2.y)));
    const P2 g3 = p2_fma(r, p2b(r3.z), p2_fma(u, p2b(r3.y), p2b(r3.x)));
    const P2 yy = p2_fma(cos_arg_p2(arg1), g1,
                         p2_fma(squaresaw_core_p2(arg2, V.pk, V.shape), g2, p2_mul(p2(nzv[k], nzv[k + 1]), g3)));
    y[k] = p2lo(yy);
    y[k + 1] = p2hi(yy);
    lpeak = fmaxf(lpeak, fmaxf(fabsf(y[k]), fabsf(y[k + 1])));
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
    // ---- tile t+1, pass 1: overwrites the increments consumed above ----
#if !IAS_SP_PITCH_FIRST
    P2 i1, i2, spn;
    pitch_pair<CLAMP>(k, ft0n, V.scale, fjn, fj1n, r0n, r1n, V.midi1, V.depth1, V.midi2, V.depth2, V.sr, V.rsr, i1, i2,
                      spn);
#endif
    x1[k] = p2lo(i1);
    x1[k + 1] = p2hi(i1);
    x2[k] = p2lo(i2);
    x2[k + 1] = p2hi(i2);
#if !IAS_SP_RECOMPUTE_SRC
    srcs[k] = p2lo(spn);
    srcs[k + 1] = p2hi(spn);
#endif
    tot1 += widen_pos(x1[k]);
    tot1 += widen_pos(x1[k + 1]);
    tot2 += widen_pos(x2[k]);
    tot2 += widen_pos(x2[k + 1]);
  }
}

// Requires T % SPT == 0 and 16-byte aligned rows (the VEC case of k_voice_audio; other calls take k_voice_audio).
template <int NT, int SPT, int MINB, bool DBG>
__global__ void __launch_bounds__(NT, MINB) k_voice_audio_sp(AudioArgs A) {
  constexpr int TILE = NT * SPT;
  constexpr int NW = NT / 32;
  __shared__ double s_wsum[2][2][NW];  // [buffer][vco][warp]
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
    float x1[SPT], x2[SPT], srcs[SPT];
    double tot1 = 0, tot2 = 0;
    float4 r1;
    if (ntiles > 0) {
      const float4 r0 = __ldg(rj + 0);
      r1 = __ldg(rj + 1);
#pragma unroll
      for (int k = 0; k < SPT; k += 2) {
        P2 i1, i2, sp;
        if (noclamp)
          pitch_pair<false>(k, ft0, V.scale, fj, fj1, r0, r1, V.midi1, V.depth1, V.midi2, V.depth2, V.sr, V.rsr, i1, i2, sp);
        else
          pitch_pair<true>(k, ft0, V.scale, fj, fj1, r0, r1, V.midi1, V.depth1, V.midi2, V.depth2, V.sr, V.rsr, i1, i2, sp);
        x1[k] = p2lo(i1); x1[k + 1] = p2hi(i1);
        x2[k] = p2lo(i2); x2[k + 1] = p2hi(i2);
        srcs[k] = p2lo(sp); srcs[k + 1] = p2hi(sp);
        tot1 += widen_pos(x1[k]);
        tot1 += widen_pos(x1[k + 1]);
        tot2 += widen_pos(x2[k]);
        tot2 += widen_pos(x2[k + 1]);
      }
    }
    for (int tile = 0; tile < ntiles; ++tile) {
      const int t0 = tile * TILE + tid * SPT;
      const int buf = tile & 1;
      // ---- block scan of tile t ----
      const double inc1 = warp_incl_scan(tot1, lane);
      const double inc2 = warp_incl_scan(tot2, lane);
      if (lane == 31) {
        s_wsum[buf][0][warp] = inc1;
        s_wsum[buf][1][warp] = inc2;
      }
      // gains and noise of tile t, pitch points of tile t+1: issued before the barrier
      const float4 r2 = __ldg(rj + 2);
      const float4 r3 = __ldg(rj + 3);
      float nzv[SPT];
      if (t0 < T) {
#pragma unroll
        for (int q = 0; q < SPT / 4; ++q) {
          const float4 n4 = __ldg(reinterpret_cast<const float4*>(nz + t0) + q);
          nzv[4 * q + 0] = n4.x; nzv[4 * q + 1] = n4.y; nzv[4 * q + 2] = n4.z; nzv[4 * q + 3] = n4.w;
        }
      } else {
#pragma unroll
        for (int k = 0; k < SPT; ++k) nzv[k] = 0.0f;
      }
      const float ft0n = add(ft0, (float)TILE);
      const int jn = min((int)mul(V.scale, fminf(ft0n, fTm1)), C - 1);
      const float fjn = (float)jn, fj1n = add(fjn, 1.0f);
      const float4* rjn = rec + (size_t)jn * 4;
      const float4 r0n = __ldg(rjn + 0);
      const float4 r1n = __ldg(rjn + 1);
      if (tile + 2 < ntiles) {  // L1 prefetch two tiles ahead (record) / one ahead (noise)
        const int jnn = min((int)mul(V.scale, fminf(add(ft0n, (float)TILE), fTm1)), C - 1);
        prefetch_l1(rec + (size_t)jnn * 4);
      }
      if (t0 + TILE < T) prefetch_l1(nz + t0 + TILE);
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
      float y[SPT];
      float lpeak = 0.0f;
      tot1 = 0;
      tot2 = 0;
#ifdef IAS_AUDIO_COUNT_NOCLAMP_ONLY  // tools/issue_model.py: object for instruction counting only (never shipped)
      if (true)
#else
      if (noclamp)
#endif
        merged_passes<SPT, false, DBG>(x1, x2, srcs, y, nzv, acc1, acc2, tot1, tot2, lpeak, V, ft0, fj, fj1, r1, r2, r3,
                                       ft0n, fjn, fj1n, r0n, r1n, dbg1, dbg2, t0, T);
      else
        merged_passes<SPT, true, DBG>(x1, x2, srcs, y, nzv, acc1, acc2, tot1, tot2, lpeak, V, ft0, fj, fj1, r1, r2, r3,
                                      ft0n, fjn, fj1n, r0n, r1n, dbg1, dbg2, t0, T);
      // a thread is wholly live or wholly past the end of the clip (T % SPT == 0); threads past the end still ran both
      // passes: their samples do not exist in the reference and must not reach max|mixed|
      if (t0 < T) {
        tpeak = fmaxf(tpeak, lpeak);
#pragma unroll
        for (int q = 0; q < SPT / 4; ++q)
          reinterpret_cast<float4*>(out + t0)[q] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
      }
      ft0 = ft0n; fj = fjn; fj1 = fj1n; rj = rjn; r1 = r1n;
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
    if (tid == 0 && A.peak) A.peak[b] = A.normalize == 2 ? (pkv > 1.0f ? vm::div(1.0f, pkv) : 1.0f) : pkv;
    if (A.normalize == 1 && pkv > 1.0f) {
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

struct VoiceWorkspace {
  float4* rec;
  float* ctrl;
  float* scratch;
  float* vconst;
  int* ntiles;
  int* order;
  int* counter;
};

size_t workspace_floats(int B, int C) {
  return (size_t)B * C * REC_FLOATS + (size_t)B * IAS_VOICE_NCONTROL * C + (size_t)B * 6 * C + (size_t)B * VC_COUNT +
         2 * (size_t)B + 4;
}

VoiceWorkspace carve(void* ws, int B, int C) {
  VoiceWorkspace w;
  float* f = reinterpret_cast<float*>(ws);
  w.rec = reinterpret_cast<float4*>(f);
  w.ctrl = f + (size_t)B * C * REC_FLOATS;
  w.scratch = w.ctrl + (size_t)B * IAS_VOICE_NCONTROL * C;
  w.vconst = w.scratch + (size_t)B * 6 * C;
  w.ntiles = reinterpret_cast<int*>(w.vconst + (size_t)B * VC_COUNT);
  w.order = w.ntiles + B;
  w.counter = w.order + B;
  return w;
}

// envelopes: the six ADSR rows into the workspace; modulation: LFOs, modulation matrix, records (reads those rows)
int launch_control(const float* params01, int B, int C, float cr, float eps, const float* ctrl_in, float* ctrl_out,
                   const VoiceWorkspace& w, cudaStream_t st, bool envelopes = true, bool modulation = true) {
  static const RangeTable ranges = make_range_table();
  if (envelopes) {
    ProfScope prof_(K_VOICE_ADSR, st);
    k_voice_adsr<<<dim3(B, 6), ADSR_THREADS, 0, st>>>(params01, B, C, cr, eps, ranges, w.scratch);
  }
  IAS_LAUNCH_CHECK("k_voice_adsr");
  if (modulation) {
    // 4 s clips: phase + output rows in shared memory (7*C floats); longer clips fall back to global scratch rows
    const size_t smem = (size_t)7 * C * sizeof(float);
    static unsigned long long attr_devs = 0;
    constexpr size_t SMEM_LIMIT = 56 * 1024;  // four CTAs per SM
    if (ias_first_use_on_device(attr_devs)) {
      IAS_CUDA(cudaFuncSetAttribute(k_voice_control<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
    }
    ProfScope prof_(K_VOICE_CONTROL, st);
    if (smem <= SMEM_LIMIT)
      k_voice_control<true><<<B, CTRL_THREADS, smem, st>>>(params01, B, C, cr, eps, ranges, ctrl_in, ctrl_out, w.ctrl,
                                                           w.scratch, w.vconst, w.rec);
    else
      k_voice_control<false><<<B, CTRL_THREADS, 0, st>>>(params01, B, C, cr, eps, ranges, ctrl_in, ctrl_out, w.ctrl,
                                                         w.scratch, w.vconst, w.rec);
  }
  IAS_LAUNCH_CHECK("k_voice_control");
  return IAS_OK;
}

int sm_count() {
  int dev = 0, v = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
  return v;
}

// persistent grid of the audio kernels; IAS_VOICE_GRID_PER_SM=<n> (tuning runs) launches fewer CTAs than fit
int audio_grid(int B, int minb) {
  static const int env = getenv("IAS_VOICE_GRID_PER_SM") ? atoi(getenv("IAS_VOICE_GRID_PER_SM")) : 0;
  return std::min(B, sm_count() * ((env > 0 && env < minb) ? env : minb));
}

// Shape of the audio kernel: threads per CTA, samples per thread per tile, resident CTAs per SM.
struct AudioShape {
  int nt, spt, ctas_per_sm;
  int sp;  // 0 classic, 1 software-pipelined (k_voice_audio_sp: needs the 128-bit path, otherwise the classic kernel runs)
};

template <int NT, int SPT, int MINB>
int launch_audio_shape(const AudioArgs& a, bool vec, bool dbg, cudaStream_t st) {
  const int grid = audio_grid(a.B, MINB);
  ProfScope prof_(K_VOICE_AUDIO, st);
  if (vec && !dbg)
    k_voice_audio<NT, SPT, MINB, true, false><<<grid, NT, 0, st>>>(a);
  else if (vec)
    k_voice_audio<NT, SPT, MINB, true, true><<<grid, NT, 0, st>>>(a);
  else if (!dbg)
    k_voice_audio<NT, SPT, MINB, false, false><<<grid, NT, 0, st>>>(a);
  else
    k_voice_audio<NT, SPT, MINB, false, true><<<grid, NT, 0, st>>>(a);
  return IAS_OK;
}

// IAS_VOICE_SHAPE=<threads>x<samples per thread>x<CTAs per SM> overrides the default shape (tuning runs only).
// Measured on B200 (profiles/r01f sweep, 1024 x 4 s): 128x16x4 0.84 ms, 256x16x2 0.86, 128x16x3 0.87, 256x8x3 0.93,
// 128x8x6 0.95, 128x8x7 1.02.  16 samples per thread halve the per-tile scan/barrier cost; it needs T % 16 == 0 for
// the 128-bit path (4 s: yes; 30 s: T % 8 == 0 only).
#ifndef IAS_AUDIO_SP_DEFAULT
#define IAS_AUDIO_SP_DEFAULT 1
#endif
AudioShape pick_shape(int T) {
  AudioShape s = (T % 16 == 0) ? AudioShape{128, 16, 4, IAS_AUDIO_SP_DEFAULT} : AudioShape{128, 8, 6, 0};
  if (const char* e = getenv("IAS_VOICE_SHAPE")) {  // "128x16x4" classic kernel, "p128x16x4" software-pipelined
    int nt = 0, spt = 0, c = 0;
    const int sp = e[0] == 'p' ? 1 : 0;
    if (sscanf(e + (sp ? 1 : 0), "%dx%dx%d", &nt, &spt, &c) == 3) s = AudioShape{nt, spt, c, sp};
  }
  return s;
}

template <int NT, int SPT, int MINB>
int launch_audio_sp(const AudioArgs& a, bool dbg, cudaStream_t st) {
  const int grid = audio_grid(a.B, MINB);
  ProfScope prof_(K_VOICE_AUDIO, st);
  if (!dbg)
    k_voice_audio_sp<NT, SPT, MINB, false><<<grid, NT, 0, st>>>(a);
  else
    k_voice_audio_sp<NT, SPT, MINB, true><<<grid, NT, 0, st>>>(a);
  return IAS_OK;
}

int launch_audio(const AudioArgs& a, const VoiceWorkspace& w, bool dbg, bool schedule, bool audio, cudaStream_t st) {
  const AudioShape s = pick_shape(a.T);
  const bool vec = (a.T % s.spt == 0) && ias_aligned16(a.noise) && ias_aligned16(a.audio);
  if (schedule) {
    ProfScope prof_(K_VOICE_SCHEDULE, st);
    // IAS_VOICE_RENDER_ALL=1 (tuning runs): silent tails are rendered like any other tile (same audio: their gains
    // are exactly 0), so every voice costs the same and a shape's throughput can be read without its load balance
    static const bool render_all_env = getenv("IAS_VOICE_RENDER_ALL") != nullptr;
    k_voice_schedule<<<1, SCHED_THREADS, 0, st>>>(a.vconst, a.B, a.T, a.scale, s.nt * s.spt,
                                                  (dbg || render_all_env) ? 1 : 0, w.ntiles,
                                                  w.order, w.counter);
  }
  IAS_LAUNCH_CHECK("k_voice_schedule");
  if (!audio) return IAS_OK;
#define IAS_SHAPE_SP(NT, SPT, MINB) \
  if (s.sp == 1 && vec && s.nt == NT && s.spt == SPT && s.ctas_per_sm == MINB) { launch_audio_sp<NT, SPT, MINB>(a, dbg, st); } else
  IAS_SHAPE_SP(128, 16, 4)
  IAS_SHAPE_SP(128, 16, 3)
#undef IAS_SHAPE_SP
#define IAS_SHAPE(NT, SPT, MINB) \
  if (s.nt == NT && s.spt == SPT && s.ctas_per_sm == MINB) { launch_audio_shape<NT, SPT, MINB>(a, vec, dbg, st); } else
  IAS_SHAPE(128, 8, 7)
  IAS_SHAPE(128, 8, 6)
  IAS_SHAPE(128, 16, 4)
  IAS_SHAPE(128, 16, 3)
  IAS_SHAPE(256, 8, 3)
  IAS_SHAPE(256, 16, 2)
  IAS_SHAPE(128, 16, 5)
  IAS_SHAPE(128, 12, 5)
  IAS_SHAPE(96, 16, 5)
  IAS_SHAPE(64, 16, 8)
  return set_err(IAS_ERR_UNSUPPORTED, "ias_voice_render: no audio kernel of shape %dx%dx%d", s.nt, s.spt, s.ctas_per_sm);
#undef IAS_SHAPE
  IAS_LAUNCH_CHECK("k_voice_audio");
  return IAS_OK;
}

}  // namespace
}  // namespace ias

using namespace ias;

extern "C" const char* ias_voice_param_name(int reg_index) {
  if (reg_index < 0 || reg_index >= vm::NROWS) return nullptr;
  return names().reg[reg_index].c_str();
}

extern "C" int ias_voice_sorted_index(int reg_index) {
  if (reg_index < 0 || reg_index >= vm::NROWS) return -1;
  return names().sorted_index[reg_index];
}

namespace {
int seed_params(int64_t first_sound_id, const int64_t* batch_idx_dev, int B, const uint8_t* frozen78_host,
                float* params01, uint8_t* is_train, ias_stream_t stream, const char* who) {
  IAS_REQUIRE(B > 0, IAS_ERR_INVALID, "%s: B=%d", who, B);
  IAS_REQUIRE(params01 != nullptr, IAS_ERR_INVALID, "%s: params01 is NULL", who);
  SeedTable tab;
  for (int i = 0; i < vm::NROWS; ++i) {
    tab.reg_of_sorted[i] = names().reg_of_sorted[i];
    tab.frozen[i] = frozen78_host ? frozen78_host[i] : 0;
  }
  {
    ProfScope prof_(K_SEED_PARAMS, as_stream(stream));
    k_seed_params<<<(B + SEED_THREADS - 1) / SEED_THREADS, SEED_THREADS, 0, as_stream(stream)>>>(
        (long long)first_sound_id, reinterpret_cast<const long long*>(batch_idx_dev), B, tab, params01, is_train);
  }
  IAS_LAUNCH_CHECK("k_seed_params");
  return IAS_OK;
}
}  // namespace

extern "C" int ias_voice_seed_params(int64_t first_sound_id, int B, const uint8_t* frozen78_host, float* params01,
                                     uint8_t* is_train, ias_stream_t stream) {
  return seed_params(first_sound_id, nullptr, B, frozen78_host, params01, is_train, stream, "ias_voice_seed_params");
}

extern "C" int ias_voice_seed_params_dev(const int64_t* batch_idx_dev, int B, const uint8_t* frozen78_host,
                                         float* params01, uint8_t* is_train, ias_stream_t stream) {
  IAS_REQUIRE(batch_idx_dev != nullptr, IAS_ERR_INVALID, "ias_voice_seed_params_dev: batch_idx_dev is NULL");
  return seed_params(0, batch_idx_dev, B, frozen78_host, params01, is_train, stream, "ias_voice_seed_params_dev");
}

extern "C" size_t ias_voice_workspace_bytes(int B, int T, int C) {
  (void)T;
  if (B <= 0 || C <= 0) return 0;
  return workspace_floats(B, C) * sizeof(float);
}

extern "C" int ias_voice_control(const float* params01, int B, int C, float control_rate, float eps, float* ctrl,
                                 void* workspace, size_t workspace_bytes, ias_stream_t stream) {
  IAS_REQUIRE(B > 0 && C > 1, IAS_ERR_INVALID, "ias_voice_control: B=%d C=%d", B, C);
  IAS_REQUIRE(params01 && ctrl, IAS_ERR_INVALID, "ias_voice_control: NULL pointer");
  IAS_REQUIRE(workspace && workspace_bytes >= ias_voice_workspace_bytes(B, 0, C), IAS_ERR_WORKSPACE,
              "ias_voice_control: workspace %zu < %zu bytes", workspace_bytes, ias_voice_workspace_bytes(B, 0, C));
  IAS_REQUIRE(ias_aligned16(workspace), IAS_ERR_INVALID, "ias_voice_control: workspace must be 16-byte aligned");
  VoiceWorkspace w = carve(workspace, B, C);
  return launch_control(params01, B, C, control_rate, eps, nullptr, ctrl, w, as_stream(stream));
}

extern "C" int ias_voice_render(const float* params01, const float* noise, int noise_rows, float* audio, float* peak,
                                int B, int T, int C, float sample_rate, float control_rate, float eps, int normalize,
                                const float* ctrl_in, float* phase_dbg, void* workspace, size_t workspace_bytes,
                                ias_stream_t stream) {
  return ias_voice_render_stages(params01, noise, noise_rows, audio, peak, B, T, C, sample_rate, control_rate, eps,
                                 normalize, ctrl_in, phase_dbg, workspace, workspace_bytes,
                                 IAS_VOICE_STAGE_CONTROL | IAS_VOICE_STAGE_AUDIO, stream);
}

extern "C" int ias_voice_render_stages(const float* params01, const float* noise, int noise_rows, float* audio,
                                       float* peak, int B, int T, int C, float sample_rate, float control_rate,
                                       float eps, int normalize, const float* ctrl_in, float* phase_dbg,
                                       void* workspace, size_t workspace_bytes, int stages, ias_stream_t stream) {
  const bool do_env = (stages & (IAS_VOICE_STAGE_CONTROL | IAS_VOICE_STAGE_ENVELOPES)) != 0;
  const bool do_mod = (stages & (IAS_VOICE_STAGE_CONTROL | IAS_VOICE_STAGE_MODULATION)) != 0;
  const bool do_control = do_env || do_mod, do_audio = (stages & IAS_VOICE_STAGE_AUDIO) != 0;
  IAS_REQUIRE(do_control || do_audio, IAS_ERR_INVALID, "ias_voice_render: stages=%d selects nothing", stages);
  IAS_REQUIRE(!(do_audio && do_env && !do_mod), IAS_ERR_INVALID,
              "ias_voice_render: stages=%d renders audio from envelopes without the modulation stage", stages);
  if (!do_audio) {  // control stage only: the audio-side pointers are not used
    noise_rows = noise_rows > 0 ? noise_rows : 1;
  }
  IAS_REQUIRE(B > 0 && T > 1 && C > 1 && noise_rows > 0, IAS_ERR_INVALID, "ias_voice_render: B=%d T=%d C=%d R=%d", B,
              T, C, noise_rows);
  IAS_REQUIRE(T < (1 << 24), IAS_ERR_UNSUPPORTED, "ias_voice_render: T=%d exceeds 2^24 samples", T);
  IAS_REQUIRE((long long)(T - 1) >= 16ll * (C - 1), IAS_ERR_UNSUPPORTED,
              "ias_voice_render: needs at least %d audio samples per control sample (T=%d C=%d)", 16, T, C);
  IAS_REQUIRE((!do_control || params01) && (!do_audio || (noise && audio)), IAS_ERR_INVALID,
              "ias_voice_render: NULL pointer");
  IAS_REQUIRE(sample_rate > 0.f && control_rate > 0.f, IAS_ERR_INVALID, "ias_voice_render: rates must be positive");
  IAS_REQUIRE(workspace && workspace_bytes >= ias_voice_workspace_bytes(B, T, C), IAS_ERR_WORKSPACE,
              "ias_voice_render: workspace %zu < %zu bytes", workspace_bytes, ias_voice_workspace_bytes(B, T, C));
  IAS_REQUIRE(ias_aligned16(workspace), IAS_ERR_INVALID, "ias_voice_render: workspace must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  VoiceWorkspace w = carve(workspace, B, C);
  if (do_control) {
    int rc = launch_control(params01, B, C, control_rate, eps, ctrl_in, nullptr, w, st, do_env, do_mod);
    if (rc) return rc;
  }
  AudioArgs a;
  a.rec = w.rec;
  a.vconst = w.vconst;
  a.noise = noise;
  a.ntiles = w.ntiles;
  a.order = w.order;
  a.counter = w.counter;
  a.audio = audio;
  a.peak = peak;
  a.phase_dbg = phase_dbg;
  a.B = B; a.T = T; a.C = C; a.noise_rows = noise_rows;
  a.scale = (float)(C - 1) / (float)(T - 1);
  a.sr = sample_rate;
  a.rsr = 1.0f / sample_rate;
  a.normalize = normalize;
  // the work queue (k_voice_schedule) belongs to the control stage: it depends on the control signals only
  return launch_audio(a, w, phase_dbg != nullptr, do_mod, do_audio, st);
}
