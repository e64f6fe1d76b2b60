// TEST-ONLY host build of csrc/voice_math.cuh: drives the same per-sample functions the CUDA kernels use, one
// voice at a time on the CPU, so rounding behaviour can be checked against oracle/voice.py without a GPU.
// Built by tests/conftest.py with g++ -O2 -mfma -ffp-contract=off.  Never loaded by the product package.
#include <stdlib.h>
#include <vector>

#include "../../inverse-audio-synthesis_b200/csrc/voice_math.cuh"

using namespace ias::vm;

extern "C" {

void shim_exp2_fast(const float* x, float* y, long n) { for (long i = 0; i < n; ++i) y[i] = exp2_fast(x[i]); }
void shim_exp2_full(const float* x, float* y, long n) { for (long i = 0; i < n; ++i) y[i] = exp2_full(x[i]); }
void shim_div_const(const float* x, float c, float* y, long n) {
  float rc = 1.0f / c;
  for (long i = 0; i < n; ++i) y[i] = div_const(x[i], c, rc);
}
// a[i] / c[i] through the per-voice-constant path (rcp + two Markstein steps)
void shim_div_pre(const float* a, const float* c, float* y, long n) {
  for (long i = 0; i < n; ++i) y[i] = div_pre(a[i], c[i], rcp(c[i]));
}
void shim_sincos_arg(const float* x, float* s, float* c, long n) { for (long i = 0; i < n; ++i) sincos_arg(x[i], s[i], c[i]); }
void shim_cos_arg(const float* x, float* c, long n) { for (long i = 0; i < n; ++i) c[i] = cos_arg(x[i]); }

// params01[78][B] -> P[B][78]
void shim_from_0to1(const float* params01, int B, float* P) {
  RangeTable t = make_range_table();
  for (int b = 0; b < B; ++b)
    for (int r = 0; r < NROWS; ++r) P[(size_t)b * NROWS + r] = from_0to1(params01[(size_t)r * B + b], t.r[r]);
}

// control stage: ctrl[B][5][C], vconst[B][16], dbg[B][8][C] = adsr_1, adsr_2, lfo1_amp, lfo2_amp, lfo1_rate, lfo2_rate, lfo_1, lfo_2
void shim_control(const float* params01, int B, int C, float cr, float eps, float* ctrl, float* vconst, float* dbg) {
  RangeTable t = make_range_table();
  for (int b = 0; b < B; ++b) {
    float P[NROWS];
    for (int r = 0; r < NROWS; ++r) P[r] = from_0to1(params01[(size_t)r * B + b], t.r[r]);
    const int base[6] = {ADSR1, ADSR2, LFO1_AMP, LFO2_AMP, LFO1_RATE, LFO2_RATE};
    Adsr ad[6];
    for (int i = 0; i < 6; ++i) ad[i] = adsr_setup(&P[base[i]], P[KEY_DURATION], cr, eps, C);
    Lfo lf[2] = {lfo_setup(&P[LFO1]), lfo_setup(&P[LFO2])};
    ModMatrix mm = modmatrix_setup(&P[MODM]);
    voice_constants(P, vconst + (size_t)b * VC_COUNT);
    double acc[2] = {0.0, 0.0};
    for (int j = 0; j < C; ++j) {
      float n = (float)j;
      float e[6];
      for (int i = 0; i < 6; ++i) e[i] = adsr_eval(ad[i], n, eps);
      float l[2];
      for (int k = 0; k < 2; ++k) {
        acc[k] += (double)lfo_increment(lf[k], e[4 + k], cr, rcp(cr));
        float arg = add((float)acc[k], lf[k].initial_phase);
        l[k] = mul(lfo_shapes_mix(lf[k], arg), e[2 + k]);
      }
      for (int o = 0; o < 5; ++o) ctrl[((size_t)b * 5 + o) * C + j] = modmatrix_out(mm, o, e[0], e[1], l[0], l[1]);
      if (dbg) {
        for (int i = 0; i < 6; ++i) dbg[((size_t)b * 8 + i) * C + j] = e[i];
        dbg[((size_t)b * 8 + 6) * C + j] = l[0];
        dbg[((size_t)b * 8 + 7) * C + j] = l[1];
      }
    }
  }
}

// audio stage from ctrl/vconst: audio[B][T], peak[B], phase[B][2][T] (may be null)
void shim_audio(const float* ctrl, const float* vconst, const float* noise, int R, int B, int T, int C, float sr,
                int normalize, float* audio, float* peak, float* phase) {
  const float scale = (float)(C - 1) / (float)(T - 1);
  const float rsr = 1.0f / sr;
  for (int b = 0; b < B; ++b) {
    const float* vc = vconst + (size_t)b * VC_COUNT;
    const float* ct = ctrl + (size_t)b * 5 * C;
    const float* nz = noise + (size_t)(b % R) * T;
    float* out = audio + (size_t)b * T;
    double a1 = 0.0, a2 = 0.0;
    float pk = 0.0f;
    for (int i = 0; i < T; ++i) {
      int i0, i1;
      float l0, l1;
      upsample_coords(i, scale, C, i0, i1, l0, l1);
      float u[5];
      for (int s = 0; s < 5; ++s) u[s] = upsample_mix(ct[s * C + i0], ct[s * C + i1], l0, l1);
      a1 += (double)vco_increment(vc[VC_MIDI1], vc[VC_DEPTH1], u[0], sr, rsr);
      a2 += (double)vco_increment(vc[VC_MIDI2], vc[VC_DEPTH2], u[2], sr, rsr);
      float arg1 = add((float)a1, vc[VC_PHASE1]);
      float arg2 = add((float)a2, vc[VC_PHASE2]);
      float v1 = mul(cos_arg(arg1), u[1]);
      float v2 = mul(squaresaw(arg2, vc[VC_PK], vc[VC_SHAPE], vc[VC_GAIN2]), u[3]);
      float v3 = mul(nz[i], u[4]);
      float y = mix3(vc[VC_LEVEL1], v1, vc[VC_LEVEL2], v2, vc[VC_LEVEL3], v3);
      out[i] = y;
      pk = fmaxf(pk, fabsf(y));
      if (phase) {
        phase[((size_t)b * 2 + 0) * T + i] = arg1;
        phase[((size_t)b * 2 + 1) * T + i] = arg2;
      }
    }
    if (peak) peak[b] = pk;
    if (normalize && pk > 1.0f)
      for (int i = 0; i < T; ++i) out[i] = ias::vm::div(out[i], pk);
  }
}

}  // extern "C"
extern "C" void shim_pow(const float* x, const float* a, float* y, long n) { for (long i = 0; i < n; ++i) y[i] = pow_sleef(x[i], a[i]); }
