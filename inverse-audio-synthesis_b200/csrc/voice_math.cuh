// Per-sample arithmetic of the Voice render, written once for device and host.
//
// The device kernels in voice.cu and the host build in tests/shim/voice_host.cpp (a test-only
// sequential driver of the same functions, used to check rounding behaviour without a GPU) both include
// this header.  The torch CPU path of torchsynth (the parity target, oracle/voice.py) rounds after every
// fp32 tensor op, so every op that feeds a VCO phase is spelled out with a non-contracting primitive
// (mul/add/sub/div/fma below) in exactly the order torchsynth issues it.  See DESIGN.md "Voice numerics".
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define IAS_HD __host__ __device__ __forceinline__
#else
#define IAS_HD inline
#endif

namespace ias {
namespace vm {

// ---- IEEE-754 binary32 primitives, never contracted ---------------------------------------------------------
IAS_HD float mul(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
IAS_HD float add(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
IAS_HD float sub(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
IAS_HD float div(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
IAS_HD float fma(float a, float b, float c) { return fmaf(a, b, c); }

IAS_HD int f2i(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_int(f);
#else
  int i;
  memcpy(&i, &f, 4);
  return i;
#endif
}
IAS_HD float i2f(int i) {
#ifdef __CUDA_ARCH__
  return __int_as_float(i);
#else
  float f;
  memcpy(&f, &i, 4);
  return f;
#endif
}

// a / c for a compile-time-known positive constant c with rc = RN(1/c): Markstein's correction step gives the
// correctly rounded quotient (tests/test_voice_math.py checks it against IEEE division).
IAS_HD float div_const(float a, float c, float rc) {
  float q = mul(a, rc);
  float r = fma(-q, c, a);
  return fma(r, rc, q);
}

// 1 / x, correctly rounded (== div(1.0f, x))
IAS_HD float rcp(float x) {
#ifdef __CUDA_ARCH__
  return __frcp_rn(x);
#else
  return 1.0f / x;
#endif
}

// a / c for a divisor that is constant per voice, rc = rcp(c): two Markstein correction steps.  The first makes the
// quotient faithful for any normal c, the second rounds it correctly (tests/test_voice_math.py checks it against IEEE
// division over the ranges the ADSR ramps use).  Requires c, rc and the quotient to be normal numbers.
IAS_HD float div_pre(float a, float c, float rc) {
  float q = mul(a, rc);
  q = fma(fma(-q, c, a), rc, q);
  return fma(fma(-q, c, a), rc, q);
}

#define IAS_PI_F 3.14159274101257324f       /* (float)math.pi       */
#define IAS_TWO_PI_F 6.28318548202514648f   /* (float)(2 * math.pi) */

// ---- exp2: bit-exact restatement of the SLEEF 1.0-ULP single precision exp2 that torch's CPU vectorised
// path evaluates (Sleef_exp2f*_u10, FMA variant): round-to-nearest split, degree-6 FMA Horner, exponent add.
IAS_HD float exp2_poly(float s) {
  float u = +0.1535920892e-3f;
  u = fma(u, s, +0.1339262701e-2f);
  u = fma(u, s, +0.9618384764e-2f);
  u = fma(u, s, +0.5550347269e-1f);
  u = fma(u, s, +0.2402264476e+0f);
  u = fma(u, s, +0.6931471825e+0f);
  u = fma(u, s, 1.0f);
  return u;
}

// Valid for |d| < 100 (normal results): the pitch path, d in [-5.75, 12.9].
IAS_HD float exp2_fast(float d) {
  const float magic = 12582912.0f;  // 1.5 * 2^23: (d + magic) - magic == rint(d), low mantissa bits hold the integer
  float t = add(d, magic);
  int q = f2i(t) - 0x4B400000;
  float s = sub(d, sub(t, magic));
  return i2f(f2i(exp2_poly(s)) + (q << 23));
}

// Full-range version (parameter curves: exp2(log2(u)/curve) reaches -inf).
IAS_HD float exp2_full(float d) {
  if (d >= 128.0f) return INFINITY;
  if (d < -150.0f) return 0.0f;
  float qf = rintf(d);
  int q = (int)qf;
  float u = exp2_poly(sub(d, qf));
  float a = i2f(((q >> 1) + 127) << 23);
  float b = i2f(((q - (q >> 1)) + 127) << 23);
  return mul(mul(u, a), b);
}

// ---- pow: bit-exact restatement of the SLEEF 1.0-ULP single precision pow that torch's CPU vectorised path
// evaluates (Sleef_powf*_u10, FMA build): float-float log (atanh series on m in [0.75,1.5)), float-float product
// with the exponent, float-float exp.  Verified bit for bit against torch.pow (tests/test_voice_math.py).
struct F2 {
  float x, y;
};
IAS_HD F2 f2(float x, float y) {
  F2 r;
  r.x = x;
  r.y = y;
  return r;
}
IAS_HD F2 df_add2_f_f(float x, float y) {
  float s = add(x, y), v = sub(s, x);
  return f2(s, add(sub(x, sub(s, v)), sub(y, v)));
}
IAS_HD F2 df_add2_f2_f(F2 x, float y) {
  float s = add(x.x, y), v = sub(s, x.x);
  float t = add(sub(x.x, sub(s, v)), sub(y, v));
  return f2(s, add(t, x.y));
}
IAS_HD F2 df_add_f2_f2(F2 x, F2 y) {
  float s = add(x.x, y.x);
  return f2(s, add(add(add(sub(x.x, s), y.x), x.y), y.y));
}
IAS_HD F2 df_add_f_f2(float x, F2 y) {
  float s = add(x, y.x);
  return f2(s, add(add(sub(x, s), y.x), y.y));
}
IAS_HD F2 df_add2_f2_f2(F2 x, F2 y) {
  float s = add(x.x, y.x), v = sub(s, x.x);
  float t = add(sub(x.x, sub(s, v)), sub(y.x, v));
  return f2(s, add(t, add(x.y, y.y)));
}
IAS_HD F2 df_mul_f2_f(F2 x, float y) {
  float s = mul(x.x, y);
  return f2(s, fma(x.y, y, fma(x.x, y, -s)));
}
IAS_HD F2 df_mul_f2_f2(F2 x, F2 y) {
  float s = mul(x.x, y.x);
  return f2(s, fma(x.x, y.y, fma(x.y, y.x, fma(x.x, y.x, -s))));
}
IAS_HD F2 df_squ(F2 x) {
  float s = mul(x.x, x.x);
  return f2(s, fma(add(x.x, x.x), x.y, fma(x.x, x.x, -s)));
}
IAS_HD F2 df_div(F2 n, F2 d) {
  float t = rcp(d.x);
  float s = mul(n.x, t);
  float u = fma(t, n.x, -s);
  float v = fma(-d.y, t, fma(-d.x, t, 1.0f));
  return f2(s, fma(s, v, fma(n.y, t, u)));
}

IAS_HD F2 sleef_logk(float d) {  // d > 0
  const bool tiny = d < 1.17549435e-38f;
  if (tiny) d = mul(d, 1.8446744073709552e19f);
  int e = ((f2i(mul(d, 1.0f / 0.75f)) >> 23) & 0xff) - 0x7f;
  float m = i2f(f2i(d) - (e << 23));
  if (tiny) e -= 64;
  F2 x = df_div(df_add2_f_f(-1.0f, m), df_add2_f_f(1.0f, m));
  F2 x2 = df_squ(x);
  float t = 0.240320354700088500976562f;
  t = fma(t, x2.x, 0.285112679004669189453125f);
  t = fma(t, x2.x, 0.400007992982864379882812f);
  F2 c = f2(0.66666662693023681640625f, 3.69183861259614332084311e-09f);
  F2 s = df_mul_f2_f(f2(0.69314718246459960938f, -1.904654323148236017e-09f), (float)e);
  s = df_add_f2_f2(s, f2(mul(x.x, 2.0f), mul(x.y, 2.0f)));
  s = df_add_f2_f2(s, df_mul_f2_f2(df_mul_f2_f2(x2, x), df_add2_f2_f2(df_mul_f2_f(x2, t), c)));
  return s;
}

IAS_HD float sleef_expk(F2 d) {
  float qf = rintf(mul(add(d.x, d.y), 1.442695040888963407359924681001892137426645954152985934135449406931f));
  int q = (int)qf;
  F2 s = df_add2_f2_f(d, mul(qf, -0.693145751953125f));
  s = df_add2_f2_f(s, mul(qf, -1.428606765330187045e-06f));
  {  // normalise
    float n = add(s.x, s.y);
    s = f2(n, add(sub(s.x, n), s.y));
  }
  float u = 0.00136324646882712841033936f;
  u = fma(u, s.x, 0.00836596917361021041870117f);
  u = fma(u, s.x, 0.0416710823774337768554688f);
  u = fma(u, s.x, 0.166665524244308471679688f);
  u = fma(u, s.x, 0.499999850988388061523438f);
  F2 t = df_add_f2_f2(s, df_mul_f2_f(df_squ(s), u));
  t = df_add_f_f2(1.0f, t);
  u = add(t.x, t.y);
  float a = i2f(((q >> 1) + 127) << 23);
  float b = i2f(((q - (q >> 1)) + 127) << 23);
  u = mul(mul(u, a), b);
  if (d.x < -104.0f) u = 0.0f;
  return u;
}

// One out-of-line copy on the device: the control kernel calls it from 18 ramps + 10 LFO weights, and 28 inlined
// copies (~300 instructions each) do not fit the instruction cache.
#if defined(__CUDACC__) && defined(IAS_POW_NOINLINE)
__host__ __device__ __noinline__ float pow_sleef_core(float r, float a) {
#else
IAS_HD float pow_sleef_core(float r, float a) {
#endif
  return sleef_expk(df_mul_f2_f(sleef_logk(r), a));
}

IAS_HD float pow_sleef(float r, float a) {  // r >= 0, a > 0 (ADSR ramps, LFO mode weights)
  if (r == 1.0f) return 1.0f;
  if (r == 0.0f) return 0.0f;
  return pow_sleef_core(r, a);
}

// ---- correctly rounded (double evaluation, one final rounding) stand-ins for the SLEEF functions that are not
// restated bit for bit.  They agree with torch's CPU result except where SLEEF itself is not correctly rounded
// (measured: log2 0.05 %, log10 0.06 %, cos 4.9 % of arguments, by 1 ulp).
IAS_HD float log2_cr(float x) { return (float)log2((double)x); }
IAS_HD float log10_cr(float x) { return (float)log10((double)x); }
IAS_HD float cos_cr(float x) { return (float)cos((double)x); }

// ---- ModuleParameterRange.from_0to1 (torchsynth parameter.py; oracle/voice.py from_0to1) --------------------
struct ParamRange {
  float lo;     // (float)minimum
  float scale;  // (float)(maximum - minimum), or (float)((maximum - minimum) / 2) when symmetric
  float curve;
  int symmetric;
};

IAS_HD float from_0to1(float u, const ParamRange& r) {
  if (!r.symmetric) {
    if (r.curve != 1.0f) u = exp2_full(div(log2_cr(u), r.curve));
    return add(r.lo, mul(r.scale, u));
  }
  float dist = sub(mul(2.0f, u), 1.0f);
  float sg = dist > 0.0f ? 1.0f : (dist < 0.0f ? -1.0f : 0.0f);
  float shaped = mul(sg, exp2_full(div(log2_cr(fabsf(dist)), r.curve)));
  return add(r.lo, mul(r.scale, add(shaped, 1.0f)));
}

// Registration-order row indices of the 78 parameters (ias_b200.h: params01[78][B]).
enum Row {
  KEY_MIDI_F0 = 0, KEY_DURATION = 1,
  ADSR1 = 2, ADSR2 = 7,            // + {attack, decay, sustain, release, alpha}
  LFO1 = 12, LFO2 = 20,            // + {frequency, mod_depth, initial_phase, sin, tri, saw, rsaw, sqr}
  LFO1_AMP = 28, LFO2_AMP = 33, LFO1_RATE = 38, LFO2_RATE = 43,
  MODM = 48,                       // + input * 5 + output
  VCO1 = 68,                       // + {tuning, mod_depth, initial_phase}
  VCO2 = 71,                       // + {tuning, mod_depth, initial_phase, shape}
  MIX = 75,                        // + {vco_1, vco_2, noise}
  NROWS = 78
};

struct RangeTable {
  ParamRange r[NROWS];
};

inline RangeTable make_range_table() {
  struct D { double lo, hi; float curve; int sym; };
  const double PI = 3.14159265358979323846;
  const D key[2] = {{0.0, 127.0, 1.0f, 0}, {0.01, 4.0, 0.5f, 0}};
  const D adsr[5] = {{0.0, 2.0, 0.5f, 0}, {0.0, 2.0, 0.5f, 0}, {0.0, 1.0, 1.0f, 0}, {0.0, 5.0, 0.5f, 0}, {0.1, 6.0, 1.0f, 0}};
  const D lfo[8] = {{0.0, 20.0, 0.25f, 0}, {-10.0, 20.0, 0.5f, 1}, {-PI, PI, 1.0f, 0}, {0, 1, 1.0f, 0}, {0, 1, 1.0f, 0},
                    {0, 1, 1.0f, 0}, {0, 1, 1.0f, 0}, {0, 1, 1.0f, 0}};
  const D vco[4] = {{-24.0, 24.0, 1.0f, 0}, {-96.0, 96.0, 0.2f, 1}, {-PI, PI, 1.0f, 0}, {0.0, 1.0, 1.0f, 0}};
  const D mix[3] = {{0, 1, 1.0f, 0}, {0, 1, 1.0f, 0}, {0, 1, 0.1f, 0}};
  D all[NROWS];
  int k = 0;
  for (int i = 0; i < 2; ++i) all[k++] = key[i];
  for (int m = 0; m < 2; ++m) for (int i = 0; i < 5; ++i) all[k++] = adsr[i];
  for (int m = 0; m < 2; ++m) for (int i = 0; i < 8; ++i) all[k++] = lfo[i];
  for (int m = 0; m < 4; ++m) for (int i = 0; i < 5; ++i) all[k++] = adsr[i];
  for (int i = 0; i < 20; ++i) all[k++] = D{0.0, 1.0, 0.5f, 0};
  for (int i = 0; i < 3; ++i) all[k++] = vco[i];
  for (int i = 0; i < 4; ++i) all[k++] = vco[i];
  for (int i = 0; i < 3; ++i) all[k++] = mix[i];
  RangeTable t;
  for (int i = 0; i < NROWS; ++i) {
    t.r[i].lo = (float)all[i].lo;
    t.r[i].scale = all[i].sym ? (float)((all[i].hi - all[i].lo) / 2.0) : (float)(all[i].hi - all[i].lo);
    t.r[i].curve = all[i].curve;
    t.r[i].symmetric = all[i].sym;
  }
  return t;
}

// ---- ADSR (torchsynth module.py ADSR; oracle/voice.py _adsr) -------------------------------------------------
// One ramp of the envelope: r(n) = min((max(n - start, 0) + eps) / dur + eps, 1), inverted to 1 - r when `inverse` and
// dur > 0, then raised to alpha.  r is non-decreasing in n, so it is constant before `start` and exactly 1 from some
// integer n on; both thresholds are found per voice (with the very same fp32 expression) so that the per-point
// evaluation skips the division and the pow outside the ramp's active span.
struct Ramp {
  float dur, rdur, start;
  float sat_from;  // first integer n (as float) with r(n) == 1; >= 1e30 if none was found below the search limit
  float pre, post; // value of the finished ramp (after inversion and pow) for n <= start / n >= sat_from
  int fast_div;    // dur normal and well inside the exponent range: division through rdur = rcp(dur)
  int inverse;
};

IAS_HD float ramp_raw(const Ramp& p, float n, float eps) {
  float r = fmaxf(sub(n, p.start), 0.0f);
  r = add(r, eps);
  r = p.fast_div ? div_pre(r, p.dur, p.rdur) : div(r, p.dur);
  return fminf(add(r, eps), 1.0f);
}

IAS_HD float ramp_finish(const Ramp& p, float r, float alpha) {
  if (p.inverse && p.dur > 0.0f) r = sub(1.0f, r);
  return pow_sleef(r, alpha);
}

IAS_HD Ramp ramp_setup(float dur, float start, bool inverse, float alpha, float eps, int n_limit) {
  Ramp p;
  p.dur = dur;
  p.start = start;
  p.inverse = inverse ? 1 : 0;
  p.fast_div = (dur > 1e-20f && dur < 1e20f) ? 1 : 0;
  p.rdur = p.fast_div ? rcp(dur) : 0.0f;
  const float r0 = ramp_raw(p, 0.0f < start ? 0.0f : start, eps);  // max(n - start, 0) == 0 for every n <= start
  p.pre = ramp_finish(p, r0, alpha);
  p.post = ramp_finish(p, 1.0f, alpha);
  if (r0 == 1.0f) {  // zero (or tiny) duration: saturated everywhere
    p.sat_from = 0.0f;
    return p;
  }
  // first integer n with r(n) == 1: it lies above `start`; begin at the real-valued estimate and walk (r is monotone)
  float est = ceilf(add(start, dur));
  int n0 = est < 0.0f ? 0 : (est > (float)n_limit ? n_limit : (int)est);
  while (n0 > 0 && (float)(n0 - 1) > start && ramp_raw(p, (float)(n0 - 1), eps) == 1.0f) --n0;
  while (n0 < n_limit && ramp_raw(p, (float)n0, eps) != 1.0f) ++n0;
  p.sat_from = n0 < n_limit ? (float)n0 : 1e30f;
  return p;
}

IAS_HD float ramp_eval(const Ramp& p, float n, float alpha, float eps) {
  if (n >= p.sat_from) return p.post;
  if (n <= p.start) return p.pre;
  return ramp_finish(p, ramp_raw(p, n, eps), alpha);
}

struct Adsr {
  Ramp a, d, r;
  float sustain, one_minus_sustain, alpha;
};

// v[] = {attack, decay, sustain, release, alpha} already through from_0to1; n_limit = number of control points.
// `which` selects the part to fill (0 attack + scalars, 1 decay, 2 release) so three threads can share one envelope.
IAS_HD void adsr_setup_part(Adsr& p, int which, const float* v, float note_on, float cr, float eps, int n_limit) {
  const float new_attack = fminf(v[0], note_on);
  const float alpha = v[4];
  if (which == 0) {
    p.sustain = v[2];
    p.one_minus_sustain = sub(1.0f, v[2]);
    p.alpha = alpha;
    p.a = ramp_setup(mul(new_attack, cr), 0.0f, false, alpha, eps, n_limit);
  } else if (which == 1) {
    const float new_decay = fminf(fmaxf(sub(note_on, v[0]), 0.0f), v[1]);
    p.d = ramp_setup(mul(new_decay, cr), mul(new_attack, cr), true, alpha, eps, n_limit);
  } else {
    p.r = ramp_setup(mul(v[3], cr), mul(note_on, cr), true, alpha, eps, n_limit);
  }
}

IAS_HD Adsr adsr_setup(const float* v, float note_on, float cr, float eps, int n_limit) {
  Adsr p;
  for (int which = 0; which < 3; ++which) adsr_setup_part(p, which, v, note_on, cr, eps, n_limit);
  return p;
}

IAS_HD float adsr_eval(const Adsr& p, float n, float eps) {
  float a = ramp_eval(p.a, n, p.alpha, eps);
  float d = ramp_eval(p.d, n, p.alpha, eps);
  float dk = add(mul(p.one_minus_sustain, d), p.sustain);
  float r = ramp_eval(p.r, n, p.alpha, eps);
  return mul(mul(a, dk), r);
}

// ---- LFO (torchsynth module.py LFO; oracle/voice.py _lfo) ----------------------------------------------------
struct Lfo {
  float frequency, mod_depth, initial_phase;
  float w[5];  // normalised mode weights: sin, tri, saw, rsaw, sqr
};

// v[] = {frequency, mod_depth, initial_phase, sin, tri, saw, rsaw, sqr} already through from_0to1.
IAS_HD Lfo lfo_setup(const float* v) {
  Lfo l;
  l.frequency = v[0];
  l.mod_depth = v[1];
  l.initial_phase = v[2];
  float m[5];
  for (int i = 0; i < 5; ++i) m[i] = pow_sleef(v[3 + i], 2.718281828f);
  // torch.sum over a contiguous length-5 row: 4 ILP lanes, remainder folded into lane 0 first
  float s = add(add(add(add(m[0], m[4]), m[1]), m[2]), m[3]);
  for (int i = 0; i < 5; ++i) l.w[i] = div(m[i], s);
  return l;
}

// The same in two steps, so the five pows can be spread over threads: the caller fills l.w[i] = v[3+i] ** e first.
IAS_HD void lfo_finish(Lfo& l, const float* v) {
  l.frequency = v[0];
  l.mod_depth = v[1];
  l.initial_phase = v[2];
  float s = add(add(add(add(l.w[0], l.w[4]), l.w[1]), l.w[2]), l.w[3]);
  float m[5];
  for (int i = 0; i < 5; ++i) m[i] = l.w[i];
  for (int i = 0; i < 5; ++i) l.w[i] = div(m[i], s);
}

// phase increment of one control sample: 2*pi*max(frequency + mod_depth*mod, 0) / control_rate
IAS_HD float lfo_increment(const Lfo& l, float mod, float cr) {
  float f = fmaxf(add(l.frequency, mul(l.mod_depth, mod)), 0.0f);
  return div(mul(IAS_TWO_PI_F, f), cr);
}
// same with rcr = rcp(cr) hoisted by the caller (cr is a launch constant)
IAS_HD float lfo_increment(const Lfo& l, float mod, float cr, float rcr) {
  float f = fmaxf(add(l.frequency, mul(l.mod_depth, mod)), 0.0f);
  return div_pre(mul(IAS_TWO_PI_F, f), cr, rcr);
}

IAS_HD float lfo_shapes_mix(const Lfo& l, float arg) {
  float c = cos_cr(add(arg, IAS_PI_F));
  float sq = c > 0.0f ? 1.0f : (c < 0.0f ? -1.0f : 0.0f);
  c = mul(add(c, 1.0f), 0.5f);
  sq = mul(add(sq, 1.0f), 0.5f);
  float m = fmodf(arg, IAS_TWO_PI_F);
  if (m != 0.0f && m < 0.0f) m = add(m, IAS_TWO_PI_F);
  float saw = div_const(m, IAS_TWO_PI_F, 0.159154936671257019f);  // RN(1 / (float)(2 pi))
  float rsaw = sub(1.0f, saw);
  float tri = mul(2.0f, saw);
  if (tri > 1.0f) tri = sub(2.0f, tri);
  float acc = mul(l.w[0], c);  // matmul [1,5]x[5,C]: FMA chain in k order
  acc = fma(l.w[1], tri, acc);
  acc = fma(l.w[2], saw, acc);
  acc = fma(l.w[3], rsaw, acc);
  acc = fma(l.w[4], sq, acc);
  return acc;
}

// ---- Modulation matrix (torchsynth ModulationMixer; oracle/voice.py _mod_matrix) -----------------------------
struct ModMatrix {
  float w[5][4];  // [output][input], rows normalised
};

// v[20] input-major ({adsr_1, adsr_2, lfo_1, lfo_2} x 5 outputs), already through from_0to1.
IAS_HD ModMatrix modmatrix_setup(const float* v) {
  ModMatrix mm;
  for (int o = 0; o < 5; ++o) {
    float s = add(add(add(v[0 * 5 + o], v[1 * 5 + o]), v[2 * 5 + o]), v[3 * 5 + o]);
    for (int i = 0; i < 4; ++i) mm.w[o][i] = div(v[i * 5 + o], s);
  }
  return mm;
}

IAS_HD float modmatrix_out(const ModMatrix& mm, int o, float a1, float a2, float l1, float l2) {
  float acc = mul(mm.w[o][0], a1);
  acc = fma(mm.w[o][1], a2, acc);
  acc = fma(mm.w[o][2], l1, acc);
  acc = fma(mm.w[o][3], l2, acc);
  return acc;
}

// ---- per-voice constants handed from the control stage to the audio stage ------------------------------------
enum VoiceConst {
  VC_MIDI1 = 0, VC_DEPTH1, VC_PHASE1, VC_MIDI2, VC_DEPTH2, VC_PHASE2,
  VC_PK,      // pi * partials_constant * 0.5 * 2*log2(e): tanh argument scale folded with tanh_from_scaled's
  VC_SHAPE, VC_GAIN2,  // shape, 1 - shape/2
  VC_LEVEL1, VC_LEVEL2, VC_LEVEL3,
  VC_SILENT_FROM,  // control index from which all three amplitude signals are exactly 0 to the end (as float)
  VC_NOCLAMP,      // 1.0 if midi + depth*mod stays inside (0.5, 126.5) for both VCOs over the whole clip: the clamp
                   // to [0, 127] is then the identity and the audio stage skips it
  VC_COUNT = 16
};

IAS_HD float midi_to_hz(float m) { return mul(440.0f, exp2_fast(div_const(sub(m, 69.0f), 12.0f, 0.0833333358168601989746f))); }

// P[78]: parameters through from_0to1, registration order.
IAS_HD void voice_constants(const float* P, float* vc) {
  float midi_f0 = P[KEY_MIDI_F0];
  vc[VC_MIDI1] = add(midi_f0, P[VCO1 + 0]);
  vc[VC_DEPTH1] = P[VCO1 + 1];
  vc[VC_PHASE1] = P[VCO1 + 2];
  vc[VC_MIDI2] = add(midi_f0, P[VCO2 + 0]);
  vc[VC_DEPTH2] = P[VCO2 + 1];
  vc[VC_PHASE2] = P[VCO2 + 2];
  float max_f0 = midi_to_hz(add(midi_f0, fmaxf(P[VCO2 + 1], 0.0f)));
  float partials = div(12000.0f, mul(max_f0, log10_cr(max_f0)));
  vc[VC_PK] = mul(mul(mul(IAS_PI_F, partials), 0.5f), 2.88539008177792681472f);
  vc[VC_SHAPE] = P[VCO2 + 3];
  vc[VC_GAIN2] = sub(1.0f, mul(P[VCO2 + 3], 0.5f));
  vc[VC_LEVEL1] = P[MIX + 0];
  vc[VC_LEVEL2] = P[MIX + 1];
  vc[VC_LEVEL3] = P[MIX + 2];
  for (int i = VC_LEVEL3 + 1; i < VC_COUNT; ++i) vc[i] = 0.0f;
  vc[VC_SILENT_FROM] = 1e30f;  // filled in by the control stage once the signals are known
}

// ---- audio rate ----------------------------------------------------------------------------------------------
// nn.Upsample(mode="linear", align_corners=True) source index/weights of output sample i (torch CPU formula:
// scale = float(C-1)/float(T-1); src = scale*i; out = fma(l0, x[i0], l1*x[i1])).
IAS_HD void upsample_coords(int i, float scale, int C, int& i0, int& i1, float& l0, float& l1) {
  float src = mul(scale, (float)i);
  i0 = (int)src;
  if (i0 > C - 1) i0 = C - 1;
  l1 = fminf(fmaxf(sub(src, (float)i0), 0.0f), 1.0f);
  l0 = sub(1.0f, l1);
  i1 = i0 < C - 1 ? i0 + 1 : i0;
}
IAS_HD float upsample_mix(float v0, float v1, float l0, float l1) { return fma(l0, v0, mul(l1, v1)); }

// Same coordinates without int<->float conversions, for a thread whose samples all fall in control intervals j or
// j+1: fi = float(i), fj = float(j).  Returns d = i0 - j (0 or 1) and the two weights.
IAS_HD bool upsample_coords_f(float fi, float scale, float fj, float& l0, float& l1) {
  const float src = mul(scale, fi);
  const float fj1 = add(fj, 1.0f);
  const bool d = src >= fj1;
  l1 = sub(src, d ? fj1 : fj);  // in [0,1) by construction: j = floor(src) of the thread's first sample
  l0 = sub(1.0f, l1);
  return d;
}

// VCO phase increment of one audio sample: 2*pi*hz(clamp(midi + depth*mod, 0, 127)) / sample_rate
IAS_HD float vco_increment(float midi, float depth, float mod, float sr, float rsr) {
  float m = fminf(fmaxf(add(midi, mul(depth, mod)), 0.0f), 127.0f);
  return div_const(mul(IAS_TWO_PI_F, midi_to_hz(m)), sr, rsr);
}

// a = (n + f) * pi with f in [-0.5, 0.5]; accurate for |a| < 2^22 * pi.  n = rint(a * C1) from one FMA against the
// 1.5 * 2^23 magic constant; the second FMA forms a*C1 - n with a single rounding (relative to the small result, so
// accuracy near the zeros of sin is kept), the third adds the 1/pi tail.
IAS_HD void reduce_half_turns(float a, float& f, int& n) {
  const float C1 = 0.318309873342514038f;                                            // (float)(1/pi)
  const float C2 = (float)(0.31830988618379067154 - (double)0.318309873342514038f);  // 1/pi - C1
  const float magic = 12582912.0f;
  float t = fma(a, C1, magic);
  n = f2i(t);  // low bit = parity of rint(a/pi)
  float nf = sub(t, magic);
  f = fma(a, C2, fma(a, C1, -nf));
}

// sin(pi*f) for |f| <= 0.5625 (reduce_half_turns can overshoot 0.5 by the ulp of a/pi for 30 s clips): odd degree-9
// polynomial, relative error <= 1.9e-7, so the zero crossings that SquareSawVCO amplifies through tanh stay
// accurate in absolute terms.
IAS_HD float sinpi_poly(float f) {
  const float u = mul(f, f);
  float p = 0.07634329050779343f;
  p = fma(p, u, -0.59761643409729f);
  p = fma(p, u, 2.5499696731567383f);
  p = fma(p, u, -5.1677045822143555f);
  p = fma(p, u, 3.141592502593994f);
  return mul(p, f);
}

// cos(pi*f) for f in [-0.5, 0.5], absolute error ~5e-7 (amplitude path only): the SFU cosine on the device.
IAS_HD float cospi_fast(float f) {
#ifdef __CUDA_ARCH__
  return __cosf(mul(f, IAS_PI_F));
#else
  return (float)cos(3.14159265358979323846 * (double)f);
#endif
}

// tanh(z) given u = 2*log2(e)*z:  1 - 2 / (exp2(u) + 1), absolute error ~1.5e-7; saturates correctly for huge |u|.
IAS_HD float tanh_from_scaled(float u) {
#ifdef __CUDA_ARCH__
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(u));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(add(e, 1.0f)));
  return fma(-2.0f, r, 1.0f);
#else
  return (float)tanh((double)u / 2.88539008177792681472);
#endif
}

IAS_HD void sincos_arg(float a, float& s, float& c) {
  float f;
  int n;
  reduce_half_turns(a, f, n);
  const int flip = (n & 1) << 31;
  s = i2f(f2i(sinpi_poly(f)) ^ flip);
  c = i2f(f2i(cospi_fast(f)) ^ flip);
}

IAS_HD float cos_arg(float a) {
  float f;
  int n;
  reduce_half_turns(a, f, n);
  return i2f(f2i(cospi_fast(f)) ^ ((n & 1) << 31));
}

// SquareSawVCO.oscillator: (1 - shape/2) * tanh(pi*k*sin(arg)/2) * (1 + shape*cos(arg)); pk = VC_PK (pre-scaled)
IAS_HD float squaresaw(float arg, float pk, float shape, float gain) {
  float s, c;
  sincos_arg(arg, s, c);
  float sq = tanh_from_scaled(mul(pk, s));
  return mul(mul(gain, sq), fma(shape, c, 1.0f));
}

// ---- amplitude path of the audio stage ------------------------------------------------------------------------
// The three VCA gains (vco_1_amp, vco_2_amp, noise_amp) are linear interpolations of control points j, j+1, j+2 for a
// thread whose samples start in control interval j.  With u = src - j (src = the reference's fp32 source coordinate,
// so its quantisation is reproduced) and r = max(u - 1, 0):
//     gain(u) = v0 + u*(v1 - v0) + r*((v2 - v1) - (v1 - v0))
// which equals the reference's l0*x[i0] + l1*x[i1] on both sides of the interval boundary up to fp32 rounding
// (<= 2e-7; amplitude path only, never feeds a phase).  The mixer level (and SquareSawVCO's 1 - shape/2) is folded in.
struct AmpLine {
  float c0, d0, dd;
};
IAS_HD AmpLine amp_line(float v0, float v1, float v2, float level) {
  AmpLine a;
  const float d0 = sub(v1, v0);
  a.c0 = mul(level, v0);
  a.d0 = mul(level, d0);
  a.dd = mul(level, sub(sub(v2, v1), d0));
  return a;
}
IAS_HD float amp_eval(const AmpLine& a, float u, float r) { return fma(r, a.dd, fma(u, a.d0, a.c0)); }

// SquareSawVCO.oscillator without its constant gain: tanh(pi*k*sin(arg)/2) * (1 + shape*cos(arg))
IAS_HD float squaresaw_core(float arg, float pk, float shape) {
  float s, c;
  sincos_arg(arg, s, c);
  return mul(tanh_from_scaled(mul(pk, s)), fma(shape, c, 1.0f));
}

// AudioMixer matmul [1,3]x[3,T]: FMA chain in k order
IAS_HD float mix3(float l1, float v1, float l2, float v2, float l3, float v3) {
  float acc = mul(l1, v1);
  acc = fma(l2, v2, acc);
  acc = fma(l3, v3, acc);
  return acc;
}

}  // namespace vm
}  // namespace ias
