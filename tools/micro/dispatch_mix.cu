// Microbenchmark (B200): marginal dispatch cost of the instruction classes k_voice_audio mixes with packed fp32.
// Every mode runs 16 warps per scheduler of independent chains; cycles per iteration and scheduler are printed, so the
// difference to the "FFMA2 x8" line is what the added instructions cost the dispatch port.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dispatch_mix dispatch_mix.cu && ./dispatch_mix
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
  float r;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float ex2(float a) {
  float r;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ double f2d(float a) {
  double r;
  asm volatile("cvt.f64.f32 %0, %1;" : "=d"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ float d2f(double a) {
  float r;
  asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(r) : "d"(a));
  return r;
}
__device__ __forceinline__ double my_dadd(double a, double b) {
  double r;
  asm volatile("add.rn.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b));
  return r;
}
__device__ __forceinline__ float fsel(float a, float b, int p) {
  float r;
  asm volatile("{.reg .pred q; setp.ne.s32 q, %3, 0; selp.f32 %0, %1, %2, q;}" : "=f"(r) : "f"(a), "f"(b), "r"(p));
  return r;
}
__device__ __forceinline__ int lea(int a, int b) {
  int r;
  asm volatile("{.reg .b32 t; shl.b32 t, %1, 23; add.s32 %0, t, %2;}" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ float fmnmx(float a, float b) {
  float r;
  asm volatile("max.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

// NF2: FFMA2 per iteration; NF1: scalar FFMA; NX: MUFU; NC: f32->f64->f32 round trips; ND: DADD; NS: FSEL; NL: LEA; NM: FMNMX
template <int NF2, int NF1, int NX, int NC, int ND, int NS, int NL, int NM>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
  u64 p[8];
  float a[8], x[4], c[4], s[4], m[4];
  double d[4];
  int l[4];
  for (int i = 0; i < 8; ++i) {
    a[i] = seed + i + threadIdx.x;
    float2 v = make_float2(a[i], a[i] * 0.5f);
    p[i] = *reinterpret_cast<u64*>(&v);
  }
  for (int i = 0; i < 4; ++i) {
    x[i] = 0.001f * (threadIdx.x + i);
    c[i] = 1.0f + i + threadIdx.x;
    d[i] = 1.0 + i;
    s[i] = 2.0f + i;
    l[i] = threadIdx.x + i;
    m[i] = 0.5f * i;
  }
  float2 cc = make_float2(0.999f, 1.001f);
  const u64 c2 = *reinterpret_cast<u64*>(&cc);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < NF2) p[i] = fma2(p[i], c2, c2);
      if (i < NF1) a[i] = fma1(a[i], 0.999f, 0.5f);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i < NX) x[i] = ex2(x[i]);
      if (i < NC) c[i] = d2f(f2d(c[i]));
      if (i < ND) d[i] = my_dadd(d[i], 1.0);
      if (i < NS) s[i] = fsel(s[i], x[0], it & 1);
      if (i < NL) l[i] = lea(l[i], it);
      if (i < NM) m[i] = fmnmx(m[i], s[i]);
    }
  }
  float t = 0;
  for (int i = 0; i < 8; ++i) {
    float2 v = *reinterpret_cast<float2*>(&p[i]);
    t += a[i] + v.x + v.y;
  }
  for (int i = 0; i < 4; ++i) t += x[i] + c[i] + (float)d[i] + s[i] + l[i] + m[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int NF2, int NF1, int NX, int NC, int ND, int NS, int NL, int NM>
void run(const char* name, float* out, double base) {
  const int iters = 20000, blocks = 148 * 8;  // 8 blocks x 8 warps per SM = 16 warps per scheduler
  k<NF2, NF1, NX, NC, ND, NS, NL, NM><<<blocks, 256>>>(out, 100, 1.0f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<NF2, NF1, NX, NC, ND, NS, NL, NM><<<blocks, 256>>>(out, iters, 1.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double cyc = ms * 1e-3 * 1.965e9 / (iters * 16.0);
  printf("%-44s %8.3f ms  %6.2f cycles / iteration / scheduler", name, ms, cyc);
  if (base > 0) printf("   (+%.2f over FFMA2 x8)", cyc - base);
  printf("\n");
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  run<0, 8, 0, 0, 0, 0, 0, 0>("FFMA x8", out, 0);
  run<8, 0, 0, 0, 0, 0, 0, 0>("FFMA2 x8", out, 0);
  const double b = 16.4;  // printed reference; see the first lines for the measured value
  run<8, 0, 2, 0, 0, 0, 0, 0>("FFMA2 x8 + MUFU x2", out, b);
  run<8, 0, 4, 0, 0, 0, 0, 0>("FFMA2 x8 + MUFU x4", out, b);
  run<8, 0, 0, 2, 0, 0, 0, 0>("FFMA2 x8 + (F2F.64.32 + F2F.32.64) x2", out, b);
  run<8, 0, 0, 0, 2, 0, 0, 0>("FFMA2 x8 + DADD x2", out, b);
  run<8, 0, 0, 0, 4, 0, 0, 0>("FFMA2 x8 + DADD x4", out, b);
  run<8, 0, 0, 0, 0, 4, 0, 0>("FFMA2 x8 + FSEL x4", out, b);
  run<8, 0, 0, 0, 0, 0, 4, 0>("FFMA2 x8 + SHL/IADD(LEA) x4", out, b);
  run<8, 0, 0, 0, 0, 0, 0, 4>("FFMA2 x8 + FMNMX x4", out, b);
  run<8, 0, 2, 2, 2, 2, 2, 2>("FFMA2 x8 + 2 of each", out, b);
  run<0, 0, 4, 0, 0, 0, 0, 0>("MUFU x4", out, 0);
  run<0, 0, 0, 4, 0, 0, 0, 0>("(F2F + F2F) x4", out, 0);
  run<0, 0, 0, 0, 4, 0, 0, 0>("DADD x4", out, 0);
  run<0, 0, 0, 0, 0, 4, 0, 0>("FSEL x4", out, 0);
  run<0, 8, 2, 0, 0, 0, 0, 0>("FFMA x8 + MUFU x2", out, 0);
  run<0, 8, 0, 0, 0, 4, 0, 0>("FFMA x8 + FSEL x4", out, 0);
  return 0;
}
