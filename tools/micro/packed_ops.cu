// Microbenchmark (B200): throughput of the three packed fp32 instructions (FFMA2 / FADD2 / FMUL2) and of the audio
// kernel's pitch chain (vco_increment_p2, no-clamp form) as independent chains, at 4 and 16 warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o packed_ops packed_ops.cu && ./packed_ops
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 r;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 r;
  asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 pk(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk(u64 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); }
__device__ __forceinline__ u64 bc(float a) { return pk(a, a); }

// MODE 0: FFMA2 x8, 1: FADD2 x8, 2: FMUL2 x8, 3: FFMA2 x4 + FADD2 x2 + FMUL2 x2, 4: pitch chain x NCH independent pairs
template <int MODE, int NCH>
__global__ void __launch_bounds__(128) k(float* out, int iters, float seed) {
  u64 p[8];
  for (int i = 0; i < 8; ++i) p[i] = pk(seed + i + threadIdx.x, seed * 0.5f + i);
  const u64 c2 = pk(0.999f, 1.001f), c3 = pk(1e-3f, -1e-3f);
  for (int it = 0; it < iters; ++it) {
    if (MODE <= 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) p[i] = fma2(p[i], c2, c3);
        if (MODE == 1) p[i] = add2(p[i], c3);
        if (MODE == 2) p[i] = mul2(p[i], c2);
        if (MODE == 3) p[i] = (i < 4) ? fma2(p[i], c2, c3) : (i < 6) ? add2(p[i], c3) : mul2(p[i], c2);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        // vco_increment_p2<false>: 20 packed instructions + 2 x (shift, add) per pair
        const u64 mod = p[i];
        u64 m = add2(bc(60.0f), fma2(bc(0.37f), mod, bc(0.0f)));
        const u64 a = add2(m, bc(-69.0f));
        const u64 q = mul2(a, bc(0.0833333358168601989746f));
        const u64 d = fma2(fma2(q, bc(-12.0f), a), bc(0.0833333358168601989746f), q);
        const float magic = 12582912.0f;
        const u64 t = add2(d, bc(magic));
        const u64 sf = add2(d, mul2(add2(t, bc(-magic)), bc(-1.0f)));
        u64 u = bc(+0.1535920892e-3f);
        u = fma2(u, sf, bc(+0.1339262701e-2f));
        u = fma2(u, sf, bc(+0.9618384764e-2f));
        u = fma2(u, sf, bc(+0.5550347269e-1f));
        u = fma2(u, sf, bc(+0.2402264476e+0f));
        u = fma2(u, sf, bc(+0.6931471825e+0f));
        u = fma2(u, sf, bc(1.0f));
        float u0, u1, t0, t1;
        unpk(u, u0, u1);
        unpk(t, t0, t1);
        const float e0 = __int_as_float(__float_as_int(u0) + (__float_as_int(t0) << 23));
        const float e1 = __int_as_float(__float_as_int(u1) + (__float_as_int(t1) << 23));
        const u64 w = mul2(bc(6.2831855f), mul2(bc(440.0f), pk(e0, e1)));
        const u64 qq = mul2(w, bc(2.2675737e-5f));
        const u64 inc = fma2(fma2(qq, bc(-44100.0f), w), bc(2.2675737e-5f), qq);
        p[i] = add2(mul2(inc, bc(0.01f)), bc(0.3f));  // keep the chain alive and bounded (2 more packed instructions)
      }
    }
  }
  float t = 0;
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    unpk(p[i], lo, hi);
    t += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE, int NCH>
void run(const char* name, float* out, int ctas_per_sm, double instr_per_iter) {
  const int iters = 4000, blocks = 148 * ctas_per_sm;  // 128-thread CTAs: ctas_per_sm warps per scheduler
  k<MODE, NCH><<<blocks, 128>>>(out, 50, 1.0f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE, NCH><<<blocks, 128>>>(out, iters, 1.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double cyc = ms * 1e-3 * 1.965e9 / ((double)iters * ctas_per_sm);
  printf("%-40s %2d warps/sched %9.2f cycles / iteration / warp   %5.2f cycles / packed instruction\n", name, ctas_per_sm,
         cyc, cyc / instr_per_iter);
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 16 * 128 * sizeof(float));
  for (int w : {4, 16}) {
    run<0, 8>("FFMA2 x8", out, w, 8);
    run<1, 8>("FADD2 x8", out, w, 8);
    run<2, 8>("FMUL2 x8", out, w, 8);
    run<3, 8>("FFMA2 x4 + FADD2 x2 + FMUL2 x2", out, w, 8);
    run<4, 1>("pitch chain x1 (22 packed)", out, w, 22);
    run<4, 2>("pitch chain x2", out, w, 44);
    run<4, 4>("pitch chain x4", out, w, 88);
    run<4, 8>("pitch chain x8", out, w, 176);
  }
  return 0;
}
