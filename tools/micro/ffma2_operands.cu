// Microbenchmark (B200): does the cost of a packed FFMA2 depend on where its operands come from?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_operands ffma2_operands.cu && ./ffma2_operands
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ u64 pk(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
// MODE 0: acc = fma2(acc, c, c)            one loop-invariant pair used twice
// MODE 1: acc_i = fma2(acc_i, s_i, t_i)    three distinct register pairs per instruction (s_i, t_i loop invariant)
// MODE 2: acc_i = fma2(acc_i, bcast(s_i), bcast(t_i))   scalar registers broadcast to both lanes
// MODE 3: acc_i = fma2(acc_i, acc_j, t_i)  two accumulators + one invariant
// MODE 4: scalar: acc_i = fma(acc_i, s_i, t_i) x16 (same lane-FMA count as 8 FFMA2)
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
  u64 acc[8], s[8], t[8];
  float fa[16], fs[8], ft[8];
  for (int i = 0; i < 8; ++i) {
    fs[i] = 0.999f + 1e-4f * (i + threadIdx.x % 3);
    ft[i] = 0.5f + 1e-3f * i;
    acc[i] = pk(seed + i, seed - i);
    s[i] = pk(fs[i], fs[i] + 1e-5f);
    t[i] = pk(ft[i], ft[i] + 1e-5f);
    fa[2 * i] = seed + i;
    fa[2 * i + 1] = seed - i;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) acc[i] = fma2(acc[i], s[0], s[0]);
      if (MODE == 1) acc[i] = fma2(acc[i], s[i], t[i]);
      if (MODE == 2) acc[i] = fma2(acc[i], pk(fs[i], fs[i]), pk(ft[i], ft[i]));
      if (MODE == 3) acc[i] = fma2(acc[i], acc[(i + 3) & 7], t[i]);
      if (MODE == 4) {
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(fa[2 * i]) : "f"(fs[i]), "f"(ft[i]));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(fa[2 * i + 1]) : "f"(fs[i]), "f"(ft[i]));
      }
    }
  }
  float r = 0;
  for (int i = 0; i < 8; ++i) {
    float2 v = *reinterpret_cast<float2*>(&acc[i]);
    r += v.x + v.y + fa[2 * i] + fa[2 * i + 1];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char* name, float* out, int warps_per_sched) {
  const int iters = 20000, blocks = 148 * warps_per_sched / 2;  // 256 threads = 8 warps = 2 per scheduler
  k<MODE><<<blocks, 256>>>(out, 100, 1.0f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, iters, 1.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("%-58s %2d warps/sched  %7.2f cycles / iteration / scheduler\n", name, warps_per_sched,
         ms * 1e-3 * 1.965e9 / (iters * (double)warps_per_sched));
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  for (int w : {16, 4, 2}) {
    if (w == 16) {
      run<0>("FFMA2 x8: acc, c, c", out, 16);
      run<1>("FFMA2 x8: three distinct register pairs", out, 16);
      run<2>("FFMA2 x8: acc, broadcast scalar, broadcast scalar", out, 16);
      run<3>("FFMA2 x8: acc, other acc, invariant pair", out, 16);
      run<4>("FFMA x16 scalar", out, 16);
    } else if (w == 4) {
      run<1>("FFMA2 x8: three distinct register pairs", out, 4);
      run<4>("FFMA x16 scalar", out, 4);
    } else {
      run<1>("FFMA2 x8: three distinct register pairs", out, 2);
      run<4>("FFMA x16 scalar", out, 2);
    }
  }
  return 0;
}
