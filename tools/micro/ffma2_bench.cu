// Microbenchmark (B200): issue rate of packed fp32 (FFMA2/FADD2/FMUL2) against scalar FFMA, alone and mixed with
// ALU-pipe instructions.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
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
__device__ __forceinline__ int alu1(int a, int b) {
  int r;
  asm volatile("lop3.b32 %0, %1, %2, %1, 0x96;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
  float a[8];
  u64 p[8];
  int z[4];
  for (int i = 0; i < 8; ++i) {
    a[i] = seed + i + threadIdx.x;
    float2 v = make_float2(a[i], a[i] * 0.5f);
    p[i] = *reinterpret_cast<u64*>(&v);
  }
  for (int i = 0; i < 4; ++i) z[i] = threadIdx.x + i;
  float2 cc = make_float2(0.999f, 1.001f);
  const u64 c2 = *reinterpret_cast<u64*>(&cc);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0 || MODE == 2) a[i] = fma1(a[i], 0.999f, 0.5f);
      if (MODE == 1 || MODE == 3) p[i] = fma2(p[i], c2, c2);
    }
    if (MODE == 2 || MODE == 3) {
#pragma unroll
      for (int i = 0; i < 4; ++i) z[i] = alu1(z[i], it);
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) {
    float2 v = *reinterpret_cast<float2*>(&p[i]);
    s += a[i] + v.x + v.y;
  }
  for (int i = 0; i < 4; ++i) s += z[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double lane_fma_per_iter, float* out) {
  const int iters = 20000, blocks = 148 * 8;
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
  double fmas = (double)blocks * 256 * iters * lane_fma_per_iter;
  printf("%-28s %8.3f ms  %7.2f T lane-FMA/s  (%.1f lane-FMA/clk/SM at 1.965 GHz)\n", name, ms, fmas / ms / 1e9,
         fmas / (ms * 1e-3) / 148 / 1.965e9);
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  run<0>("FFMA x8", 8, out);
  run<1>("FFMA2 x8", 16, out);
  run<2>("FFMA x8 + LOP3 x4", 8, out);
  run<3>("FFMA2 x8 + LOP3 x4", 16, out);
  return 0;
}
