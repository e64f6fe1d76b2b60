// Harness utility (NOT a reference surface): adaptive average pooling of |bands| to P bins per sound.
//
// The reference maps PQMF bands to embeddings with a MobileNet backbone that is out of scope; the
// benchmark and tests stand in a fixed "bridge": adaptive_avg_pool1d(|bands|.reshape(B,1,N*L), 256) followed by two
// small projections.  torch's adaptive_avg_pool kernel runs that pooling at ~380 GB/s (1.9 ms for 1024 sounds, as
// long as the whole synth); this kernel does it at memory speed so the stand-in does not distort the measurement.
// One warp per (sound, bin): coalesced strided reads, lane partial sums, shuffle tree -- deterministic.
#include "ias_common.cuh"

namespace ias {
namespace {

__global__ void __launch_bounds__(256) k_abs_avg_pool(const float* __restrict__ x, float* __restrict__ out, int B,
                                                      long long S, int P) {
  const long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= (long long)B * P) return;
  const int lane = threadIdx.x & 31;
  // sounds are taken last-first: the producer (PQMF analysis) wrote them in ascending order, so the tail of its
  // output is still in L2 when this kernel starts
  const int b = B - 1 - (int)(w / P), i = (int)(w % P);
  // torch adaptive pooling bin: [floor(i*S/P), ceil((i+1)*S/P))
  const long long s0 = (i * S) / P;
  const long long s1 = ((i + 1) * S + P - 1) / P;
  const float* row = x + (size_t)b * S;
  float acc = 0.0f;
  if (((reinterpret_cast<uintptr_t>(row) & 15u) == 0) && s1 - s0 <= 4 * 32 * 8) {
    // unaligned head and tail as scalar loads, the body as 128-bit loads, all issued before the first add
    const long long a0 = (s0 + 3) & ~3LL, a1 = s1 & ~3LL;
    const int n4 = a1 > a0 ? (int)((a1 - a0) >> 2) : 0;
    const float4* r4 = reinterpret_cast<const float4*>(row + a0);
    float4 v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = (lane + 32 * q < n4) ? __ldg(r4 + lane + 32 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
    float e = 0.0f;
    if (a1 > a0) {
      if (s0 + lane < a0) e = fabsf(__ldg(row + s0 + lane));
      if (a1 + lane < s1) e += fabsf(__ldg(row + a1 + lane));
    } else {
      for (long long j = s0 + lane; j < s1; j += 32) e += fabsf(__ldg(row + j));
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) acc += (fabsf(v[q].x) + fabsf(v[q].y)) + (fabsf(v[q].z) + fabsf(v[q].w));
    acc += e;
  } else {
    for (long long j = s0 + lane; j < s1; j += 32) acc += fabsf(__ldg(row + j));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[(size_t)b * P + i] = acc / (float)(s1 - s0);
}

}  // namespace
}  // namespace ias

using namespace ias;

extern "C" int ias_abs_avg_pool(const float* x, float* out, int B, long long S, int P, ias_stream_t stream) {
  IAS_REQUIRE(B > 0 && S > 0 && P > 0 && P <= S, IAS_ERR_INVALID, "ias_abs_avg_pool: B=%d S=%lld P=%d", B, S, P);
  IAS_REQUIRE(x && out, IAS_ERR_INVALID, "ias_abs_avg_pool: NULL pointer");
  const long long warps = (long long)B * P;
  {
    ProfScope prof_(K_ABS_AVG_POOL, as_stream(stream));
    k_abs_avg_pool<<<(unsigned)((warps + 7) / 8), 256, 0, as_stream(stream)>>>(x, out, B, S, P);
  }
  IAS_LAUNCH_CHECK("k_abs_avg_pool");
  return IAS_OK;
}
