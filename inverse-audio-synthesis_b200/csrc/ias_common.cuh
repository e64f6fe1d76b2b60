// Shared host-side plumbing for libias_b200.so: error state, argument checks, launch checks.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/ias_b200.h"

namespace ias {

// Thread-local last-error buffer (ias_last_error()).
char* err_buf();
int set_err(int code, const char* fmt, ...);

inline cudaStream_t as_stream(ias_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Kernel ids of the launch counter / event profiler (ias_prof_*).
enum KernelId {
  K_SEED_PARAMS = 0, K_VOICE_CONTROL, K_VOICE_AUDIO, K_PQMF_ANALYSIS, K_PQMF_SYNTHESIS, K_VICREG_COLSUM,
  K_VICREG_PACK, K_VICREG_GRAM_TC, K_VICREG_COV_REDUCE, K_VICREG_FINALIZE, K_VICREG_GRAM_SIMT, K_VICREG_BWD, K_ABS_AVG_POOL, K_VOICE_SCHEDULE, K_VOICE_ADSR, K_POOL_FINALIZE,
  K_VICREG_STATS_PUBLISH, K_VICREG_STATS_COMBINE, K_COUNT
};

// RAII bracket around one kernel launch: always counts it; when profiling is on also records a CUDA event pair on
// the launch stream (elapsed times are read back by ias_prof_read after a synchronise).
struct ProfScope {
  ProfScope(int id, cudaStream_t st);
  ~ProfScope();
  int id_;
  int slot_;
  cudaStream_t st_;
};

}  // namespace ias

#define IAS_REQUIRE(cond, code, ...)                     \
  do {                                                   \
    if (!(cond)) return ias::set_err((code), __VA_ARGS__); \
  } while (0)

#define IAS_CUDA(call)                                                                               \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess)                                                                          \
      return ias::set_err(IAS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                          __LINE__);                                                                 \
  } while (0)

#define IAS_LAUNCH_CHECK(name)                                                                      \
  do {                                                                                              \
    cudaError_t e__ = cudaGetLastError();                                                           \
    if (e__ != cudaSuccess)                                                                         \
      return ias::set_err(IAS_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__));   \
  } while (0)

static inline bool ias_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Function attributes (opt-in shared memory) are per device: `seen` is a caller-owned bitmask of the devices already
// configured.  Returns true the first time it is called for the current device.
static inline bool ias_first_use_on_device(unsigned long long& seen) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return true;
  const unsigned long long bit = 1ull << dev;
  if (seen & bit) return false;
  seen |= bit;
  return true;
}
