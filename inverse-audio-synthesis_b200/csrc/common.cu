// Library-level entry points: version, last error, device check.
#include "ias_common.cuh"

namespace ias {

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

namespace {
constexpr int PROF_RING = 8192;
const char* const kNames[K_COUNT] = {"k_seed_params", "k_voice_control", "k_voice_audio", "k_pqmf_analysis",
                                     "k_pqmf_synthesis", "k_vicreg_colsum", "k_vicreg_center_pack", "k_vicreg_gram_tc",
                                     "k_vicreg_cov_reduce", "k_vicreg_finalize", "k_vicreg_gram_simt", "k_vicreg_backward",
                                     "k_abs_avg_pool", "k_voice_schedule", "k_voice_adsr", "k_pool_finalize",
                                     "k_vicreg_stats_publish", "k_vicreg_stats_combine"};
struct Prof {
  bool on = false;
  long long launches[K_COUNT] = {0};
  cudaEvent_t beg[PROF_RING], end[PROF_RING];
  int ids[PROF_RING];
  int created = 0, used = 0;
};
Prof& prof() {
  static Prof p;
  return p;
}
}  // namespace

ProfScope::ProfScope(int id, cudaStream_t st) : id_(id), slot_(-1), st_(st) {
  Prof& p = prof();
  p.launches[id]++;
  if (!p.on || p.used >= PROF_RING) return;
  if (p.used >= p.created) {
    if (cudaEventCreate(&p.beg[p.created]) != cudaSuccess || cudaEventCreate(&p.end[p.created]) != cudaSuccess) return;
    p.created++;
  }
  slot_ = p.used++;
  p.ids[slot_] = id;
  cudaEventRecord(p.beg[slot_], st);
}
ProfScope::~ProfScope() {
  if (slot_ >= 0) cudaEventRecord(prof().end[slot_], st_);
}

}  // namespace ias

extern "C" int ias_prof_enable(int on) {
  ias::prof().on = on != 0;
  return IAS_OK;
}
extern "C" int ias_prof_reset(void) {
  ias::Prof& p = ias::prof();
  p.used = 0;
  for (int i = 0; i < ias::K_COUNT; ++i) p.launches[i] = 0;
  return IAS_OK;
}
extern "C" int ias_prof_kernel_count(void) { return ias::K_COUNT; }
extern "C" const char* ias_prof_kernel_name(int id) { return (id >= 0 && id < ias::K_COUNT) ? ias::kNames[id] : nullptr; }
extern "C" long long ias_prof_launches(int id) {
  ias::Prof& p = ias::prof();
  if (id < 0) {
    long long t = 0;
    for (int i = 0; i < ias::K_COUNT; ++i) t += p.launches[i];
    return t;
  }
  return id < ias::K_COUNT ? p.launches[id] : 0;
}
extern "C" int ias_prof_read(int id, double* total_ms, long long* timed_launches) {
  ias::Prof& p = ias::prof();
  IAS_REQUIRE(id >= 0 && id < ias::K_COUNT && total_ms && timed_launches, IAS_ERR_INVALID, "ias_prof_read: bad arguments");
  double t = 0.0;
  long long n = 0;
  for (int i = 0; i < p.used; ++i) {
    if (p.ids[i] != id) continue;
    IAS_CUDA(cudaEventSynchronize(p.end[i]));
    float ms = 0.f;
    IAS_CUDA(cudaEventElapsedTime(&ms, p.beg[i], p.end[i]));
    t += ms;
    ++n;
  }
  *total_ms = t;
  *timed_launches = n;
  return IAS_OK;
}

extern "C" int ias_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char* ias_last_error(void) { return ias::err_buf(); }

extern "C" int ias_device_check(int device) {
  cudaDeviceProp p;
  IAS_CUDA(cudaGetDeviceProperties(&p, device));
  IAS_REQUIRE(p.major == 10, IAS_ERR_UNSUPPORTED, "device %d is sm_%d%d; libias_b200 is built for sm_100a only", device,
              p.major, p.minor);
  return IAS_OK;
}
