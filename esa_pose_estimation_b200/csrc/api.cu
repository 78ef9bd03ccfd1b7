// Library-level entry points of the C ABI (include/esa_pose_b200.h).
#include "common.cuh"

#include <atomic>
#include <mutex>
#include <vector>

namespace epb {
thread_local int g_last_cuda_error = 0;
unsigned long long g_launch_count = 0;

static std::atomic<bool> g_prof_on{false};   // read on every launch without the mutex
static std::mutex g_prof_mu;
struct ProfRec { int cls; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof;
static thread_local cudaEvent_t g_prof_open[PROF_CLASSES];

void prof_begin(int cls, cudaStream_t s) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  cudaEvent_t a;
  if (cudaEventCreate(&a) != cudaSuccess) return;
  cudaEventRecord(a, s);
  g_prof_open[cls] = a;
}
void prof_end(int cls, cudaStream_t s) {
  if (!g_prof_on.load(std::memory_order_relaxed) || !g_prof_open[cls]) return;
  cudaEvent_t b;
  if (cudaEventCreate(&b) != cudaSuccess) return;
  cudaEventRecord(b, s);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back({cls, g_prof_open[cls], b});
  g_prof_open[cls] = nullptr;
}
}  // namespace epb

extern "C" int epb_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(epb::g_prof_mu);
  epb::g_prof_on.store(on != 0, std::memory_order_relaxed);
  if (!on) {
    for (auto& r : epb::g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    epb::g_prof.clear();
  }
  return EPB_OK;
}

extern "C" int epb_profile_read(int cls, double* total_ms, int* launches) {
  std::lock_guard<std::mutex> lk(epb::g_prof_mu);
  double t = 0; int n = 0;
  for (auto& r : epb::g_prof) {
    if (r.cls != cls) continue;
    if (cudaEventSynchronize(r.b) != cudaSuccess) return EPB_ERR_CUDA;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) return EPB_ERR_CUDA;
    t += ms; ++n;
  }
  if (total_ms) *total_ms = t;
  if (launches) *launches = n;
  return EPB_OK;
}

extern "C" int epb_version(void) { return EPB_VERSION; }

extern "C" int epb_last_cuda_error(void) { return epb::g_last_cuda_error; }

extern "C" const char* epb_last_cuda_error_string(void) {
  return cudaGetErrorString((cudaError_t)epb::g_last_cuda_error);
}

extern "C" unsigned long long epb_launch_count(void) {
  return __atomic_load_n(&epb::g_launch_count, __ATOMIC_RELAXED);
}

extern "C" int epb_device_info(int* sm_count, int* sm_clock_khz, size_t* l2_bytes) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return EPB_ERR_NO_DEVICE;
  int v = 0;
  if (sm_count) { cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); *sm_count = v; }
  if (sm_clock_khz) { cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, dev); *sm_clock_khz = v; }
  if (l2_bytes) { cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev); *l2_bytes = (size_t)v; }
  return EPB_OK;
}
