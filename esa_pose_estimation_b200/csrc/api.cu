// Library-level entry points of the C ABI (include/esa_pose_b200.h).
#include "common.cuh"

namespace epb {
thread_local int g_last_cuda_error = 0;
unsigned long long g_launch_count = 0;
}  // namespace epb

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
