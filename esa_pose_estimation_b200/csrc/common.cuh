// Shared helpers for the sm_100a kernels of esa_pose_estimation_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/esa_pose_b200.h"

namespace epb {

extern thread_local int g_last_cuda_error;
extern unsigned long long g_launch_count;

inline int check_launch() {
  cudaError_t e = cudaGetLastError();
  __atomic_fetch_add(&g_launch_count, 1ull, __ATOMIC_RELAXED);
  if (e != cudaSuccess) {
    g_last_cuda_error = (int)e;
    return EPB_ERR_CUDA;
  }
  return EPB_OK;
}

inline int check_api(cudaError_t e) {
  if (e != cudaSuccess) {
    g_last_cuda_error = (int)e;
    return EPB_ERR_CUDA;
  }
  return EPB_OK;
}

// Optional per-kernel-class timing with CUDA events on the launching stream (bench.py roofline):
// enabled by epb_profile_enable(); otherwise a no-op.
enum { PROF_COMPACT = 0, PROF_HYPOTHESIS = 1, PROF_VOTE_COUNT = 2, PROF_REFINE = 3, PROF_POSE = 4,
       PROF_DECODE = 5, PROF_CLASSES = 6 };
void prof_begin(int cls, cudaStream_t s);
void prof_end(int cls, cudaStream_t s);
struct ProfScope {
  int cls; cudaStream_t s;
  ProfScope(int c, cudaStream_t st) : cls(c), s(st) { prof_begin(cls, s); }
  ~ProfScope() { prof_end(cls, s); }
};

#define EPB_RETURN_IF(expr)            \
  do {                                 \
    int _s = (expr);                   \
    if (_s != EPB_OK) return _s;       \
  } while (0)

// Shared-memory carve-out.  An SM's L1 / shared-memory split is fixed while any CTA is resident, and it is
// chosen by whichever kernel lands on the empty SM first.  Our kernels run concurrently on several streams
// (gather of the next batch chunk under the voting of the current one, pose solve under the next call's
// voting); a kernel without shared memory that arrives first leaves the SM with a split in which only one
// or two of vote_count's 16 KB CTAs fit, which slowed a voting chunk from 0.56 to 2.6 ms in the co-run
// microbenchmark (tools/corun.py).  Every kernel of this library therefore asks for a large shared-memory
// carve-out (60 % measured best end to end: 2.36 ms per batch vs 2.44 at 72 %, 2.57 at 100 %, 2.69 without);
// none of them depends on L1 capacity (streaming loads, shared-memory tiles) except the compaction
// scatter, which is handled at its launch site.
// Tuning hooks.  The shipped library reads NO environment variables: overrides exist only in builds with
// -DEPB_TUNING (EPB_EXTRA_NVCC_FLAGS=-DEPB_TUNING python -m esa_pose_estimation_b200.build --force), which the
// profiling scripts under tools/ use.  Knobs: EPB_CARVEOUT, EPB_VOTE_IMPL, EPB_VOTE_ITEM, EPB_VOTE_R,
// EPB_GATHER_SPLIT, EPB_GATHER_CTAS[_PER_SM], EPB_GATHER_LIGHT, EPB_GATHER_VSPLIT.
inline int tuning_int(const char* name, int dflt) {
#ifdef EPB_TUNING
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
#else
  (void)name;
  return dflt;
#endif
}
inline int carveout_percent() {
  static const int pct = tuning_int("EPB_CARVEOUT", 60);
  return pct;
}
// SM count of the current device, queried once per device
inline int device_sm_count() {
  static int cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = __atomic_load_n(&cached[dev], __ATOMIC_RELAXED);
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    __atomic_store_n(&cached[dev], n, __ATOMIC_RELAXED);
  }
  return n;
}
// Is `ptr` page-locked host memory?  (cudaPointerGetAttributes costs ~1 us; the answer only steers which of two
// equivalent gather kernels runs, so a one-entry per-thread cache keyed by the pointer is safe.)
inline bool is_host_pointer(const void* ptr) {
  static thread_local const void* last_ptr = nullptr;
  static thread_local bool last_host = false;
  if (ptr != last_ptr || ptr == nullptr) {
    cudaPointerAttributes attr;
    last_host = ptr && cudaPointerGetAttributes(&attr, ptr) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    last_ptr = ptr;
  }
  return last_host;
}
template <typename K>
inline void prefer_max_shared(K kernel) {
  cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), cudaFuncAttributePreferredSharedMemoryCarveout,
                       carveout_percent());
}
// runs `init` once per device and file
#define EPB_INIT_ONCE_PER_DEVICE(init)                                   \
  do {                                                                   \
    static unsigned long long done_mask = 0;                             \
    int dev_ = 0;                                                        \
    if (cudaGetDevice(&dev_) == cudaSuccess && dev_ >= 0 && dev_ < 64 &&  \
        !((__atomic_load_n(&done_mask, __ATOMIC_ACQUIRE) >> dev_) & 1ull)) { \
      init();                                                            \
      __atomic_fetch_or(&done_mask, 1ull << dev_, __ATOMIC_RELEASE);     \
    }                                                                    \
  } while (0)

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double shfl_xor_d(double v, int m) {
  return __shfl_xor_sync(FULL, v, m);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
  return v;
}
__device__ __forceinline__ double bcast_d(double v, int lane) { return __shfl_sync(FULL, v, lane); }

// streaming 128-bit load that does not allocate in L1 (data is touched once)
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// streaming 32-bit load, no L1 allocation (works for device memory and for mapped host memory)
__device__ __forceinline__ float ldg_stream(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// ---- mbarrier + TMA bulk copy (cp.async.bulk global -> shared; SASS: UBLKCP + SYNCS) -------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP + SYNCS)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

}  // namespace epb
