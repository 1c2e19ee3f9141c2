// api.cu — library bookkeeping: ABI version, last-error string, launch counter.
#include <atomic>
#include <string.h>

#include <cstring>

#include "common.cuh"
#include "internal.cuh"

namespace tpg {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return kNumSMsB200;
  }
  return cached;
}

}  // namespace tpg

TPG_API int tpg_abi_version(void) { return TPG_ABI_VERSION; }
TPG_API const char* tpg_last_error(void) { return tpg::g_err; }
TPG_API uint64_t tpg_launch_count(void) { return tpg::g_launches.load(std::memory_order_relaxed); }
TPG_API int tpg_set_option(const char* name, long value) {
  TPG_REQUIRE(name != nullptr, TPG_EINVAL, "set_option: null name");
  if (strcmp(name, "fps.sms_per_cloud") == 0) {
    TPG_REQUIRE(value == 1 || value == 2 || value == 4 || value == 8, TPG_EINVAL, "set_option: fps.sms_per_cloud must be 1, 2, 4 or 8");
    tpg::fps_cluster_option().store((int)value, std::memory_order_relaxed);
    return TPG_OK;
  }
  if (strcmp(name, "fps.exclusive_sm") == 0) {
    TPG_REQUIRE(value == 0 || value == 1, TPG_EINVAL, "set_option: fps.exclusive_sm must be 0 or 1");
    tpg::fps_exclusive_option().store((int)value, std::memory_order_relaxed);
    return TPG_OK;
  }
  tpg::set_error("set_option: unknown option '%s'", name);
  return TPG_EINVAL;
}

// tuning hook (tools/timeline.py; not part of the ABI): writes the GPU global timer (ns) to *slot on `stream`
namespace {
__global__ void timestamp_kernel(long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *slot = (long long)t;
}
}  // namespace
TPG_API int tpg_debug_timestamp(long long* slot, tpg_stream_t stream) {
  timestamp_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(slot);
  return TPG_OK;
}
