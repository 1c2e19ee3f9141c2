// api.cu — library bookkeeping: ABI version, last-error string, launch counter.
#include <atomic>
#include <string.h>

#include "common.cuh"

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
