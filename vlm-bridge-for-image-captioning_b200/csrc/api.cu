// Process-wide state of libb200_bridge.so: last-error string, launch counter, device checks.
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "../../include/b200_bridge.h"
#include "launch.h"

namespace b200b {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return B200B_OK;
}

int device_sm_count(int* out) {
  static int cached_dev = -1, cached_sms = 0;  // one process per GPU; a race only repeats the query
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_last_error("cudaGetDevice failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  if (dev != cached_dev) {
    int major = 0, sms = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (major != 10) {
      set_last_error("device %d has compute capability %d.x; this library is sm_100a only", dev, major);
      return B200B_ERR_DEVICE;
    }
    cached_sms = sms;
    cached_dev = dev;
  }
  *out = cached_sms;
  return B200B_OK;
}

}  // namespace b200b

extern "C" int b200b_abi_version(void) { return B200B_ABI_VERSION; }
extern "C" const char* b200b_last_error(void) { return b200b::g_err; }
extern "C" uint64_t b200b_launch_count(void) { return b200b::g_launches.load(std::memory_order_relaxed); }
