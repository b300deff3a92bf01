// Process-wide state of libb200_bridge.so: last-error string, launch counter, device checks.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200_bridge.h"
#include "launch.h"

namespace b200b {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_sm_limit{0};  // b200b_set_sm_limit: 0 = all SMs

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- optional per-kernel timing (bench.py's roofline leg; off by default) ----------------------
struct ProfEvent {
  cudaEvent_t ev;
  const char* what;
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static cudaStream_t g_prof_stream = nullptr;  // only launches on the profiled stream are timed
static std::vector<ProfEvent> g_prof_events;

int check_launch(const char* what, cudaStream_t stream) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (g_prof_on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (g_prof_on && stream == g_prof_stream) {
      ProfEvent pe;
      pe.what = what;
      if (cudaEventCreate(&pe.ev) == cudaSuccess) {
        cudaEventRecord(pe.ev, stream);
        g_prof_events.push_back(pe);
      }
    }
  }
  return B200B_OK;
}

int pdl_mask() {
  static const int mask = [] { const char* e = getenv("B200B_PDL"); return e ? atoi(e) : 7; }();
  return mask;
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
  const int limit = g_sm_limit.load(std::memory_order_relaxed);
  *out = (limit > 0 && limit < cached_sms) ? limit : cached_sms;
  return B200B_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn tmap_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

struct HeadsMapKey {
  const void* base;
  long long ld;
  int head_dim, heads, rows, batch, box_rows;
  bool operator==(const HeadsMapKey& o) const {
    return base == o.base && ld == o.ld && head_dim == o.head_dim && heads == o.heads && rows == o.rows &&
           batch == o.batch && box_rows == o.box_rows;
  }
};
struct HeadsMapEntry {
  HeadsMapKey key;
  CUtensorMap map;
  bool used;
};
static std::mutex g_tmap_mu;
static HeadsMapEntry g_tmap_cache[128];
static unsigned g_tmap_next = 0;

int make_tmap_heads_sw64(CUtensorMap* tm, const void* base, long long ld_elems, int head_dim, int heads, int rows,
                         int batch, int box_rows) {
  const HeadsMapKey key{base, ld_elems, head_dim, heads, rows, batch, box_rows};
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    for (auto& e : g_tmap_cache)
      if (e.used && e.key == key) {
        *tm = e.map;
        return B200B_OK;
      }
  }
  EncodeTiledFn fn = tmap_encode_fn();
  if (fn == nullptr) {
    set_last_error("cuTensorMapEncodeTiled not available from the driver");
    return B200B_ERR_DRIVER;
  }
  cuuint64_t dims[4] = {(cuuint64_t)head_dim, (cuuint64_t)heads, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[3] = {(cuuint64_t)head_dim * 2, (cuuint64_t)ld_elems * 2, (cuuint64_t)rows * ld_elems * 2};
  cuuint32_t box[4] = {32, 1, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled (head slices) failed (%d): base=%p ld=%lld head_dim=%d heads=%d rows=%d "
                   "batch=%d box_rows=%d", (int)r, base, ld_elems, head_dim, heads, rows, batch, box_rows);
    return B200B_ERR_TENSORMAP;
  }
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  HeadsMapEntry& e = g_tmap_cache[g_tmap_next++ % 128];
  e.key = key;
  e.map = *tm;
  e.used = true;
  return B200B_OK;
}

}  // namespace b200b

extern "C" int b200b_abi_version(void) { return B200B_ABI_VERSION; }
extern "C" void b200b_set_sm_limit(int sms) { b200b::g_sm_limit.store(sms > 0 ? sms : 0, std::memory_order_relaxed); }
extern "C" int b200b_get_sm_limit(void) { return b200b::g_sm_limit.load(std::memory_order_relaxed); }
extern "C" const char* b200b_last_error(void) { return b200b::g_err; }
extern "C" uint64_t b200b_launch_count(void) { return b200b::g_launches.load(std::memory_order_relaxed); }

// Per-kernel timing. begin(stream) records a start event; every later launch of this library
// records one more; end() synchronises and returns the number of timed launches. Entry i is the
// time between event i-1 and event i, i.e. kernel i plus whatever else ran on the stream between
// the two launches. Not thread safe against concurrent begin/end; meant for bench.py.
extern "C" int b200b_profile_begin(void* stream_) {
  using namespace b200b;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& pe : g_prof_events) cudaEventDestroy(pe.ev);
  g_prof_events.clear();
  ProfEvent pe;
  pe.what = "<begin>";
  cudaError_t e = cudaEventCreate(&pe.ev);
  if (e != cudaSuccess) {
    set_last_error("profile_begin: %s", cudaGetErrorString(e));
    return (int)e;
  }
  g_prof_stream = reinterpret_cast<cudaStream_t>(stream_);
  cudaEventRecord(pe.ev, g_prof_stream);
  g_prof_events.push_back(pe);
  g_prof_on = true;
  return B200B_OK;
}

extern "C" int b200b_profile_end(void) {
  using namespace b200b;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = false;
  if (g_prof_events.empty()) return 0;
  cudaEventSynchronize(g_prof_events.back().ev);
  return (int)g_prof_events.size() - 1;
}

// name and milliseconds of timed launch i in [0, b200b_profile_end()); returns NULL past the end
extern "C" const char* b200b_profile_entry(int i, float* ms) {
  using namespace b200b;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (i < 0 || i + 1 >= (int)g_prof_events.size()) return nullptr;
  float t = 0.f;
  cudaEventElapsedTime(&t, g_prof_events[i].ev, g_prof_events[i + 1].ev);
  if (ms) *ms = t;
  return g_prof_events[i + 1].what;
}
