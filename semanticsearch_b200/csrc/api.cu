// Library-level plumbing: error strings, device introspection.
#include <vector>

#include "ss_common.cuh"

namespace ss {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  return SS_ERR_CUDA;
}

struct DevInfo {
  bool ok = false;
  int sms = 0, major = 0, minor = 0;
  size_t smem_optin = 0;
};
static DevInfo g_dev[64];

static const DevInfo& devinfo() {
  static DevInfo none;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return none;
  DevInfo& d = g_dev[dev];
  if (!d.ok) {
    int v = 0;
    cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&d.major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&d.minor, cudaDevAttrComputeCapabilityMinor, dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    d.smem_optin = static_cast<size_t>(v);
    d.ok = d.sms > 0;
  }
  return d;
}
int sm_count() {
  const int s = devinfo().sms;
  return s > 0 ? s : 148;
}
size_t smem_optin() {
  const size_t s = devinfo().smem_optin;
  return s > 0 ? s : 227 * 1024;
}

// ---- dominant-kernel timing (bench.py's roofline leg) --------------------------------------
// Between ss_profile_begin() and ss_profile_end() every launch of a kernel marked with
// ProfileScope is bracketed by a CUDA event pair on the launching stream; nothing synchronises
// until ss_profile_end() reads the pairs back.
struct ProfState {
  bool on = false;
  int cap = 0, used = 0, dropped = 0;
  std::vector<cudaEvent_t> ev;  // 2 * cap
};
static ProfState g_prof;

ProfileScope::ProfileScope(cudaStream_t st) : st_(st), slot_(-1) {
  if (!g_prof.on) return;
  if (g_prof.used >= g_prof.cap) {
    ++g_prof.dropped;
    return;
  }
  slot_ = g_prof.used++;
  cudaEventRecord(g_prof.ev[2 * slot_], st_);
}
ProfileScope::~ProfileScope() {
  if (slot_ >= 0) cudaEventRecord(g_prof.ev[2 * slot_ + 1], st_);
}

}  // namespace ss

extern "C" int ss_profile_begin(int max_records) {
  using namespace ss;
  if (max_records <= 0 || max_records > 65536) return fail(SS_ERR_INVALID_ARG, "ss_profile_begin: bad max_records");
  while (static_cast<int>(g_prof.ev.size()) < 2 * max_records) {
    cudaEvent_t e;
    SS_CUDA_CHECK(cudaEventCreate(&e));
    g_prof.ev.push_back(e);
  }
  g_prof.cap = max_records;
  g_prof.used = 0;
  g_prof.dropped = 0;
  g_prof.on = true;
  return SS_OK;
}

extern "C" int ss_profile_end(float* ms_out, int capacity, int* n_recorded, int* n_dropped) {
  using namespace ss;
  g_prof.on = false;
  int n = 0;
  for (int i = 0; i < g_prof.used; ++i) {
    SS_CUDA_CHECK(cudaEventSynchronize(g_prof.ev[2 * i + 1]));
    float ms = 0.f;
    SS_CUDA_CHECK(cudaEventElapsedTime(&ms, g_prof.ev[2 * i], g_prof.ev[2 * i + 1]));
    if (ms_out && n < capacity) ms_out[n] = ms;
    ++n;
  }
  if (n_recorded) *n_recorded = n;
  if (n_dropped) *n_dropped = g_prof.dropped;
  return SS_OK;
}

extern "C" int ss_version(void) { return 200; }
extern "C" const char* ss_last_error(void) { return ss::g_last_error.c_str(); }
extern "C" int ss_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* smem_per_block_optin) {
  const ss::DevInfo& d = ss::devinfo();
  if (!d.ok) return ss::fail(SS_ERR_CUDA, "ss_device_info: no usable CUDA device");
  if (sm_count) *sm_count = d.sms;
  if (cc_major) *cc_major = d.major;
  if (cc_minor) *cc_minor = d.minor;
  if (smem_per_block_optin) *smem_per_block_optin = d.smem_optin;
  return SS_OK;
}
