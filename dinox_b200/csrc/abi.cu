// Library-level plumbing of the C ABI: version, error string, device check, launch counter.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <mutex>

namespace dinox {

static thread_local char g_err[512] = "";
// process-wide: backward kernels are launched from autograd worker threads
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- launch trace (diagnostics): a one-thread kernel behind every launch writes %globaltimer into the caller's
// buffer, so the END time of every kernel of a step - also inside a replayed CUDA graph, across its streams - can be
// read back and laid out as a timeline (tools/trace_step.py).  Off unless dinox_trace_begin() was called.
constexpr int kTraceMax = 4096;
static std::mutex g_trace_mu;
static unsigned long long* g_trace_buf = nullptr;
static int g_trace_cap = 0, g_trace_n = 0;
static char g_trace_name[kTraceMax][48];
static unsigned long long g_trace_stream[kTraceMax];

__global__ void trace_stamp_kernel(unsigned long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *slot = t;
}

static void trace_launch(const char* what, cudaStream_t stream) {
  std::lock_guard<std::mutex> lk(g_trace_mu);
  if (!g_trace_buf || g_trace_n >= g_trace_cap) return;
  const int i = g_trace_n++;
  strncpy(g_trace_name[i], what, sizeof(g_trace_name[i]) - 1);
  g_trace_name[i][sizeof(g_trace_name[i]) - 1] = 0;
  g_trace_stream[i] = (unsigned long long)(uintptr_t)stream;
  trace_stamp_kernel<<<1, 1, 0, stream>>>(g_trace_buf + i);
}

int check_launch(const char* what, cudaStream_t stream) {
  ++g_launches;
  if (g_trace_buf) trace_launch(what, stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return DINOX_E_CUDA;
  }
  return DINOX_OK;
}

static int g_cc_major[64];
static int g_sms[64];
static bool g_probed[64];

static int probe(int* dev_out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 64) {
    set_error("cudaGetDevice failed: %s", cudaGetErrorString(e));
    return DINOX_E_CUDA;
  }
  if (!g_probed[dev]) {
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) {
      set_error("cudaGetDeviceProperties failed: %s", cudaGetErrorString(e));
      return DINOX_E_CUDA;
    }
    g_cc_major[dev] = p.major;
    g_sms[dev] = p.multiProcessorCount;
    g_probed[dev] = true;
  }
  *dev_out = dev;
  return DINOX_OK;
}

int require_sm100(void) {
  int dev;
  int rc = probe(&dev);
  if (rc != DINOX_OK) return rc;
  if (g_cc_major[dev] != 10) {
    set_error("dinox_b200 needs an sm_100 (B200) device; device %d is sm_%d0 - there is no fallback",
              dev, g_cc_major[dev]);
    return DINOX_E_ARCH;
  }
  return DINOX_OK;
}

int num_sms(void) {
  int dev;
  if (probe(&dev) != DINOX_OK) return 148;
  return g_sms[dev] > 0 ? g_sms[dev] : 148;
}

}  // namespace dinox

extern "C" {
int dinox_version(void) { return 100; }
const char* dinox_last_error_string(void) { return dinox::g_err; }
int dinox_device_check(void) { return dinox::require_sm100(); }
int64_t dinox_launch_count(void) { return dinox::g_launches.load(); }
void dinox_launch_count_reset(void) { dinox::g_launches.store(0); }

int dinox_trace_begin(uint64_t* device_slots, int capacity) {
  std::lock_guard<std::mutex> lk(dinox::g_trace_mu);
  if (!device_slots || capacity <= 0) { dinox::set_error("trace_begin: bad arguments"); return DINOX_E_BADARG; }
  dinox::g_trace_buf = reinterpret_cast<unsigned long long*>(device_slots);
  dinox::g_trace_cap = capacity < dinox::kTraceMax ? capacity : dinox::kTraceMax;
  dinox::g_trace_n = 0;
  return DINOX_OK;
}
int dinox_trace_end(void) {
  std::lock_guard<std::mutex> lk(dinox::g_trace_mu);
  dinox::g_trace_buf = nullptr;
  return dinox::g_trace_n;
}
const char* dinox_trace_name(int i) { return (i >= 0 && i < dinox::g_trace_n) ? dinox::g_trace_name[i] : ""; }
uint64_t dinox_trace_stream(int i) { return (i >= 0 && i < dinox::g_trace_n) ? dinox::g_trace_stream[i] : 0; }
}
