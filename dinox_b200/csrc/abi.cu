// Library-level plumbing of the C ABI: version, error string, device check, launch counter.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <atomic>

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

int check_launch(const char* what, cudaStream_t) {
  ++g_launches;
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
}
