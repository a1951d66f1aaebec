// Host-side plumbing shared by every entry point: error strings, device query.
#include <stdarg.h>
#include <string.h>

#include "snb_common.cuh"

namespace snb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return (int)e;
}

int num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("no CUDA device");
    return -1;
  }
  if (dev != cached_dev) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
      set_error("cudaGetDeviceProperties failed");
      return -1;
    }
    if (p.major != 10) {
      set_error("libsnb is built for sm_100a only; device is sm_%d%d", p.major, p.minor);
      return -1;
    }
    cached = p.multiProcessorCount;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace snb

extern "C" int snb_version(void) { return SNB_VERSION; }
extern "C" const char* snb_last_error(void) { return snb::g_err; }
extern "C" int snb_device_sms(void) { return snb::num_sms(); }
