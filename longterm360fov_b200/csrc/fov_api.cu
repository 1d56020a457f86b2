// Error reporting and library-level queries of libfov360.
#include <stdarg.h>

#include "fov_common.cuh"

static thread_local char g_err[512] = "";
unsigned long long g_fov_launches = 0;

void fov_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* fov_last_error(void) { return g_err; }
extern "C" int fov_version(void) { return 100; }
extern "C" unsigned long long fov_launch_count(void) { return g_fov_launches; }

extern "C" int fov_device_is_sm100(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}
