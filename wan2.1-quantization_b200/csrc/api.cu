// Library-level entry points and the error/device plumbing shared by every translation unit.
#include "common.cuh"
#include <string.h>

namespace b200q {

static thread_local char g_err[512] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void clear_error() { g_err[0] = 0; }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace b200q

using namespace b200q;

extern "C" int b200q_version(void) { return (0 << 16) | (1 << 8) | 0; }

extern "C" const char* b200q_last_error(void) { return g_err; }

extern "C" int b200q_device_info(int* sm_count_out, int* cc_major, int* cc_minor) {
  clear_error();
  int dev = 0, major = 0, minor = 0, sms = 0;
  B200Q_CUDA_OK(cudaGetDevice(&dev));
  B200Q_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  B200Q_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  B200Q_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (sm_count_out) *sm_count_out = sms;
  if (cc_major) *cc_major = major;
  if (cc_minor) *cc_minor = minor;
  B200Q_REQUIRE(major == 10 && minor == 0, B200Q_ERR_UNSUPPORTED,
                "libb200q is built for sm_100a only; device is sm_%d%d", major, minor);
  return B200Q_OK;
}
