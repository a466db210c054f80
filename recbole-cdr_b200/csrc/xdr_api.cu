// xdr_api.cu -- library-level entry points: version, thread-local error text, device info.
#include <stdarg.h>
#include <string.h>
#include "xdr_common.cuh"

namespace xdr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached_dev = dev;
    cached_sms = n;
  }
  return cached_sms;
}

}  // namespace xdr

extern "C" {

int xdr_version(void) { return XDR_VERSION; }

const char* xdr_last_error(void) { return xdr::g_err; }

int xdr_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host) {
  int dev = 0;
  XDR_CUDA_OK(cudaGetDevice(&dev));
  int sms = 0, maj = 0, min = 0;
  XDR_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  XDR_CUDA_OK(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  XDR_CUDA_OK(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count_host) *sm_count_host = sms;
  if (cc_major_host) *cc_major_host = maj;
  if (cc_minor_host) *cc_minor_host = min;
  return XDR_OK;
}

size_t xdr_workspace_bytes(void) { return xdr::kWsBytes; }

}  // extern "C"
