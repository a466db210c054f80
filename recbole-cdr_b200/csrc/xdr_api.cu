// xdr_api.cu -- library-level entry points: version, thread-local error text, device info.
#include <stdarg.h>
#include <string.h>
#include <mutex>
#include <vector>
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

#ifndef XDR_EMU
// ---- persistent (co-resident) launches: one-time kernel setup + residency check, cached per (kernel, device) ----------
namespace {
struct CoopEntry {
  const void* kern;
  int dev, block;
  size_t smem_limit;   // dynamic shared-memory limit already set on the kernel
  size_t smem_checked; // (block, smem) pair the occupancy was computed for
  int ctas_per_sm;
};
std::mutex g_coop_mu;
std::vector<CoopEntry> g_coop;
int g_coop_enabled = 1;
}  // namespace

bool coop_enabled() { return g_coop_enabled != 0; }

int coop_prepare(const void* kern, int grid, int block, size_t smem, const char* name) {
  int dev = 0;
  XDR_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_coop_mu);
  CoopEntry* e = nullptr;
  for (auto& c : g_coop)
    if (c.kern == kern && c.dev == dev) e = &c;
  if (!e) {
    g_coop.push_back(CoopEntry{kern, dev, 0, 0, 0, 0});
    e = &g_coop.back();
  }
  if (smem > e->smem_limit) {
    XDR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    e->smem_limit = smem;
  }
  if (e->block != block || e->smem_checked != smem) {
    int n = 0;
    XDR_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, block, smem));
    e->block = block;
    e->smem_checked = smem;
    e->ctas_per_sm = n;
  }
  if ((long long)e->ctas_per_sm * sm_count() < grid) {
    set_error("%s: a grid of %d CTAs (%d threads, %zu B shared memory) cannot be co-resident on this device (%d per SM x %d "
              "SMs); its CTAs wait for each other, so it is not launched", name, grid, block, smem, e->ctas_per_sm, sm_count());
    return XDR_ERR_UNSUPPORTED;
  }
  return XDR_OK;
}
#endif  // !XDR_EMU

}  // namespace xdr

extern "C" {

#ifndef XDR_EMU
// 1 (default): persistent kernels go through cudaLaunchCooperativeKernel (driver-guaranteed co-residency); 0: plain launches
// (the static occupancy check still applies).  Returns the previous setting.
int xdr_set_coop_launch(int on) {
  const int prev = xdr::g_coop_enabled;
  xdr::g_coop_enabled = on ? 1 : 0;
  return prev;
}
#endif

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
