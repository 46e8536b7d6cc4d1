// Error state, version and device probing for the C ABI (include/witw_b200.h).
#include <cstring>

#include "common.cuh"

namespace witw {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

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

}  // namespace witw

extern "C" const char* witw_last_error(void) { return witw::g_err; }

extern "C" int witw_version(void) { return 200; }

// sizes of the argument structures, so that a binding can check its own declaration of them
extern "C" size_t witw_sizeof_sweep_args(void) { return sizeof(witw_sweep_args); }
extern "C" size_t witw_sizeof_finish_args(void) { return sizeof(witw_finish_args); }

extern "C" int witw_device_check(void) {
  int dev = 0, major = 0;
  WITW_CUDA(cudaGetDevice(&dev));
  WITW_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  WITW_REQUIRE(major == 10, WITW_ERR_DEVICE, "device %d has compute capability %d.x; libwitw_b200 is built for sm_100a only", dev, major);
  return WITW_OK;
}

// Opt-in L2 residency for one buffer (the query operand of a sweep, re-read once per 8 gallery items): an access-policy window on
// the stream, backed by the device's persisting-L2 set-aside.  bytes == 0 removes the window.
extern "C" int witw_stream_l2_window(const void* ptr_dev, size_t bytes, witw_stream_t stream) {
  cudaStreamAttrValue attr;
  memset(&attr, 0, sizeof(attr));
  if (bytes == 0 || ptr_dev == nullptr) {
    attr.accessPolicyWindow.num_bytes = 0;
    WITW_CUDA(cudaStreamSetAttribute(witw::as_stream(stream), cudaStreamAttributeAccessPolicyWindow, &attr));
    return WITW_OK;
  }
  int dev = 0, max_persist = 0, max_window = 0;
  WITW_CUDA(cudaGetDevice(&dev));
  WITW_CUDA(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
  WITW_CUDA(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev));
  WITW_REQUIRE(max_persist > 0 && max_window > 0, WITW_ERR_UNSUPPORTED, "witw_stream_l2_window: the device has no persisting L2 set-aside");
  static thread_local size_t set_aside[64] = {0};
  const size_t want = bytes < (size_t)max_persist ? bytes : (size_t)max_persist;
  if (dev >= 0 && dev < 64 && set_aside[dev] < want) {     // raising the set-aside is a device-wide setting: done once, never lowered here
    WITW_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
    set_aside[dev] = want;
  }
  const size_t win = bytes < (size_t)max_window ? bytes : (size_t)max_window;
  attr.accessPolicyWindow.base_ptr = const_cast<void*>(ptr_dev);
  attr.accessPolicyWindow.num_bytes = win;
  attr.accessPolicyWindow.hitRatio = want >= win ? 1.0f : (float)((double)want / (double)win);
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  WITW_CUDA(cudaStreamSetAttribute(witw::as_stream(stream), cudaStreamAttributeAccessPolicyWindow, &attr));
  return WITW_OK;
}
