// Error state, version and device probing for the C ABI (include/witw_b200.h).
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
