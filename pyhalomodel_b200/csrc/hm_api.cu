// hm_api.cu -- ABI version, error reporting and device probing for libhalob200.
#include <stdarg.h>
#include <string.h>

#include "hm_common.cuh"

static thread_local char g_err[512] = "";

void hm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* hm_last_error(void) { return g_err; }
extern "C" int hm_abi_version(void) { return HM_ABI_VERSION; }

extern "C" int hm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();   // clear the sticky "no device" error
    return 0;
  }
  return n;
}

// There is no CPU fallback anywhere in this library: without a device every compute entry point fails loudly.
int hm_check_device() {
  static int cached = 1;   // 1 = unknown
  if (cached == 1) cached = (hm_device_count() > 0) ? HM_OK : HM_ERR_NODEVICE;
  if (cached != HM_OK) hm_set_error("no CUDA device visible: libhalob200 has no CPU fallback");
  return cached;
}
