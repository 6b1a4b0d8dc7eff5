// Status / last-error plumbing of the C ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace cb {
static thread_local char g_last_error[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: CUDA error %d (%s)", what, int(e), cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;  // B200
    cached[dev] = n;
  }
  return cached[dev];
}
}  // namespace cb

extern "C" const char* cb_last_error(void) { return cb::g_last_error; }
extern "C" int cb_abi_version(void) { return 1; }
