// Error reporting and per-device properties for the C-ABI library.
#include <stdarg.h>

#include "common.cuh"
#include "../../include/lstc_vad_b200.h"

namespace lstc {

static thread_local char g_last_error[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int num_sms() {
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

}  // namespace lstc

extern "C" int lstc_abi_version(void) { return LSTC_ABI_VERSION; }
extern "C" const char* lstc_last_error(void) { return lstc::g_last_error; }

// Registers (or clears, with NULL) the device-resident dropout step counter on the CURRENT device: every kernel that
// draws a dropout mask adds *counter to its `offset` argument.
extern "C" int lstc_set_rng_step(const void* counter_dev) {
  int rc = 0;
  rc |= lstc::set_rng_step_gemm(counter_dev);
  rc |= lstc::set_rng_step_attention(counter_dev);
  rc |= lstc::set_rng_step_attention_cls(counter_dev);
  rc |= lstc::set_rng_step_layernorm(counter_dev);
  rc |= lstc::set_rng_step_elementwise(counter_dev);
  rc |= lstc::set_rng_step_heads(counter_dev);
  if (rc != 0) lstc::set_last_error("lstc_set_rng_step: cudaMemcpyToSymbol failed");
  return rc == 0 ? LSTC_OK : LSTC_ERR_CUDA;
}
