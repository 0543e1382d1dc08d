// Error reporting, per-device properties and the tensor-map encoder / cache for the C-ABI library.
#include <cuda.h>
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

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

// ---------------------------------------------------------------------------
// cuTensorMapEncodeTiled, resolved once (C++11 thread-safe static initialiser), + a cache of encoded maps
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn resolve_encode_fn() {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    return (EncodeTiledFn)ptr;
  return nullptr;
}
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = resolve_encode_fn();
  return fn;
}

namespace {
struct TmapKey {
  uint64_t v[11];
  bool operator==(const TmapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (uint64_t x : k.v) {
      h ^= x + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    }
    return (size_t)h;
  }
};
struct TmapVal {
  CUtensorMap m;
};
std::mutex g_tmap_mu;
std::unordered_map<TmapKey, TmapVal, TmapKeyHash> g_tmap_cache;
}  // namespace

int encode_tmap_bf16_sw128(void* out, const void* ptr, int rank, const uint64_t* gdim, const uint64_t* gstride,
                           const uint32_t* box, int l2_promotion_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_last_error("cuTensorMapEncodeTiled not available from the CUDA driver");
    return LSTC_ERR_DRIVER;
  }
  if (rank < 2 || rank > 3) {
    set_last_error("encode_tmap: rank %d unsupported", rank);
    return LSTC_ERR_INVALID_ARG;
  }
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.v[0] = (uint64_t)(uintptr_t)ptr;
  key.v[1] = (uint64_t)rank | ((uint64_t)l2_promotion_bytes << 8);
  for (int i = 0; i < rank; ++i) key.v[2 + i] = gdim[i];
  for (int i = 0; i < rank - 1; ++i) key.v[5 + i] = gstride[i];
  for (int i = 0; i < rank; ++i) key.v[7 + i] = box[i];
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      memcpy(out, &it->second.m, sizeof(CUtensorMap));
      return LSTC_OK;
    }
  }
  cuuint64_t gd[3] = {1, 1, 1}, gs[2] = {0, 0};
  cuuint32_t bx[3] = {1, 1, 1}, es[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gd[i] = gdim[i];
    bx[i] = box[i];
  }
  for (int i = 0; i < rank - 1; ++i) gs[i] = gstride[i];
  const CUtensorMapL2promotion promo = l2_promotion_bytes >= 256   ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                       : l2_promotion_bytes >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                                   : CU_TENSOR_MAP_L2_PROMOTION_NONE;
  TmapVal val;
  memset(&val, 0, sizeof(val));
  CUresult r = fn(&val.m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (CUresult %d) ptr=%p rank=%d dims=%llu,%llu,%llu strides=%llu,%llu "
                   "box=%u,%u,%u",
                   (int)r, ptr, rank, (unsigned long long)gd[0], (unsigned long long)gd[1], (unsigned long long)gd[2],
                   (unsigned long long)gs[0], (unsigned long long)gs[1], bx[0], bx[1], bx[2]);
    return LSTC_ERR_DRIVER;
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmap_cache.size() > 8192) g_tmap_cache.clear();  // bounded: shapes / pointers of a training job repeat
    g_tmap_cache.emplace(key, val);
  }
  memcpy(out, &val.m, sizeof(CUtensorMap));
  return LSTC_OK;
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
  rc |= lstc::set_rng_step_attention_tc(counter_dev);
  rc |= lstc::set_rng_step_attention_cls(counter_dev);
  rc |= lstc::set_rng_step_layernorm(counter_dev);
  rc |= lstc::set_rng_step_elementwise(counter_dev);
  rc |= lstc::set_rng_step_heads(counter_dev);
  if (rc != 0) lstc::set_last_error("lstc_set_rng_step: cudaMemcpyToSymbol failed");
  return rc == 0 ? LSTC_OK : LSTC_ERR_CUDA;
}
