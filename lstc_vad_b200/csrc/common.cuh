// Shared device/host helpers for the lstc_vad_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

// ---- status codes returned through the C-ABI (include/lstc_vad_b200.h) ----
#define LSTC_OK 0
#define LSTC_ERR_INVALID_ARG 1
#define LSTC_ERR_CUDA 2
#define LSTC_ERR_UNSUPPORTED 3
#define LSTC_ERR_DRIVER 4

namespace lstc {

void set_last_error(const char* fmt, ...);

// every matrix operand is read / written with 16-byte vector accesses
#define LSTC_ALIGNED16(p) ((reinterpret_cast<uintptr_t>(p) & 15u) == 0)
#define LSTC_CHECK_ARG(cond, ...)                                   \
  do {                                                              \
    if (!(cond)) {                                                  \
      ::lstc::set_last_error(__VA_ARGS__);                          \
      return LSTC_ERR_INVALID_ARG;                                  \
    }                                                               \
  } while (0)

#define LSTC_CHECK_CUDA(expr)                                                        \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ::lstc::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                    \
      return LSTC_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

#define LSTC_CHECK_LAUNCH() LSTC_CHECK_CUDA(cudaGetLastError())

int num_sms();  // SM count of the current device (cached per device)

// Encodes (or fetches from a process-wide cache keyed on every argument) a SWIZZLE_128B bf16 tiled tensor map of rank
// 2 or 3: gdim[] elements (innermost first), gstride[] BYTES of dims 1.. , box[] elements, zero fill out of bounds.
// Thread-safe (nn.DataParallel replica threads call the launchers concurrently).  `out` is the driver's 128-byte
// CUtensorMap.
int encode_tmap_bf16_sw128(void* out, const void* ptr, int rank, const uint64_t* gdim, const uint64_t* gstride,
                           const uint32_t* box, int l2_promotion_bytes);

// per translation unit setters of the device-side dropout step pointer (see rng_step() below)
int set_rng_step_gemm(const void* p);
int set_rng_step_attention(const void* p);
int set_rng_step_attention_tc(const void* p);
int set_rng_step_attention_cls(const void* p);
int set_rng_step_layernorm(const void* p);
int set_rng_step_elementwise(const void* p);
int set_rng_step_heads(const void* p);

// ---------------------------------------------------------------------------
// bf16 <-> fp32 packing helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ float bf16lo_to_f32(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16hi_to_f32(uint32_t packed) { return __uint_as_float(packed & 0xffff0000u); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16lo_to_f32(v.x); f[1] = bf16hi_to_f32(v.x);
  f[2] = bf16lo_to_f32(v.y); f[3] = bf16hi_to_f32(v.y);
  f[4] = bf16lo_to_f32(v.z); f[5] = bf16hi_to_f32(v.z);
  f[6] = bf16lo_to_f32(v.w); f[7] = bf16hi_to_f32(v.w);
}

__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  v.x = pack_bf16x2(f[0], f[1]);
  v.y = pack_bf16x2(f[2], f[3]);
  v.z = pack_bf16x2(f[4], f[5]);
  v.w = pack_bf16x2(f[6], f[7]);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------
// Counter-based dropout RNG: Philox4x32 with 7 rounds (the shortest variant that is
// Crush-resistant in Salmon et al., SC'11; dropout masks need no larger margin and
// the rounds are pure ALU work inside bandwidth- and issue-bound kernels).  One call
// yields 128 random bits = eight 16-bit lanes, i.e. the keep/drop decision for 8
// consecutive elements.
// The mask of element (row, col) of a [rows, ld8*8] grid depends only on
// (seed, offset, row*ld8 + col/8) so forward and backward kernels regenerate
// the same mask without storing it (reference keeps masks inside autograd:
// nn.Dropout at models/MultiHeadAttention.py:46,50, models/FFN.py:11).
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

constexpr int PHILOX_ROUNDS = 7;
__host__ __device__ __forceinline__ void philox4x32(uint64_t seed, uint64_t offset, uint64_t idx,
                                                    uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < PHILOX_ROUNDS; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// Same generator with the round keys expanded once (they depend on the seed only): kernels that draw many masks keep
// the 2*PHILOX_ROUNDS keys in (uniform) registers instead of re-deriving them per call.
struct PhiloxKeys {
  uint32_t k0[PHILOX_ROUNDS], k1[PHILOX_ROUNDS];
};
__host__ __device__ __forceinline__ PhiloxKeys philox_keys(uint64_t seed) {
  PhiloxKeys k;
  uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < PHILOX_ROUNDS; ++r) {
    k.k0[r] = a;
    k.k1[r] = b;
    a += 0x9E3779B9u;
    b += 0xBB67AE85u;
  }
  return k;
}
__host__ __device__ __forceinline__ void philox4x32_keyed(const PhiloxKeys& k, uint64_t offset, uint64_t idx,
                                                          uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
#pragma unroll
  for (int r = 0; r < PHILOX_ROUNDS; ++r) philox_round(c, k.k0[r], k.k1[r]);
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// threshold in 1/65536 units: element is DROPPED when its 16-bit lane < thr.
__host__ __device__ __forceinline__ uint32_t dropout_threshold16(float p) {
  float t = p * 65536.0f + 0.5f;
  if (t < 0.f) t = 0.f;
  if (t > 65535.f) t = 65535.f;
  return (uint32_t)t;
}

// raw form of dropout_keep8 for kernels that select per element: element 8*idx8 + j is KEPT when
// dropout_kept(r, j, thr16), with r the four words this returns
__host__ __device__ __forceinline__ void dropout_rand8(uint64_t seed, uint64_t offset, uint64_t idx8, uint32_t (&r)[4]) {
  philox4x32(seed, offset, idx8, r);
}
__host__ __device__ __forceinline__ bool dropout_kept(const uint32_t (&r)[4], int j, uint32_t thr16) {
  return ((j & 1) ? (r[j >> 1] >> 16) : (r[j >> 1] & 0xffffu)) >= thr16;
}

// 8-bit keep mask (bit j set => element 8*idx8 + j is kept)
__host__ __device__ __forceinline__ uint32_t dropout_keep8(uint64_t seed, uint64_t offset, uint64_t idx8,
                                                           uint32_t thr16) {
  uint32_t r[4];
  philox4x32(seed, offset, idx8, r);
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    m |= ((r[j] & 0xffffu) >= thr16 ? 1u : 0u) << (2 * j);
    m |= ((r[j] >> 16) >= thr16 ? 1u : 0u) << (2 * j + 1);
  }
  return m;
}

}  // namespace lstc

// ---------------------------------------------------------------------------
// Optional device-resident dropout step counter.  A CUDA graph bakes every kernel's (seed, offset) arguments by value;
// when a counter has been registered with lstc_set_rng_step(), each kernel adds its current value to `offset`, so a
// replayed graph draws fresh masks every step (the harness bumps the counter inside the graph).  Without -rdc every
// translation unit owns a copy of the constant; lstc_set_rng_step() sets them all.
// ---------------------------------------------------------------------------
static __constant__ const unsigned long long* g_rng_step_ptr = nullptr;
__device__ __forceinline__ uint64_t rng_step() {
  const unsigned long long* p = g_rng_step_ptr;
  return p != nullptr ? (uint64_t)(*p) : 0ull;
}
#define LSTC_DEFINE_RNG_STEP_SETTER(name)                                                        \
  int lstc::name(const void* p) {                                                                \
    return cudaMemcpyToSymbol(g_rng_step_ptr, &p, sizeof(p)) == cudaSuccess ? LSTC_OK : LSTC_ERR_CUDA; \
  }
