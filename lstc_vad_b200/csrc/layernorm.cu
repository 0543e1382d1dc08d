// LayerNorm forward / backward over the last dimension, fp32 statistics, bf16 or fp32 I/O.
// Replaces nn.LayerNorm(d_model, eps=1e-6) at models/Encoder.py:31,48-49,
// models/MultiHeadAttention.py:47,125-126 and models/FFN.py:10,20-21 (and its autograd backward).
//
// One CTA walks over rows (grid-stride); thread t owns the same 8*NV columns of every row, so
//   - each row is one fully coalesced 16-byte-per-thread read and write,
//   - the per-column dgamma / dbeta sums of the backward stay in registers for the whole kernel and
//     are flushed once per CTA into a partial buffer that a second tiny kernel reduces
//     (deterministic, no atomics).
#include "common.cuh"
#include "../../include/lstc_vad_b200.h"

namespace lstc {
namespace ln {

constexpr int MAX_WARPS = 8;

template <bool IS_F32>
__device__ __forceinline__ void load8(const void* base, int64_t off, float (&f)[8]) {
  if (IS_F32) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off));
    const float4 b = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off + 4));
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off));
    unpack8(v, f);
  }
}
template <bool IS_F32>
__device__ __forceinline__ void store8(void* base, int64_t off, const float (&f)[8]) {
  if (IS_F32) {
    float* p = reinterpret_cast<float*>(base) + off;
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + off) = pack8(f);
  }
}

// block-wide sum of (a, b); `buf` is a [2][MAX_WARPS][2] smem scratch, `phase` alternates per call
__device__ __forceinline__ void block_sum2(float& a, float& b, float (*buf)[MAX_WARPS][2], int& phase) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) {
    buf[phase][warp][0] = a;
    buf[phase][warp][1] = b;
  }
  __syncthreads();
  float ra = 0.f, rb = 0.f;
  for (int i = 0; i < nw; ++i) {
    ra += buf[phase][i][0];
    rb += buf[phase][i][1];
  }
  a = ra;
  b = rb;
  phase ^= 1;
}

template <int NV, bool X_F32, bool Y_F32>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const void* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, void* __restrict__ y,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                     int64_t rows, int D, float eps) {
  __shared__ float buf[2][MAX_WARPS][2];
  int phase = 0;
  const int tid = threadIdx.x;
  float gm[NV][8], bt[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = (v * blockDim.x + tid) * 8;
    if (c < D) {
      load8<true>(gamma, c, gm[v]);
      load8<true>(beta, c, bt[v]);
    }
  }
  const float invD = 1.0f / (float)D;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    float xv[NV][8];
    float s = 0.f, dummy = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = (v * blockDim.x + tid) * 8;
      if (c < D) {
        load8<X_F32>(x, row * D + c, xv[v]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += xv[v][j];
      }
    }
    block_sum2(s, dummy, buf, phase);
    const float mean = s * invD;
    float q = 0.f;
    dummy = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = (v * blockDim.x + tid) * 8;
      if (c < D) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = xv[v][j] - mean;
          q += d * d;
        }
      }
    }
    block_sum2(q, dummy, buf, phase);
    const float rstd = rsqrtf(q * invD + eps);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = (v * blockDim.x + tid) * 8;
      if (c < D) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (xv[v][j] - mean) * rstd * gm[v][j] + bt[v][j];
        store8<Y_F32>(y, row * D + c, o);
      }
    }
    if (tid == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
  }
}

template <int NV, bool DY_F32, bool X_F32, bool DX_F32>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ mean_in, const float* __restrict__ rstd_in, void* __restrict__ dx,
              __nv_bfloat16* __restrict__ dx_drop, float drop_scale, uint32_t drop_thr16, uint64_t seed,
              uint64_t offset, float* __restrict__ partial /*[grid][2][D]*/, int64_t rows, int D) {
  __shared__ float buf[2][MAX_WARPS][2];
  int phase = 0;
  const int tid = threadIdx.x;
  float gm[NV][8], dg[NV][8], db[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = (v * blockDim.x + tid) * 8;
    if (c < D) load8<true>(gamma, c, gm[v]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dg[v][j] = 0.f;
      db[v][j] = 0.f;
    }
  }
  const float invD = 1.0f / (float)D;
  const int64_t ld8 = D >> 3;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const float mean = __ldg(mean_in + row), rstd = __ldg(rstd_in + row);
    float xh[NV][8], dyg[NV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = (v * blockDim.x + tid) * 8;
      if (c < D) {
        float dyv[8];
        load8<X_F32>(x, row * D + c, xh[v]);
        load8<DY_F32>(dy, row * D + c, dyv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[v][j] = (xh[v][j] - mean) * rstd;
          dg[v][j] += dyv[j] * xh[v][j];
          db[v][j] += dyv[j];
          dyg[v][j] = dyv[j] * gm[v][j];
          s1 += dyg[v][j];
          s2 += dyg[v][j] * xh[v][j];
        }
      }
    }
    block_sum2(s1, s2, buf, phase);
    s1 *= invD;
    s2 *= invD;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = (v * blockDim.x + tid) * 8;
      if (c < D) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (dyg[v][j] - s1 - xh[v][j] * s2);
        store8<DX_F32>(dx, row * D + c, o);
        if (dx_drop != nullptr) {
          const uint32_t keep = dropout_keep8(seed, offset, (uint64_t)(row * ld8 + (c >> 3)), drop_thr16);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = ((keep >> j) & 1u) ? o[j] * drop_scale : 0.f;
          store8<false>(dx_drop, row * D + c, o);
        }
      }
    }
  }
  float* pg = partial + (int64_t)blockIdx.x * 2 * D;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = (v * blockDim.x + tid) * 8;
    if (c < D) {
      store8<true>(pg, c, dg[v]);
      store8<true>(pg + D, c, db[v]);
    }
  }
}

// out[k][c] = sum_b partial[b][k][c], k in {0,1}
__global__ void ln_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int D,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= 2 * D) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += partial[(int64_t)b * 2 * D + c];
  if (c < D) dgamma[c] = s; else dbeta[c - D] = s;
}

static int threads_for(int64_t D, int& nv) {
  // threads*8*nv >= D, threads a multiple of 32, <= 256
  nv = (int)((D + 2047) / 2048);
  if (nv == 3) nv = 4;
  int threads = (int)((D / 8 + nv - 1) / nv);
  threads = ((threads + 31) / 32) * 32;
  return threads;
}

static int64_t bwd_grid(int64_t rows) {
  int64_t g = (int64_t)num_sms() * 8;
  if (g > rows) g = rows;
  if (g < 1) g = 1;
  return g;
}

}  // namespace ln
}  // namespace lstc

using namespace lstc;

extern "C" int lstc_layernorm_fwd(const void* x, int x_is_f32, const float* gamma, const float* beta, void* y,
                                  int y_is_f32, float* mean, float* rstd, int64_t rows, int64_t D, float eps,
                                  void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LSTC_CHECK_ARG(x && gamma && beta && y && mean && rstd, "lstc_layernorm_fwd: null pointer");
  LSTC_CHECK_ARG(D > 0 && D % 8 == 0 && D <= 8192, "lstc_layernorm_fwd: D=%lld must be a multiple of 8, <= 8192",
                 (long long)D);
  if (rows == 0) return LSTC_OK;
  int nv;
  const int threads = ln::threads_for(D, nv);
  int64_t grid = (int64_t)num_sms() * 8;
  if (grid > rows) grid = rows;
#define LSTC_LN_FWD(NV)                                                                                         \
  do {                                                                                                          \
    if (x_is_f32 && y_is_f32)                                                                                   \
      ln::ln_fwd_kernel<NV, true, true><<<(unsigned)grid, threads, 0, stream>>>(x, gamma, beta, y, mean, rstd, \
                                                                                 rows, (int)D, eps);           \
    else if (x_is_f32)                                                                                          \
      ln::ln_fwd_kernel<NV, true, false><<<(unsigned)grid, threads, 0, stream>>>(x, gamma, beta, y, mean, rstd, \
                                                                                  rows, (int)D, eps);          \
    else if (y_is_f32)                                                                                          \
      ln::ln_fwd_kernel<NV, false, true><<<(unsigned)grid, threads, 0, stream>>>(x, gamma, beta, y, mean, rstd, \
                                                                                  rows, (int)D, eps);          \
    else                                                                                                        \
      ln::ln_fwd_kernel<NV, false, false><<<(unsigned)grid, threads, 0, stream>>>(x, gamma, beta, y, mean,     \
                                                                                   rstd, rows, (int)D, eps);   \
  } while (0)
  if (nv == 1) LSTC_LN_FWD(1);
  else if (nv == 2) LSTC_LN_FWD(2);
  else LSTC_LN_FWD(4);
#undef LSTC_LN_FWD
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

extern "C" int64_t lstc_layernorm_bwd_workspace(int64_t rows, int64_t D) {
  return ln::bwd_grid(rows) * 2 * D * (int64_t)sizeof(float);
}

extern "C" int lstc_layernorm_bwd(const void* dy, int dy_is_f32, const void* x, int x_is_f32, const float* gamma,
                                  const float* mean, const float* rstd, void* dx, int dx_is_f32, void* dx_drop,
                                  float drop_p, uint64_t seed, uint64_t offset, float* dgamma, float* dbeta,
                                  void* workspace, int64_t rows, int64_t D, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LSTC_CHECK_ARG(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && workspace,
                 "lstc_layernorm_bwd: null pointer");
  LSTC_CHECK_ARG(D > 0 && D % 8 == 0 && D <= 8192, "lstc_layernorm_bwd: D=%lld must be a multiple of 8, <= 8192",
                 (long long)D);
  LSTC_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "lstc_layernorm_bwd: drop_p out of range");
  if (rows == 0) {
    LSTC_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, D * sizeof(float), stream));
    LSTC_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, D * sizeof(float), stream));
    return LSTC_OK;
  }
  int nv;
  const int threads = ln::threads_for(D, nv);
  const int64_t grid = ln::bwd_grid(rows);
  __nv_bfloat16* dd = (drop_p > 0.f) ? (__nv_bfloat16*)dx_drop : nullptr;
  const float dscale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const uint32_t thr = dropout_threshold16(drop_p);
  float* partial = (float*)workspace;
#define LSTC_LN_BWD_K(NV, A, B, C)                                                                             \
  ln::ln_bwd_kernel<NV, A, B, C><<<(unsigned)grid, threads, 0, stream>>>(dy, x, gamma, mean, rstd, dx, dd,    \
                                                                         dscale, thr, seed, offset, partial,  \
                                                                         rows, (int)D)
#define LSTC_LN_BWD(NV)                                             \
  do {                                                              \
    if (dy_is_f32 && x_is_f32 && dx_is_f32) LSTC_LN_BWD_K(NV, true, true, true);         \
    else if (dy_is_f32 && !x_is_f32 && !dx_is_f32) LSTC_LN_BWD_K(NV, true, false, false); \
    else if (!dy_is_f32 && x_is_f32 && dx_is_f32) LSTC_LN_BWD_K(NV, false, true, true);   \
    else if (!dy_is_f32 && !x_is_f32 && !dx_is_f32) LSTC_LN_BWD_K(NV, false, false, false); \
    else {                                                          \
      set_last_error("lstc_layernorm_bwd: unsupported dtype combination (dx must match x)"); \
      return LSTC_ERR_UNSUPPORTED;                                  \
    }                                                               \
  } while (0)
  if (nv == 1) LSTC_LN_BWD(1);
  else if (nv == 2) LSTC_LN_BWD(2);
  else LSTC_LN_BWD(4);
#undef LSTC_LN_BWD
#undef LSTC_LN_BWD_K
  LSTC_CHECK_LAUNCH();
  ln::ln_bwd_finalize_kernel<<<(unsigned)((2 * D + 255) / 256), 256, 0, stream>>>(partial, (int)grid, (int)D, dgamma,
                                                                                dbeta);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}
