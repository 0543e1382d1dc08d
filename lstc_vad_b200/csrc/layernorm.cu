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
// streaming (evict-first) stores: the row is not re-read by this kernel and is larger than what L2 can keep for the
// consumer anyway; keeping it from displacing the rows still to be read helps the read stream
template <bool IS_F32>
__device__ __forceinline__ void store8(void* base, int64_t off, const float (&f)[8]) {
  if (IS_F32) {
    float* p = reinterpret_cast<float*>(base) + off;
    __stcs(reinterpret_cast<float4*>(p), make_float4(f[0], f[1], f[2], f[3]));
    __stcs(reinterpret_cast<float4*>(p + 4), make_float4(f[4], f[5], f[6], f[7]));
  } else {
    __stcs(reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + off), pack8(f));
  }
}

// block-wide sums of N values at once; `buf` is a [2][MAX_WARPS][N] smem scratch, `phase` alternates per call so a
// single barrier per reduction suffices
template <int N>
__device__ __forceinline__ void block_sum_n(float (&v)[N], float (*buf)[MAX_WARPS][N], int& phase) {
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) buf[phase][warp][i] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = 0.f;
  for (int w = 0; w < nw; ++w) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] += buf[phase][w][i];
  }
  phase ^= 1;
}

constexpr int RPI = 2;   // forward: rows per CTA iteration (half the barriers per row)
constexpr int NSTG = 4;  // depth of the per-thread cp.async row ring

// Every thread stages ITS OWN 16/32-byte chunks of the next NSTG-1 iterations in shared memory with cp.async and is
// the only reader of those bytes, so the ring needs no barrier: cp.async.wait_group on the thread's own groups is
// enough.  This keeps ~3 rows per CTA in flight without holding them in registers.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
template <bool IS_F32>
__device__ __forceinline__ void stage8(uint32_t dst, const void* base, int64_t off) {
  if (IS_F32) {
    cp_async16(dst, reinterpret_cast<const float*>(base) + off);
    cp_async16(dst + 16, reinterpret_cast<const float*>(base) + off + 4);
  } else {
    cp_async16(dst, reinterpret_cast<const __nv_bfloat16*>(base) + off);
  }
}
template <bool IS_F32>
__device__ __forceinline__ void unstage8(uint32_t src, float (&f)[8]) {
  uint4 a;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(src));
  if (IS_F32) {
    uint4 b;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(src + 16));
    f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
    f[4] = __uint_as_float(b.x); f[5] = __uint_as_float(b.y); f[6] = __uint_as_float(b.z); f[7] = __uint_as_float(b.w);
  } else {
    unpack8(a, f);
  }
}

template <int NV, bool X_F32, bool Y_F32>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const void* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, void* __restrict__ y,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                     int64_t rows, int D, float eps) {
  __shared__ float buf[2][MAX_WARPS][RPI];
  extern __shared__ __align__(16) uint8_t ring_raw[];
  constexpr int XB = X_F32 ? 32 : 16;
  int phase = 0;
  const int tid = threadIdx.x;
  const uint32_t ring = smem_addr(ring_raw);
  const uint32_t stage_bytes = RPI * NV * blockDim.x * XB;
  auto slot = [&](int st, int r, int v) { return ring + st * stage_bytes + ((r * NV + v) * blockDim.x + tid) * XB; };
  auto issue = [&](int64_t row0, int st) {
    if (row0 < rows) {
#pragma unroll
      for (int r = 0; r < RPI; ++r) {
        if (row0 + r < rows) {
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const int c = (v * blockDim.x + tid) * 8;
            if (c < D) stage8<X_F32>(slot(st, r, v), x, (row0 + r) * D + c);
          }
        }
      }
    }
    cp_async_commit();
  };
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
  const int64_t step = (int64_t)gridDim.x * RPI;
#pragma unroll
  for (int i = 0; i < NSTG - 1; ++i) issue((int64_t)blockIdx.x * RPI + i * step, i);
  int st = 0;
  for (int64_t row0 = (int64_t)blockIdx.x * RPI; row0 < rows; row0 += step) {
    cp_async_wait<NSTG - 2>();
    float xv[RPI][NV][8];
    float s[RPI];
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      s[r] = 0.f;
      if (row0 + r < rows) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c = (v * blockDim.x + tid) * 8;
          if (c < D) {
            unstage8<X_F32>(slot(st, r, v), xv[r][v]);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[r] += xv[r][v][j];
          }
        }
      }
    }
    issue(row0 + (NSTG - 1) * step, st == 0 ? NSTG - 1 : st - 1);  // refills the slot consumed one iteration ago
    st = (st + 1 == NSTG) ? 0 : st + 1;
    block_sum_n<RPI>(s, buf, phase);
    float mean[RPI], q[RPI];
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      mean[r] = s[r] * invD;
      q[r] = 0.f;
      if (row0 + r < rows) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c = (v * blockDim.x + tid) * 8;
          if (c < D) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float d = xv[r][v][j] - mean[r];
              q[r] += d * d;
            }
          }
        }
      }
    }
    block_sum_n<RPI>(q, buf, phase);
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      if (row0 + r < rows) {
        const float rstd = rsqrtf(q[r] * invD + eps);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c = (v * blockDim.x + tid) * 8;
          if (c < D) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = (xv[r][v][j] - mean[r]) * rstd * gm[v][j] + bt[v][j];
            store8<Y_F32>(y, (row0 + r) * D + c, o);
          }
        }
        if (tid == 0) {
          mean_out[row0 + r] = mean[r];
          rstd_out[row0 + r] = rstd;
        }
      }
    }
  }
  cp_async_wait<0>();
}

template <int NV, bool DY_F32, bool X_F32, bool DX_F32>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ mean_in, const float* __restrict__ rstd_in, void* __restrict__ dx,
              __nv_bfloat16* __restrict__ dx_drop, float drop_scale, uint32_t drop_thr16, uint64_t seed,
              uint64_t offset, float* __restrict__ partial /*[grid][3][D]*/, int want_dxsum, int64_t rows, int D) {
  offset += rng_step();
  __shared__ float buf[2][MAX_WARPS][2];
  extern __shared__ __align__(16) uint8_t ring_raw[];
  constexpr int XB = X_F32 ? 32 : 16, YB = DY_F32 ? 32 : 16;
  int phase = 0;
  const int tid = threadIdx.x;
  const uint32_t ring = smem_addr(ring_raw);
  const uint32_t stage_bytes = NV * blockDim.x * (XB + YB);
  auto x_slot = [&](int st, int v) { return ring + st * stage_bytes + (v * blockDim.x + tid) * XB; };
  auto dy_slot = [&](int st, int v) { return ring + st * stage_bytes + NV * blockDim.x * XB + (v * blockDim.x + tid) * YB; };
  auto issue = [&](int64_t row, int st) {
    if (row < rows) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int c = (v * blockDim.x + tid) * 8;
        if (c < D) {
          stage8<X_F32>(x_slot(st, v), x, row * D + c);
          stage8<DY_F32>(dy_slot(st, v), dy, row * D + c);
        }
      }
    }
    cp_async_commit();
  };
  float gm[NV][8], dg[NV][8], db[NV][8], dxs[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = (v * blockDim.x + tid) * 8;
    if (c < D) load8<true>(gamma, c, gm[v]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dg[v][j] = 0.f;
      db[v][j] = 0.f;
      dxs[v][j] = 0.f;
    }
  }
  const float invD = 1.0f / (float)D;
  const int64_t ld8 = D >> 3;
#pragma unroll
  for (int i = 0; i < NSTG - 1; ++i) issue((int64_t)blockIdx.x + (int64_t)i * gridDim.x, i);
  int st = 0;
  int64_t row = blockIdx.x;
  float nmean = 0.f, nrstd = 0.f;
  if (row < rows) {
    nmean = __ldg(mean_in + row);
    nrstd = __ldg(rstd_in + row);
  }
  for (; row < rows; row += gridDim.x) {
    cp_async_wait<NSTG - 2>();
    float xh[NV][8], dyg[NV][8];
    const float mean = nmean, rstd = nrstd;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = (v * blockDim.x + tid) * 8;
      if (c < D) {
        unstage8<X_F32>(x_slot(st, v), xh[v]);
        unstage8<DY_F32>(dy_slot(st, v), dyg[v]);
      }
    }
    issue(row + (int64_t)(NSTG - 1) * gridDim.x, st == 0 ? NSTG - 1 : st - 1);  // slot consumed one iteration ago
    st = (st + 1 == NSTG) ? 0 : st + 1;
    const int64_t nrow = row + gridDim.x;
    if (nrow < rows) {
      nmean = __ldg(mean_in + nrow);
      nrstd = __ldg(rstd_in + nrow);
    }
    float red[2] = {0.f, 0.f};
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = (v * blockDim.x + tid) * 8;
      if (c < D) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xhat = (xh[v][j] - mean) * rstd;
          const float d = dyg[v][j];
          xh[v][j] = xhat;
          dg[v][j] += d * xhat;
          db[v][j] += d;
          const float dgm = d * gm[v][j];
          dyg[v][j] = dgm;
          red[0] += dgm;
          red[1] += dgm * xhat;
        }
      }
    }
    block_sum_n<2>(red, buf, phase);
    const float s1 = red[0] * invD, s2 = red[1] * invD;
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
        if (want_dxsum) {
          // column sums of the gradient that enters the preceding Linear (its bias gradient), taken on the
          // bf16-rounded values so they equal a column sum of the stored tensor
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dxs[v][j] += (DX_F32 && dx_drop == nullptr) ? o[j] : __bfloat162float(__float2bfloat16(o[j]));
        }
      }
    }
  }
  cp_async_wait<0>();
  float* pg = partial + (int64_t)blockIdx.x * 3 * D;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = (v * blockDim.x + tid) * 8;
    if (c < D) {
      store8<true>(pg, c, dg[v]);
      store8<true>(pg + D, c, db[v]);
      if (want_dxsum) store8<true>(pg + 2 * D, c, dxs[v]);
    }
  }
}

// out[k][c] = sum_b partial[b][k][c], k in {0,1,(2)}: block (32 columns, 8 partial lanes), coalesced 128-byte reads
__global__ void __launch_bounds__(256)
ln_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int D, int nk, float* __restrict__ dgamma,
                       float* __restrict__ dbeta, float* __restrict__ dxsum) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;  // column index in [0, nk*D)
  float s = 0.f;
  if (c < nk * D) {
    const int k = c / D, col = c - k * D;
    for (int b = ty; b < nblocks; b += 8) s += partial[((int64_t)b * 3 + k) * D + col];
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < nk * D) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    const int k = c / D, col = c - k * D;
    if (k == 0) dgamma[col] = t;
    else if (k == 1) dbeta[col] = t;
    else dxsum[col] = t;
  }
}

static int threads_for(int64_t D, int& nv) {
  // threads*8*nv >= D, threads a multiple of 32, <= 256
  nv = (int)((D + 2047) / 2048);
  if (nv == 3) nv = 4;
  int threads = (int)((D / 8 + nv - 1) / nv);
  threads = ((threads + 31) / 32) * 32;
  return threads;
}

// persistent grids are sized to exactly one resident wave (SMs x occupancy): a partial second wave would run at a
// fraction of the occupancy.  The backward partial buffer is sized for the upper bound (8 CTAs / SM).
static int64_t bwd_grid_max(int64_t rows) {
  int64_t g = (int64_t)num_sms() * 8;
  if (g > rows) g = rows;
  if (g < 1) g = 1;
  return g;
}
template <typename K>
static int64_t resident_grid(K kern, int threads, size_t smem, int64_t work, int cap_per_sm) {
  int occ = 0;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess || occ < 1) occ = 1;
  if (occ > cap_per_sm) occ = cap_per_sm;
  int64_t g = (int64_t)num_sms() * occ;
  if (g > work) g = work;
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
  const int64_t iters = (rows + ln::RPI - 1) / ln::RPI;
#define LSTC_LN_FWD_K(NV, A, B)                                                                       \
  do {                                                                                                \
    auto kern = ln::ln_fwd_kernel<NV, A, B>;                                                          \
    const size_t smem = (size_t)ln::NSTG * ln::RPI * NV * threads * (A ? 32 : 16);                    \
    const int64_t grid = ln::resident_grid(kern, threads, smem, iters, 8);                            \
    kern<<<(unsigned)grid, threads, smem, stream>>>(x, gamma, beta, y, mean, rstd, rows, (int)D, eps); \
  } while (0)
#define LSTC_LN_FWD(NV)                                         \
  do {                                                          \
    if (x_is_f32 && y_is_f32) LSTC_LN_FWD_K(NV, true, true);    \
    else if (x_is_f32) LSTC_LN_FWD_K(NV, true, false);          \
    else if (y_is_f32) LSTC_LN_FWD_K(NV, false, true);          \
    else LSTC_LN_FWD_K(NV, false, false);                       \
  } while (0)
  if (nv == 1) LSTC_LN_FWD(1);
  else if (nv == 2) LSTC_LN_FWD(2);
  else LSTC_LN_FWD(4);
#undef LSTC_LN_FWD
#undef LSTC_LN_FWD_K
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

extern "C" int64_t lstc_layernorm_bwd_workspace(int64_t rows, int64_t D) {
  return ln::bwd_grid_max(rows) * 3 * D * (int64_t)sizeof(float);
}

extern "C" int lstc_layernorm_bwd(const void* dy, int dy_is_f32, const void* x, int x_is_f32, const float* gamma,
                                  const float* mean, const float* rstd, void* dx, int dx_is_f32, void* dx_drop,
                                  float drop_p, uint64_t seed, uint64_t offset, float* dgamma, float* dbeta,
                                  float* dxsum, void* workspace, int64_t rows, int64_t D, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LSTC_CHECK_ARG(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && workspace,
                 "lstc_layernorm_bwd: null pointer");
  LSTC_CHECK_ARG(D > 0 && D % 8 == 0 && D <= 8192, "lstc_layernorm_bwd: D=%lld must be a multiple of 8, <= 8192",
                 (long long)D);
  LSTC_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "lstc_layernorm_bwd: drop_p out of range");
  if (rows == 0) {
    LSTC_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, D * sizeof(float), stream));
    LSTC_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, D * sizeof(float), stream));
    if (dxsum) LSTC_CHECK_CUDA(cudaMemsetAsync(dxsum, 0, D * sizeof(float), stream));
    return LSTC_OK;
  }
  int nv;
  const int threads = ln::threads_for(D, nv);
  int64_t grid = 1;
  __nv_bfloat16* dd = (drop_p > 0.f) ? (__nv_bfloat16*)dx_drop : nullptr;
  const float dscale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const uint32_t thr = dropout_threshold16(drop_p);
  float* partial = (float*)workspace;
#define LSTC_LN_BWD_K(NV, A, B, C)                                                                             \
  do {                                                                                                         \
    auto kern = ln::ln_bwd_kernel<NV, A, B, C>;                                                                \
    const size_t smem = (size_t)ln::NSTG * NV * threads * ((B ? 32 : 16) + (A ? 32 : 16));                     \
    grid = ln::resident_grid(kern, threads, smem, rows, 8);                                                    \
    kern<<<(unsigned)grid, threads, smem, stream>>>(dy, x, gamma, mean, rstd, dx, dd, dscale, thr, seed, offset, \
                                                 partial, dxsum != nullptr ? 1 : 0, rows, (int)D);             \
  } while (0)
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
  const int nk = dxsum != nullptr ? 3 : 2;
  ln::ln_bwd_finalize_kernel<<<(unsigned)((nk * D + 31) / 32), 256, 0, stream>>>(partial, (int)grid, (int)D, nk, dgamma,
                                                                               dbeta, dxsum);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

LSTC_DEFINE_RNG_STEP_SETTER(set_rng_step_layernorm)
