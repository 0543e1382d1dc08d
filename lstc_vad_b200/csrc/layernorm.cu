// LayerNorm forward / backward over the last dimension, fp32 statistics, bf16 or fp32 I/O.
// Replaces nn.LayerNorm(d_model, eps=1e-6) at models/Encoder.py:31,48-49,
// models/MultiHeadAttention.py:47,125-126 and models/FFN.py:10,20-21 (and its autograd backward).
//
// One CTA walks over groups of RPI rows (grid-stride); thread t owns the same 8*NV columns of every row, so
//   - each row is one fully coalesced 16-byte-per-thread read and write,
//   - gamma / beta (forward) and the per-column dgamma / dbeta / dx sums (backward) stay in registers for the whole
//     kernel; the backward flushes them once per CTA into a partial buffer that a second tiny kernel reduces
//     (deterministic, no atomics).
// Both kernels are bound by HBM only if they stay lean: the row statistics of RPI rows are reduced TOGETHER
// (halving butterfly in the warp, one barrier, a 32-lane second stage), which costs ~1.7 instructions per element
// instead of the ~10 of per-row shuffle trees plus an all-warps smem read.
#include <mutex>
#include <vector>

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
__device__ __forceinline__ void store8_f32(float* p, const float (&f)[8]) {
  __stcs(reinterpret_cast<float4*>(p), make_float4(f[0], f[1], f[2], f[3]));
  __stcs(reinterpret_cast<float4*>(p + 4), make_float4(f[4], f[5], f[6], f[7]));
}
__device__ __forceinline__ void store8_bf16(__nv_bfloat16* p, const uint4& v) { __stcs(reinterpret_cast<uint4*>(p), v); }

// CTA-wide sums of N values per thread (N = 2, 4 or 8; one value per row statistic).  `scratch` is
// [2][MAX_WARPS][N] floats, zero-initialised once (warps that do not exist contribute zeros); `phase` alternates so
// that one barrier per call suffices.
//   stage 1: halving butterfly -- at each step a lane keeps half of its values and hands the other half to its
//            partner, so N values cost N-1 + (5 - log2 N) shuffles instead of 5 N; lane l ends with the warp total
//            of value l >> (5 - log2 N);
//   stage 2: after the barrier lane l fetches the partial of warp (l & 7) for value (l >> 3) [+4], three xor
//            steps add the 8 warps, and N broadcasts hand every total to every lane.
template <int N>
__device__ __forceinline__ void block_sum(float (&v)[N], float* scratch, int& phase, int lane, int warp) {
  static_assert(N == 2 || N == 4 || N == 8, "block_sum: N must be 2, 4 or 8");
  constexpr int LG = (N == 8) ? 3 : (N == 4) ? 2 : 1;
#pragma unroll
  for (int s = 0; s < LG; ++s) {
    const int n = N >> s, mask = 16 >> s;
    const bool up = (lane & mask) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = up ? v[i] : v[i + n / 2];
      const float keep = up ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
    }
  }
#pragma unroll
  for (int s = LG; s < 5; ++s) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16 >> s);
  float* S = scratch + phase * (MAX_WARPS * N);
  if ((lane & ((32 >> LG) - 1)) == 0) S[warp * N + (lane >> (5 - LG))] = v[0];
  __syncthreads();
  float a = S[(lane & 7) * N + ((lane >> 3) & (N - 1))];
  float b = 0.f;
  if (N == 8) b = S[(lane & 7) * N + (lane >> 3) + 4];
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    if (N == 8) b += __shfl_xor_sync(0xffffffffu, b, o);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = __shfl_sync(0xffffffffu, (i < 4) ? a : b, (i & 3) * 8);
  phase ^= 1;
}

// Every thread stages ITS OWN 16/32-byte chunks of the next NSTG-1 iterations in shared memory with cp.async and is
// the only reader of those bytes, so the ring needs no barrier: cp.async.wait_group on the thread's own groups is
// enough.  This keeps ~3 iterations per CTA in flight without holding them in registers.
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

template <bool IS_F32>
__device__ __forceinline__ void stage8b(uint32_t dst, const char* src) {
  cp_async16(dst, src);
  if (IS_F32) cp_async16(dst + 16, src + 16);
}
__device__ __forceinline__ float rsqrt_approx(float v) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
template <bool B>
struct BoolTag {
  static constexpr bool value = B;
};

// Forward: two passes over the registers (mean, then the sum of squared deviations -- the same order of operations
// as the reference's nn.LayerNorm, no E[x^2] - mean^2 cancellation), two block_sum calls per RPI rows.  Global
// pointers, ring offsets and the group counter advance incrementally; only the ragged last group takes the
// row-predicated store path.
template <int NV, int RPI, int NSTG, bool X_F32, bool Y_F32>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const void* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, void* __restrict__ y,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                     int rows, int D, float eps) {
  __shared__ float scratch[2 * MAX_WARPS * RPI];
  extern __shared__ __align__(16) uint8_t ring_raw[];
  constexpr int XB = X_F32 ? 32 : 16, XE = X_F32 ? 4 : 2, YE = Y_F32 ? 4 : 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
  for (int i = tid; i < 2 * MAX_WARPS * RPI; i += nthr) scratch[i] = 0.f;
  __syncthreads();
  int phase = 0;
  bool act[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) act[v] = (v * nthr + tid) * 8 < D;
  const int vcols = nthr * 8;  // column distance between the chunks of one thread
  const uint32_t chunk_stride = nthr * XB, stage_bytes = RPI * NV * chunk_stride, ring_bytes = NSTG * stage_bytes;
  const uint32_t ring = smem_addr(ring_raw) + tid * XB;
  const int ngroups = (rows + RPI - 1) / RPI;
  const int64_t first = ((int64_t)blockIdx.x * RPI) * D + tid * 8;  // element offset of this thread's first chunk
  const char* xg = reinterpret_cast<const char*>(x) + first * XE;
  const int64_t xg_step = (int64_t)gridDim.x * RPI * D * XE;
  int g_issue = blockIdx.x;
  uint32_t st_issue = 0;
  auto issue = [&]() {
    if (g_issue < ngroups) {
      const int nvalid = rows - g_issue * RPI;
#pragma unroll
      for (int r = 0; r < RPI; ++r) {
        if (r < nvalid) {
#pragma unroll
          for (int v = 0; v < NV; ++v)
            if (act[v])
              stage8b<X_F32>(ring + st_issue + (r * NV + v) * chunk_stride, xg + ((int64_t)r * D + v * vcols) * XE);
        }
      }
      xg += xg_step;
    }
    g_issue += gridDim.x;
    st_issue = (st_issue + stage_bytes == ring_bytes) ? 0u : st_issue + stage_bytes;
    cp_async_commit();
  };
  float gm[NV][8], bt[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    if (act[v]) {
      load8<true>(gamma, (v * nthr + tid) * 8, gm[v]);
      load8<true>(beta, (v * nthr + tid) * 8, bt[v]);
    }
  }
  const float invD = 1.0f / (float)D;
#pragma unroll
  for (int i = 0; i < NSTG - 1; ++i) issue();
  uint32_t st_read = 0;
  char* yg = reinterpret_cast<char*>(y) + first * YE;
  const int64_t yg_step = (int64_t)gridDim.x * RPI * D * YE;
  for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
    cp_async_wait<NSTG - 2>();
    float xv[RPI][NV][8];
    float s[RPI];
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      s[r] = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (act[v]) {
          unstage8<X_F32>(ring + st_read + (r * NV + v) * chunk_stride, xv[r][v]);
#pragma unroll
          for (int j = 0; j < 8; ++j) s[r] += xv[r][v][j];
        }
      }
    }
    st_read = (st_read + stage_bytes == ring_bytes) ? 0u : st_read + stage_bytes;
    issue();  // refills the slot consumed one iteration ago
    block_sum<RPI>(s, scratch, phase, lane, warp);
    float q[RPI];
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      s[r] *= invD;  // mean
      q[r] = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (act[v]) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            xv[r][v][j] -= s[r];
            q[r] = fmaf(xv[r][v][j], xv[r][v][j], q[r]);
          }
        }
      }
    }
    block_sum<RPI>(q, scratch, phase, lane, warp);
#pragma unroll
    for (int r = 0; r < RPI; ++r) q[r] = rsqrt_approx(fmaf(q[r], invD, eps));  // rstd
    const int nvalid = rows - g * RPI;
    auto emit = [&](auto full) {
#pragma unroll
      for (int r = 0; r < RPI; ++r) {
        if (decltype(full)::value || r < nvalid) {
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            if (act[v]) {
              float o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = fmaf(xv[r][v][j] * q[r], gm[v][j], bt[v][j]);
              char* dst = yg + ((int64_t)r * D + v * vcols) * YE;
              if (Y_F32) store8_f32(reinterpret_cast<float*>(dst), o);
              else store8_bf16(reinterpret_cast<__nv_bfloat16*>(dst), pack8(o));
            }
          }
          if (tid == 0) {
            mean_out[(int64_t)g * RPI + r] = s[r];
            rstd_out[(int64_t)g * RPI + r] = q[r];
          }
        }
      }
    };
    if (nvalid >= RPI) emit(BoolTag<true>());
    else emit(BoolTag<false>());
    yg += yg_step;
  }
  cp_async_wait<0>();
}

// Backward: dx = rstd * (dy*gamma - mean_D(dy*gamma) - xhat * mean_D(dy*gamma*xhat)); one block_sum of 2*RPI values
// per RPI rows.  Optionally also writes dropout(dx) (the gradient entering the sub-layer whose output was dropped
// before the residual add) and accumulates the column sums of the tensor the preceding Linear receives.
template <int NV, int RPI, int NSTG, bool DY_F32, bool X_F32, bool DX_F32>
__global__ void __launch_bounds__(256, (NV == 1 && !X_F32) ? 2 : 1)
ln_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ mean_in, const float* __restrict__ rstd_in, void* __restrict__ dx,
              __nv_bfloat16* __restrict__ dx_drop, float drop_scale, uint32_t drop_thr16, uint64_t seed,
              uint64_t offset, float* __restrict__ partial /*[grid][3][D]*/, int want_dxsum, int rows, int D) {
  offset += rng_step();
  __shared__ float scratch[2 * MAX_WARPS * 2 * RPI];
  extern __shared__ __align__(16) uint8_t ring_raw[];
  constexpr int XB = X_F32 ? 32 : 16, YB = DY_F32 ? 32 : 16;
  constexpr int XE = X_F32 ? 4 : 2, YE = DY_F32 ? 4 : 2, OE = DX_F32 ? 4 : 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
  for (int i = tid; i < 2 * MAX_WARPS * 2 * RPI; i += nthr) scratch[i] = 0.f;
  __syncthreads();
  int phase = 0;
  bool act[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) act[v] = (v * nthr + tid) * 8 < D;
  const int vcols = nthr * 8;
  // stage layout: [RPI][NV] x chunks, then [RPI][NV] dy chunks
  const uint32_t xstride = nthr * XB, ystride = nthr * YB;
  const uint32_t stage_bytes = RPI * NV * (xstride + ystride), ring_bytes = NSTG * stage_bytes;
  const uint32_t xring = smem_addr(ring_raw) + tid * XB;
  const uint32_t yring = smem_addr(ring_raw) + RPI * NV * xstride + tid * YB;
  const int ngroups = (rows + RPI - 1) / RPI;
  const int64_t first = ((int64_t)blockIdx.x * RPI) * D + tid * 8;
  const int64_t gstep = (int64_t)gridDim.x * RPI * D;  // elements between consecutive groups of this CTA
  const char* xg = reinterpret_cast<const char*>(x) + first * XE;
  const char* yg = reinterpret_cast<const char*>(dy) + first * YE;
  int g_issue = blockIdx.x;
  uint32_t st_issue = 0;
  auto issue = [&]() {
    if (g_issue < ngroups) {
      const int nvalid = rows - g_issue * RPI;
#pragma unroll
      for (int r = 0; r < RPI; ++r) {
        if (r < nvalid) {
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            if (act[v]) {
              const int64_t e = (int64_t)r * D + v * vcols;
              stage8b<X_F32>(xring + st_issue + (r * NV + v) * xstride, xg + e * XE);
              stage8b<DY_F32>(yring + st_issue + (r * NV + v) * ystride, yg + e * YE);
            }
          }
        }
      }
      xg += gstep * XE;
      yg += gstep * YE;
    }
    g_issue += gridDim.x;
    st_issue = (st_issue + stage_bytes == ring_bytes) ? 0u : st_issue + stage_bytes;
    cp_async_commit();
  };
  float gm[NV][8], dg[NV][8], db[NV][8], dxs[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    if (act[v]) load8<true>(gamma, (v * nthr + tid) * 8, gm[v]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dg[v][j] = 0.f;
      db[v][j] = 0.f;
      dxs[v][j] = 0.f;
    }
  }
  const float invD = 1.0f / (float)D;
  const PhiloxKeys keys = philox_keys(seed);
  const int ld8 = D >> 3;
  uint64_t idx8 = (uint64_t)blockIdx.x * RPI * ld8 + tid;  // dropout counter of this thread's first chunk
  const uint64_t idx8_step = (uint64_t)gridDim.x * RPI * ld8;
#pragma unroll
  for (int i = 0; i < NSTG - 1; ++i) issue();
  uint32_t st_read = 0;
  char* og = reinterpret_cast<char*>(dx) + first * OE;
  __nv_bfloat16* odg = dx_drop != nullptr ? dx_drop + first : nullptr;
  float nrs[RPI], nnb[RPI];  // next group's rstd and -mean*rstd (0 for rows past the end)
  auto fetch_stats = [&](int g) {
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      nrs[r] = 0.f;
      nnb[r] = 0.f;
      if (g < ngroups && g * RPI + r < rows) {
        nrs[r] = __ldg(rstd_in + (int64_t)g * RPI + r);
        nnb[r] = -__ldg(mean_in + (int64_t)g * RPI + r) * nrs[r];
      }
    }
  };
  fetch_stats(blockIdx.x);
  for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
    cp_async_wait<NSTG - 2>();
    float xh[RPI][NV][8], dgm[RPI][NV][8];
    float rs[RPI], nb[RPI];
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      rs[r] = nrs[r];
      nb[r] = nnb[r];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (act[v]) {
          unstage8<X_F32>(xring + st_read + (r * NV + v) * xstride, xh[r][v]);
          unstage8<DY_F32>(yring + st_read + (r * NV + v) * ystride, dgm[r][v]);
        }
      }
    }
    const int nvalid = rows - g * RPI;
    if (nvalid < RPI) {  // ragged last group: rows past the end contribute exact zeros
#pragma unroll
      for (int r = 0; r < RPI; ++r) {
        if (r >= nvalid) {
#pragma unroll
          for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              xh[r][v][j] = 0.f;
              dgm[r][v][j] = 0.f;
            }
        }
      }
    }
    st_read = (st_read + stage_bytes == ring_bytes) ? 0u : st_read + stage_bytes;
    issue();  // refills the slot consumed one iteration ago
    fetch_stats(g + gridDim.x);
    float red[2 * RPI];
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      red[2 * r] = 0.f;
      red[2 * r + 1] = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (act[v]) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xhat = fmaf(xh[r][v][j], rs[r], nb[r]);
            const float d = dgm[r][v][j];
            xh[r][v][j] = xhat;
            dg[v][j] = fmaf(d, xhat, dg[v][j]);
            db[v][j] += d;
            const float t = d * gm[v][j];
            dgm[r][v][j] = t;
            red[2 * r] += t;
            red[2 * r + 1] = fmaf(t, xhat, red[2 * r + 1]);
          }
        }
      }
    }
    block_sum<2 * RPI>(red, scratch, phase, lane, warp);
    auto emit = [&](auto full) {
#pragma unroll
      for (int r = 0; r < RPI; ++r) {
        if (decltype(full)::value || r < nvalid) {
          const float b1 = -red[2 * r] * invD * rs[r], b2 = -red[2 * r + 1] * invD * rs[r];
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            if (act[v]) {
              float o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = fmaf(xh[r][v][j], b2, fmaf(dgm[r][v][j], rs[r], b1));
              const int64_t e = (int64_t)r * D + v * vcols;
              uint4 pk = make_uint4(0u, 0u, 0u, 0u);
              if (DX_F32) store8_f32(reinterpret_cast<float*>(og + e * OE), o);
              else {
                pk = pack8(o);
                store8_bf16(reinterpret_cast<__nv_bfloat16*>(og + e * OE), pk);
              }
              if (dx_drop != nullptr) {
                uint32_t rnd[4];
                philox4x32_keyed(keys, offset, idx8 + (uint64_t)(r * ld8 + v * nthr), rnd);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = dropout_kept(rnd, j, drop_thr16) ? o[j] * drop_scale : 0.f;
                pk = pack8(o);
                store8_bf16(odg + e, pk);
              }
              if (want_dxsum) {
                // column sums of the gradient that enters the preceding Linear (its bias gradient), taken on the
                // bf16-rounded values so they equal a column sum of the stored tensor
                if (!DX_F32 || dx_drop != nullptr) unpack8(pk, o);
#pragma unroll
                for (int j = 0; j < 8; ++j) dxs[v][j] += o[j];
              }
            }
          }
        }
      }
    };
    if (nvalid >= RPI) emit(BoolTag<true>());
    else emit(BoolTag<false>());
    og += gstep * OE;
    if (dx_drop != nullptr) odg += gstep;
    idx8 += idx8_step;
  }
  cp_async_wait<0>();
  float* pg = partial + (int64_t)blockIdx.x * 3 * D;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    if (act[v]) {
      const int c = (v * nthr + tid) * 8;
      store8_f32(pg + c, dg[v]);
      store8_f32(pg + D + c, db[v]);
      if (want_dxsum) store8_f32(pg + 2 * D + c, dxs[v]);
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
// resident CTAs per SM of (kernel, block size, dynamic smem) on the current device; the function attribute and the
// occupancy query are issued once per combination and device, not per launch
struct OccEntry {
  const void* kern;
  int threads, dev, occ;
  size_t smem;
};
static int cached_occupancy(const void* kern, int threads, size_t smem) {
  static std::mutex mu;
  static std::vector<OccEntry> table;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  std::lock_guard<std::mutex> lock(mu);
  for (const OccEntry& e : table)
    if (e.kern == kern && e.threads == threads && e.smem == smem && e.dev == dev) return e.occ;
  // the opt-in limit is one attribute per kernel: only ever raise it (the same kernel runs with several block sizes)
  size_t limit = smem;
  bool raise = true;
  for (const OccEntry& e : table)
    if (e.kern == kern && e.dev == dev && e.smem >= limit) raise = false;
  if (raise) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit);
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess || occ < 1) occ = 1;
  table.push_back(OccEntry{kern, threads, dev, occ, smem});
  return occ;
}
template <typename K>
static int64_t resident_grid(K kern, int threads, size_t smem, int64_t work, int cap_per_sm) {
  int occ = cached_occupancy(reinterpret_cast<const void*>(kern), threads, smem);
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
  LSTC_CHECK_ARG(rows >= 0 && rows < (int64_t)2000000000, "lstc_layernorm_fwd: rows=%lld out of range", (long long)rows);
  LSTC_CHECK_ARG(LSTC_ALIGNED16(x) && LSTC_ALIGNED16(y) && LSTC_ALIGNED16(gamma) && LSTC_ALIGNED16(beta),
                 "lstc_layernorm_fwd: x, y, gamma, beta must be 16-byte aligned");
  if (rows == 0) return LSTC_OK;
  int nv;
  const int threads = ln::threads_for(D, nv);
  // rows per iteration / ring depth: sized so that the row registers (RPI*NV*8 floats) and the ring
  // (NS*RPI*NV*threads*{16,32} B) stay within one CTA's budget
#define LSTC_LN_FWD_K(NV, RPI, NS, A, B)                                                              \
  do {                                                                                                \
    auto kern = ln::ln_fwd_kernel<NV, RPI, NS, A, B>;                                                 \
    const size_t smem = (size_t)NS * RPI * NV * threads * (A ? 32 : 16);                              \
    const int64_t iters = (rows + RPI - 1) / RPI;                                                     \
    const int64_t grid = ln::resident_grid(kern, threads, smem, iters, 8);                            \
    kern<<<(unsigned)grid, threads, smem, stream>>>(x, gamma, beta, y, mean, rstd, (int)rows, (int)D, eps); \
  } while (0)
#define LSTC_LN_FWD(NV, RPI_BF, NS_BF, RPI_F32, NS_F32)                       \
  do {                                                                        \
    if (x_is_f32 && y_is_f32) LSTC_LN_FWD_K(NV, RPI_F32, NS_F32, true, true); \
    else if (x_is_f32) LSTC_LN_FWD_K(NV, RPI_F32, NS_F32, true, false);       \
    else if (y_is_f32) LSTC_LN_FWD_K(NV, RPI_BF, NS_BF, false, true);         \
    else LSTC_LN_FWD_K(NV, RPI_BF, NS_BF, false, false);                      \
  } while (0)
  if (nv == 1) LSTC_LN_FWD(1, 4, 6, 2, 4);  // measured on B200 at 62720 x 2048: 4 rows x 6 stages beats 4x4, 2x4, 2x8
  else if (nv == 2) LSTC_LN_FWD(2, 2, 3, 2, 3);
  else LSTC_LN_FWD(4, 2, 2, 2, 2);
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
  LSTC_CHECK_ARG(rows >= 0 && rows < (int64_t)2000000000, "lstc_layernorm_bwd: rows=%lld out of range", (long long)rows);
  LSTC_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "lstc_layernorm_bwd: drop_p out of range");
  LSTC_CHECK_ARG(LSTC_ALIGNED16(dy) && LSTC_ALIGNED16(x) && LSTC_ALIGNED16(dx) && LSTC_ALIGNED16(dx_drop) &&
                     LSTC_ALIGNED16(gamma) && LSTC_ALIGNED16(workspace),
                 "lstc_layernorm_bwd: dy, x, dx, dx_drop, gamma, workspace must be 16-byte aligned");
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
#define LSTC_LN_BWD_K(NV, RPI, NS, A, B, C)                                                                    \
  do {                                                                                                         \
    auto kern = ln::ln_bwd_kernel<NV, RPI, NS, A, B, C>;                                                       \
    const size_t smem = (size_t)NS * RPI * NV * threads * ((B ? 32 : 16) + (A ? 32 : 16));                     \
    grid = ln::resident_grid(kern, threads, smem, (rows + RPI - 1) / RPI, 8);                                  \
    kern<<<(unsigned)grid, threads, smem, stream>>>(dy, x, gamma, mean, rstd, dx, dd, dscale, thr, seed, offset, \
                                                 partial, dxsum != nullptr ? 1 : 0, (int)rows, (int)D);             \
  } while (0)
#define LSTC_LN_BWD(NV, RPI_BF, NS_BF, RPI_MIX, NS_MIX, RPI_F32, NS_F32)                                       \
  do {                                                                                                         \
    if (dy_is_f32 && x_is_f32 && dx_is_f32) LSTC_LN_BWD_K(NV, RPI_F32, NS_F32, true, true, true);              \
    else if (dy_is_f32 && !x_is_f32 && !dx_is_f32) LSTC_LN_BWD_K(NV, RPI_MIX, NS_MIX, true, false, false);     \
    else if (!dy_is_f32 && x_is_f32 && dx_is_f32) LSTC_LN_BWD_K(NV, RPI_F32, NS_F32, false, true, true);       \
    else if (!dy_is_f32 && !x_is_f32 && !dx_is_f32) LSTC_LN_BWD_K(NV, RPI_BF, NS_BF, false, false, false);     \
    else {                                                                                                     \
      set_last_error("lstc_layernorm_bwd: unsupported dtype combination (dx must match x)");                   \
      return LSTC_ERR_UNSUPPORTED;                                                                             \
    }                                                                                                          \
  } while (0)
  // measured on B200 at 62720 x 2048: 2 rows x 4 stages (120 registers, 2 CTAs / SM) beats 4x3, 4x2, 1x4 and 3 CTAs / SM
  if (nv == 1) LSTC_LN_BWD(1, 2, 4, 2, 4, 2, 3);
  else if (nv == 2) LSTC_LN_BWD(2, 1, 4, 1, 4, 1, 4);
  else LSTC_LN_BWD(4, 1, 2, 1, 2, 1, 2);
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
