// Bandwidth-bound glue kernels: dtype casts, CLS-token prepend (+ absolute position encoding),
// column sums (bias gradients), dropout apply / mask dump, relative-position bias gather/scatter,
// device-scalar scaling and the fused Adagrad update.
#include "common.cuh"
#include "../../include/lstc_vad_b200.h"

namespace lstc {
namespace ew {

static inline unsigned grid_for(int64_t work_items, int threads, int per_sm = 8) {
  int64_t g = (work_items + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms() * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

// ---------------------------------------------------------------- casts
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  const int64_t n8 = n >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    reinterpret_cast<uint4*>(dst)[i] = pack8(f);
  }
  for (int64_t i = (n8 << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __float2bfloat16(src[i]);
}

// dst[c, r] = bf16(src[r, c]) : transposed bf16 copy of an fp32 weight (the K-major B operand of the input-gradient
// products).  64 x 64 tiles through shared memory: coalesced fp32 reads along c, coalesced bf16 writes along r.
__global__ void cast_transpose_f32_bf16_kernel(const float* __restrict__ src, int64_t ld_src, __nv_bfloat16* __restrict__ dst,
                                               int64_t ld_dst, int64_t rows, int64_t cols) {
  __shared__ float tile[64][65];
  const int64_t r0 = (int64_t)blockIdx.y * 64, c0 = (int64_t)blockIdx.x * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 256 threads: 64 x 4
#pragma unroll
  for (int i = 0; i < 64; i += 4) {
    const int64_t r = r0 + ty + i, c = c0 + tx;
    tile[ty + i][tx] = (r < rows && c < cols) ? src[r * ld_src + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 64; i += 4) {
    const int64_t c = c0 + ty + i, r = r0 + tx;
    if (c < cols && r < rows) dst[c * ld_dst + r] = __float2bfloat16(tile[tx][ty + i]);
  }
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int64_t n) {
  const int64_t n8 = n >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + i);
    float f[8];
    unpack8(v, f);
    __stcs(reinterpret_cast<float4*>(dst) + 2 * i, make_float4(f[0], f[1], f[2], f[3]));
    __stcs(reinterpret_cast<float4*>(dst) + 2 * i + 1, make_float4(f[4], f[5], f[6], f[7]));
  }
  for (int64_t i = (n8 << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __bfloat162float(src[i]);
}

// ---------------------------------------------------------------- CLS prepend
// grid = (W, ceil(D/8/blockDim)); thread owns 8 columns of one window and walks over its tokens.
template <bool X_F32>
__global__ void cls_prepend_fwd_kernel(const void* __restrict__ x, const float* __restrict__ cls,
                                       const float* __restrict__ pos, float drop_scale, uint32_t thr16,
                                       uint64_t seed, uint64_t offset, int use_drop,
                                       __nv_bfloat16* __restrict__ out, int64_t L0, int64_t D) {
  offset += rng_step();
  const int64_t w = blockIdx.x;
  const int64_t c = ((int64_t)blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (c >= D) return;
  const int64_t L = L0 + 1;
  const int64_t d8 = D >> 3;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  auto emit = [&](int64_t l, float (&v)[8]) {
    if (pos != nullptr) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(pos + l * D + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(pos + l * D + c + 4));
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    if (use_drop) {
      const uint32_t keep = dropout_keep8(seed, offset, (uint64_t)((w * L + l) * d8 + (c >> 3)), thr16);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = ((keep >> j) & 1u) ? v[j] * drop_scale : 0.f;
    }
    *reinterpret_cast<uint4*>(out + (w * L + l) * D + c) = pack8(v);
  };
  for (int64_t l = 0; l < L0; ++l) {
    float v[8];
    if (X_F32) {
      const float* px = reinterpret_cast<const float*>(x) + (w * L0 + l) * D + c;
      const float4 a = __ldg(reinterpret_cast<const float4*>(px));
      const float4 b = __ldg(reinterpret_cast<const float4*>(px + 4));
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
      unpack8(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(x) + (w * L0 + l) * D + c)), v);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
    emit(l + 1, v);
  }
  float v[8];
  if (cls != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldg(cls + c + j);
  } else {
    const float inv = 1.0f / (float)L0;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = acc[j] * inv;
  }
  emit(0, v);
}

// g [W, L, D] bf16 -> dx [W, L0, D] fp32 ; optional atomics into dcls [D] / dpos [L, D] (rarely used paths)
__global__ void cls_prepend_bwd_kernel(const __nv_bfloat16* __restrict__ g, int cls_learned, float drop_scale,
                                       uint32_t thr16, uint64_t seed, uint64_t offset, int use_drop,
                                       float* __restrict__ dx, float* __restrict__ dcls, float* __restrict__ dpos,
                                       int64_t L0, int64_t D) {
  offset += rng_step();
  const int64_t w = blockIdx.x;
  const int64_t c = ((int64_t)blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (c >= D) return;
  const int64_t L = L0 + 1;
  const int64_t d8 = D >> 3;
  auto fetch = [&](int64_t l, float (&v)[8]) {
    unpack8(__ldg(reinterpret_cast<const uint4*>(g + (w * L + l) * D + c)), v);
    if (use_drop) {
      const uint32_t keep = dropout_keep8(seed, offset, (uint64_t)((w * L + l) * d8 + (c >> 3)), thr16);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = ((keep >> j) & 1u) ? v[j] * drop_scale : 0.f;
    }
    if (dpos != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(dpos + l * D + c + j, v[j]);
    }
  };
  float g0[8];
  fetch(0, g0);
  if (cls_learned) {
    if (dcls != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(dcls + c + j, g0[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) g0[j] = 0.f;
  } else {
    const float inv = 1.0f / (float)L0;
#pragma unroll
    for (int j = 0; j < 8; ++j) g0[j] *= inv;
  }
  for (int64_t l = 0; l < L0; ++l) {
    float v[8];
    fetch(l + 1, v);
    if (dx != nullptr) {
      float* p = dx + (w * L0 + l) * D + c;
      *reinterpret_cast<float4*>(p) = make_float4(v[0] + g0[0], v[1] + g0[1], v[2] + g0[2], v[3] + g0[3]);
      *reinterpret_cast<float4*>(p + 4) = make_float4(v[4] + g0[4], v[5] + g0[5], v[6] + g0[6], v[7] + g0[7]);
    }
  }
}

// ---------------------------------------------------------------- column sum
// block (32, 8): threadIdx.x -> 8 columns each (256 columns per block), threadIdx.y -> row lane.
__global__ void colsum_stage1_kernel(const __nv_bfloat16* __restrict__ x, int64_t rows, int64_t cols, int64_t ld,
                                     float* __restrict__ partial /*[gridDim.y][cols]*/) {
  __shared__ float red[8][32][9];
  const int64_t c = ((int64_t)blockIdx.x * 32 + threadIdx.x) * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < cols) {
    const bool vec = (c + 8 <= cols) && (ld % 8 == 0);
    const int64_t rstep = (int64_t)gridDim.y * 8;
    int64_t r = (int64_t)blockIdx.y * 8 + threadIdx.y;
    if (vec) {
      // 4 independent 16-byte loads in flight per thread
      for (; r + 3 * rstep < rows; r += 4 * rstep) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(x + (r + u * rstep) * ld + c));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float f[8];
          unpack8(v[u], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
      }
    }
    for (; r < rows; r += rstep) {
      if (vec) {
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + r * ld + c)), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      } else {
        for (int j = 0; j < 8 && c + j < cols; ++j) acc[j] += __bfloat162float(x[r * ld + c + j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.y][threadIdx.x][j] = acc[j];
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) s += red[y][threadIdx.x][j];
      if (c + j < cols) partial[(int64_t)blockIdx.y * cols + c + j] = s;
    }
  }
}
__global__ void colsum_stage2_kernel(const float* __restrict__ partial, int nparts, int64_t cols,
                                     float* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += partial[(int64_t)p * cols + c];
  out[c] = s;
}
static int colsum_parts(int64_t rows) {
  int64_t p = (rows + 63) / 64;
  if (p > 128) p = 128;
  if (p < 1) p = 1;
  return (int)p;
}

// ---------------------------------------------------------------- dropout helpers
__global__ void dropout_apply_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                     int64_t n8, float scale, uint32_t thr16, uint64_t seed, uint64_t offset) {
  offset += rng_step();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), f);
    const uint32_t keep = dropout_keep8(seed, offset, (uint64_t)i, thr16);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = ((keep >> j) & 1u) ? f[j] * scale : 0.f;
    reinterpret_cast<uint4*>(y)[i] = pack8(f);
  }
}
__global__ void dropout_mask_kernel(uint8_t* __restrict__ mask, int64_t rows, int64_t cols, uint32_t thr16,
                                    uint64_t seed, uint64_t offset) {
  offset += rng_step();
  const int64_t ld8 = (cols + 7) >> 3;
  const int64_t total = rows * ld8;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / ld8, c8 = i % ld8;
    const uint32_t keep = dropout_keep8(seed, offset, (uint64_t)i, thr16);
    for (int j = 0; j < 8 && c8 * 8 + j < cols; ++j) mask[r * cols + c8 * 8 + j] = (uint8_t)((keep >> j) & 1u);
  }
}

// ---------------------------------------------------------------- rel-pos bias
__global__ void relbias_gather_kernel(const float* __restrict__ table, const int64_t* __restrict__ index,
                                      int64_t index_ld, int L, int H, int64_t T, float* __restrict__ dense) {
  const int64_t total = (int64_t)H * L * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int h = (int)(i / ((int64_t)L * L));
    const int r = (int)((i / L) % L), c = (int)(i % L);
    float v = 0.f;
    if (r > 0 && c > 0) {
      const int64_t t = index[(int64_t)(r - 1) * index_ld + (c - 1)];
      if (t >= 0 && t < T) v = table[t * H + h];
    }
    dense[i] = v;
  }
}
__global__ void relbias_scatter_kernel(const float* __restrict__ ddense, const int64_t* __restrict__ index,
                                       int64_t index_ld, int L, int H, int64_t T, float* __restrict__ dtable) {
  const int64_t total = (int64_t)H * L * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int h = (int)(i / ((int64_t)L * L));
    const int r = (int)((i / L) % L), c = (int)(i % L);
    if (r > 0 && c > 0) {
      const int64_t t = index[(int64_t)(r - 1) * index_ld + (c - 1)];
      if (t >= 0 && t < T) atomicAdd(dtable + t * H + h, ddense[i]);
    }
  }
}

// ---------------------------------------------------------------- misc
__global__ void scale_by_scalar_kernel(const float* __restrict__ src, const float* __restrict__ scalar,
                                       float* __restrict__ dst, int64_t n) {
  const float s = __ldg(scalar);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i] * s;
}
__global__ void threshold_kernel(const float* __restrict__ s, float thr, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = s[i];
    out[i] = v > thr ? v : 0.f;
  }
}
// sum of squares, block-reduced then one atomic per CTA (gradient-norm clipping)
__global__ void sumsq_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ accum) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    s += v * v;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(accum, s);
  }
}
// coef = min(1, max_norm / (sqrt(sumsq) + 1e-6))   (torch.nn.utils.clip_grad_norm_)
__global__ void clip_coef_kernel(const float* __restrict__ sumsq, float max_norm, float* __restrict__ coef) {
  const float c = max_norm / (sqrtf(sumsq[0]) + 1e-6f);
  coef[0] = c < 1.f ? c : 1.f;
}
// torch.optim.Adagrad (lr_decay = 0): g = grad*grad_scale + wd*p ; sum += g*g ; p -= lr * g / (sqrt(sum)+eps)
__global__ void adagrad_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ st, int64_t n,
                               float lr, float wd, float eps, float gscale_host, const float* __restrict__ gscale_dev) {
  const float gscale = gscale_host * (gscale_dev != nullptr ? __ldg(gscale_dev) : 1.f);
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 sv = reinterpret_cast<float4*>(st)[i];
    float* pp = reinterpret_cast<float*>(&pv);
    const float* gp = reinterpret_cast<const float*>(&gv);
    float* sp = reinterpret_cast<float*>(&sv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gg = gp[j] * gscale + wd * pp[j];
      sp[j] += gg * gg;
      pp[j] -= lr * gg / (sqrtf(sp[j]) + eps);
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(st)[i] = sv;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gg = g[i] * gscale + wd * p[i];
    st[i] += gg * gg;
    p[i] -= lr * gg / (sqrtf(st[i]) + eps);
  }
}

// ---------------------------------------------------------------- window gather (device-resident corpus)
// dst[i, :] = src[idx[i], :] for rows of row_bytes bytes (multiple of 16): builds the [P*T, N*D] clip selection of
// utils/load_dataset.py:56-88 (`feat[chosen]`) from features that already live in HBM.
__global__ void gather_rows_kernel(const uint4* __restrict__ src, const int64_t* __restrict__ idx, int64_t m,
                                   int64_t row_vec, uint4* __restrict__ dst) {
  const int64_t total = m * row_vec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / row_vec, c = i - r * row_vec;
    dst[i] = __ldg(src + idx[r] * row_vec + c);
  }
}

// ---------------------------------------------------------------- UCF-style temporal pooling
// out[b, p, :] = l2norm?( mean_{c in [lo_b, hi_b)} feats[c, p, :] )  with the single clip lo_b when the bin is empty
// (Test/evaluation_UCF.py:54,66-71,77).  grid (bins, patches); thread owns columns tid, tid+blockDim, ...
__global__ void segment_mean_kernel(const float* __restrict__ feats, const int32_t* __restrict__ bounds, int64_t n_clips,
                                    int n_patch, int D, int l2norm, float* __restrict__ out) {
  __shared__ float red[32];
  const int b = blockIdx.x, pch = blockIdx.y;
  int lo = bounds[b], hi = bounds[b + 1];
  if (hi <= lo) hi = lo + 1;
  if (lo >= n_clips) lo = (int)n_clips - 1;
  if (hi > n_clips) hi = (int)n_clips;
  const float inv = 1.0f / (float)(hi - lo);
  float ss = 0.f;
  float* orow = out + ((int64_t)b * n_patch + pch) * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int c = lo; c < hi; ++c) acc += feats[((int64_t)c * n_patch + pch) * D + d];
    acc *= inv;
    orow[d] = acc;
    ss += acc * acc;
  }
  if (!l2norm) return;
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
  const float scale = 1.0f / fmaxf(sqrtf(tot), 1e-12f);  // F.normalize(p=2, dim=-1): x / max(||x||, eps)
  for (int d = threadIdx.x; d < D; d += blockDim.x) orow[d] *= scale;
}

}  // namespace ew
}  // namespace lstc

using namespace lstc;

extern "C" int lstc_gather_rows(const void* src, int64_t row_bytes, const int64_t* idx, int64_t m, void* dst,
                                void* stream) {
  LSTC_CHECK_ARG(src && idx && dst, "lstc_gather_rows: null pointer");
  LSTC_CHECK_ARG(row_bytes > 0 && row_bytes % 16 == 0, "lstc_gather_rows: row_bytes must be a positive multiple of 16");
  LSTC_CHECK_ARG(((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0), "lstc_gather_rows: 16-byte alignment");
  if (m == 0) return LSTC_OK;
  const int64_t row_vec = row_bytes / 16;
  ew::gather_rows_kernel<<<ew::grid_for(m * row_vec, 256), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)src, idx, m, row_vec, (uint4*)dst);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

extern "C" int lstc_segment_mean(const float* feats, const int32_t* bounds, int n_bins, int64_t n_clips, int n_patch,
                                 int D, int l2norm, float* out, void* stream) {
  LSTC_CHECK_ARG(feats && bounds && out, "lstc_segment_mean: null pointer");
  LSTC_CHECK_ARG(n_bins >= 1 && n_clips >= 1 && n_patch >= 1 && D >= 1, "lstc_segment_mean: empty input");
  ew::segment_mean_kernel<<<dim3((unsigned)n_bins, (unsigned)n_patch), 256, 0, (cudaStream_t)stream>>>(
      feats, bounds, n_clips, n_patch, D, l2norm, out);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

extern "C" int lstc_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  LSTC_CHECK_ARG(n == 0 || (src && dst), "lstc_cast_f32_to_bf16: null pointer");
  LSTC_CHECK_ARG(((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0), "lstc_cast_f32_to_bf16: alignment");
  if (n == 0) return LSTC_OK;
  ew::cast_f32_bf16_kernel<<<ew::grid_for((n + 7) / 8, 256), 256, 0, (cudaStream_t)stream>>>(
      src, (__nv_bfloat16*)dst, n);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}
extern "C" int lstc_cast_f32_to_bf16_transposed(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows,
                                               int64_t cols, void* stream) {
  LSTC_CHECK_ARG(rows == 0 || cols == 0 || (src && dst), "lstc_cast_f32_to_bf16_transposed: null pointer");
  LSTC_CHECK_ARG(rows >= 0 && cols >= 0 && ld_src >= cols && ld_dst >= rows,
                 "lstc_cast_f32_to_bf16_transposed: bad shape rows=%lld cols=%lld ld_src=%lld ld_dst=%lld", (long long)rows,
                 (long long)cols, (long long)ld_src, (long long)ld_dst);
  if (rows == 0 || cols == 0) return LSTC_OK;
  LSTC_CHECK_ARG((cols + 63) / 64 < (1ll << 31) && (rows + 63) / 64 < 65536, "lstc_cast_f32_to_bf16_transposed: too large");
  ew::cast_transpose_f32_bf16_kernel<<<dim3((unsigned)((cols + 63) / 64), (unsigned)((rows + 63) / 64)), 256, 0,
                                       (cudaStream_t)stream>>>(src, ld_src, (__nv_bfloat16*)dst, ld_dst, rows, cols);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}
extern "C" int lstc_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream) {
  LSTC_CHECK_ARG(n == 0 || (src && dst), "lstc_cast_bf16_to_f32: null pointer");
  LSTC_CHECK_ARG(((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0), "lstc_cast_bf16_to_f32: alignment");
  if (n == 0) return LSTC_OK;
  ew::cast_bf16_f32_kernel<<<ew::grid_for((n + 7) / 8, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)src, dst, n);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

extern "C" int lstc_cls_prepend_fwd(const void* x, int x_is_f32, const float* cls, const float* pos, float drop_p,
                                    uint64_t seed, uint64_t offset, void* out, int64_t W, int64_t L0, int64_t D,
                                    void* stream) {
  LSTC_CHECK_ARG(x && out, "lstc_cls_prepend_fwd: null pointer");
  LSTC_CHECK_ARG(L0 >= 1 && D % 8 == 0 && D > 0, "lstc_cls_prepend_fwd: need L0 >= 1 and D %% 8 == 0");
  LSTC_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "lstc_cls_prepend_fwd: drop_p out of range");
  LSTC_CHECK_ARG(LSTC_ALIGNED16(x) && LSTC_ALIGNED16(out) && LSTC_ALIGNED16(cls) && LSTC_ALIGNED16(pos),
                 "lstc_cls_prepend_fwd: x, out, cls, pos must be 16-byte aligned");
  if (W == 0) return LSTC_OK;
  const int threads = 128;
  dim3 grid((unsigned)W, (unsigned)((D / 8 + threads - 1) / threads));
  const float dscale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const int use_drop = (drop_p > 0.f && pos != nullptr) ? 1 : 0;
  if (x_is_f32)
    ew::cls_prepend_fwd_kernel<true><<<grid, threads, 0, (cudaStream_t)stream>>>(
        x, cls, pos, dscale, dropout_threshold16(drop_p), seed, offset, use_drop, (__nv_bfloat16*)out, L0, D);
  else
    ew::cls_prepend_fwd_kernel<false><<<grid, threads, 0, (cudaStream_t)stream>>>(
        x, cls, pos, dscale, dropout_threshold16(drop_p), seed, offset, use_drop, (__nv_bfloat16*)out, L0, D);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

extern "C" int lstc_cls_prepend_bwd(const void* g, int cls_learned, float drop_p, uint64_t seed, uint64_t offset,
                                    float* dx, float* dcls, float* dpos, int64_t W, int64_t L0, int64_t D,
                                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LSTC_CHECK_ARG(g != nullptr, "lstc_cls_prepend_bwd: null pointer");
  LSTC_CHECK_ARG(L0 >= 1 && D % 8 == 0 && D > 0, "lstc_cls_prepend_bwd: need L0 >= 1 and D %% 8 == 0");
  LSTC_CHECK_ARG(LSTC_ALIGNED16(g) && LSTC_ALIGNED16(dx), "lstc_cls_prepend_bwd: g, dx must be 16-byte aligned");
  if (dcls) LSTC_CHECK_CUDA(cudaMemsetAsync(dcls, 0, D * sizeof(float), stream));
  if (dpos) LSTC_CHECK_CUDA(cudaMemsetAsync(dpos, 0, (L0 + 1) * D * sizeof(float), stream));
  if (W == 0) return LSTC_OK;
  const int threads = 128;
  dim3 grid((unsigned)W, (unsigned)((D / 8 + threads - 1) / threads));
  const float dscale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const int use_drop = (drop_p > 0.f && dpos != nullptr) ? 1 : 0;
  ew::cls_prepend_bwd_kernel<<<grid, threads, 0, stream>>>((const __nv_bfloat16*)g, cls_learned, dscale,
                                                          dropout_threshold16(drop_p), seed, offset, use_drop, dx,
                                                          dcls, dpos, L0, D);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

extern "C" int64_t lstc_colsum_workspace(int64_t rows, int64_t cols) {
  return (int64_t)ew::colsum_parts(rows) * cols * (int64_t)sizeof(float);
}
extern "C" int lstc_colsum_bf16(const void* x, int64_t rows, int64_t cols, int64_t ld, float* out, void* workspace,
                                void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LSTC_CHECK_ARG(out && workspace && cols > 0, "lstc_colsum_bf16: null pointer / empty");
  if (rows == 0) {
    LSTC_CHECK_CUDA(cudaMemsetAsync(out, 0, cols * sizeof(float), stream));
    return LSTC_OK;
  }
  LSTC_CHECK_ARG(x != nullptr, "lstc_colsum_bf16: null input");
  LSTC_CHECK_ARG(ld % 8 != 0 || LSTC_ALIGNED16(x), "lstc_colsum_bf16: x must be 16-byte aligned when ld %% 8 == 0");
  const int parts = ew::colsum_parts(rows);
  dim3 grid((unsigned)((cols + 255) / 256), (unsigned)parts);
  ew::colsum_stage1_kernel<<<grid, dim3(32, 8), 0, stream>>>((const __nv_bfloat16*)x, rows, cols, ld,
                                                              (float*)workspace);
  LSTC_CHECK_LAUNCH();
  ew::colsum_stage2_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, stream>>>((const float*)workspace, parts, cols,
                                                                              out);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

extern "C" int lstc_dropout_apply_bf16(const void* x, void* y, int64_t rows, int64_t cols, float p, uint64_t seed,
                                       uint64_t offset, void* stream) {
  LSTC_CHECK_ARG(x && y, "lstc_dropout_apply_bf16: null pointer");
  LSTC_CHECK_ARG(cols % 8 == 0, "lstc_dropout_apply_bf16: cols must be a multiple of 8");
  LSTC_CHECK_ARG(p >= 0.f && p < 1.f, "lstc_dropout_apply_bf16: p out of range");
  LSTC_CHECK_ARG(LSTC_ALIGNED16(x) && LSTC_ALIGNED16(y), "lstc_dropout_apply_bf16: x, y must be 16-byte aligned");
  const int64_t n8 = rows * cols / 8;
  if (n8 == 0) return LSTC_OK;
  ew::dropout_apply_kernel<<<ew::grid_for(n8, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n8, 1.f / (1.f - p), dropout_threshold16(p), seed, offset);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}
extern "C" int lstc_dropout_mask(uint8_t* mask, int64_t rows, int64_t cols, float p, uint64_t seed,
                                 uint64_t offset, void* stream) {
  LSTC_CHECK_ARG(mask != nullptr, "lstc_dropout_mask: null pointer");
  LSTC_CHECK_ARG(p >= 0.f && p < 1.f, "lstc_dropout_mask: p out of range");
  if (rows * cols == 0) return LSTC_OK;
  ew::dropout_mask_kernel<<<ew::grid_for(rows * ((cols + 7) / 8), 256), 256, 0, (cudaStream_t)stream>>>(
      mask, rows, cols, dropout_threshold16(p), seed, offset);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

extern "C" int lstc_relbias_gather(const float* table, const int64_t* index, int64_t index_ld, int L, int H,
                                   int64_t T, float* dense, void* stream) {
  LSTC_CHECK_ARG(table && index && dense, "lstc_relbias_gather: null pointer");
  LSTC_CHECK_ARG(L >= 1 && H >= 1 && index_ld >= L - 1, "lstc_relbias_gather: index buffer smaller than L-1");
  ew::relbias_gather_kernel<<<ew::grid_for((int64_t)H * L * L, 256), 256, 0, (cudaStream_t)stream>>>(
      table, index, index_ld, L, H, T, dense);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}
extern "C" int lstc_relbias_scatter(const float* ddense, const int64_t* index, int64_t index_ld, int L, int H,
                                    int64_t T, float* dtable, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LSTC_CHECK_ARG(ddense && index && dtable, "lstc_relbias_scatter: null pointer");
  LSTC_CHECK_ARG(L >= 1 && H >= 1 && index_ld >= L - 1, "lstc_relbias_scatter: index buffer smaller than L-1");
  LSTC_CHECK_CUDA(cudaMemsetAsync(dtable, 0, T * H * sizeof(float), stream));
  ew::relbias_scatter_kernel<<<ew::grid_for((int64_t)H * L * L, 256), 256, 0, stream>>>(ddense, index, index_ld, L, H,
                                                                                       T, dtable);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

extern "C" int lstc_scale_by_device_scalar(const float* src, const float* scalar_dev, float* dst, int64_t n,
                                           void* stream) {
  LSTC_CHECK_ARG(n == 0 || (src && scalar_dev && dst), "lstc_scale_by_device_scalar: null pointer");
  if (n == 0) return LSTC_OK;
  ew::scale_by_scalar_kernel<<<ew::grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(src, scalar_dev, dst, n);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}
extern "C" int lstc_threshold_labels(const float* scores, float thr, float* out, int64_t n, void* stream) {
  LSTC_CHECK_ARG(n == 0 || (scores && out), "lstc_threshold_labels: null pointer");
  if (n == 0) return LSTC_OK;
  ew::threshold_kernel<<<ew::grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(scores, thr, out, n);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}
extern "C" int lstc_sumsq_accumulate(const float* x, int64_t n, float* accum, void* stream) {
  LSTC_CHECK_ARG(accum && (n == 0 || x), "lstc_sumsq_accumulate: null pointer");
  if (n == 0) return LSTC_OK;
  ew::sumsq_kernel<<<ew::grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(x, n, accum);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}
extern "C" int lstc_clip_coef(const float* sumsq, float max_norm, float* coef, void* stream) {
  LSTC_CHECK_ARG(sumsq && coef, "lstc_clip_coef: null pointer");
  ew::clip_coef_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sumsq, max_norm, coef);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}
extern "C" int lstc_adagrad_step(float* param, const float* grad, float* state_sum, int64_t n, float lr,
                                 float weight_decay, float eps, float grad_scale, const float* grad_scale_dev,
                                 void* stream) {
  LSTC_CHECK_ARG(n == 0 || (param && grad && state_sum), "lstc_adagrad_step: null pointer");
  LSTC_CHECK_ARG(((uintptr_t)param % 16 == 0) && ((uintptr_t)grad % 16 == 0) && ((uintptr_t)state_sum % 16 == 0),
                 "lstc_adagrad_step: 16-byte alignment");
  if (n == 0) return LSTC_OK;
  ew::adagrad_kernel<<<ew::grid_for((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(
      param, grad, state_sum, n, lr, weight_decay, eps, grad_scale, grad_scale_dev);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

LSTC_DEFINE_RNG_STEP_SETTER(set_rng_step_elementwise)
