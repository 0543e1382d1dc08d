// Multi-tensor kernels: ONE launch walks a table of tensors (passed by value as a kernel parameter, so there is no
// device-side pointer table to keep alive and the launch is CUDA-graph capturable as is).
//   lstc_multi_pack_bf16   : fp32 gradients of a bucket -> one flat bf16 buffer (the NCCL all-reduce payload; replaces
//                            the torch._foreach_copy_ into a flat fp32 bucket, halves the bytes on NVLink)
//   lstc_multi_unpack_bf16 : the reduced flat bf16 buffer -> the fp32 .grad tensors, in place
//   lstc_multi_adagrad     : torch.optim.Adagrad step (Train/temporal_transformer_shanghaitech.py:83-85,142) for every
//                            parameter of the model in one launch (was one launch per parameter)
// HBM-bound streaming kernels: 16-byte vector accesses, a block owns an 8192-element chunk of one tensor.
#include "common.cuh"
#include "../../include/lstc_vad_b200.h"

namespace lstc {
namespace mt {

constexpr int MT_MAX = 48;          // tensors per launch (the table is a kernel parameter: 48 * 44 B < 4 KB)
constexpr int CHUNK = 8192;         // elements per block
constexpr int THREADS = 256;

struct Table {
  const void* a[MT_MAX];  // pack: fp32 src ; unpack: bf16 src ; adagrad: grad (fp32 or bf16)
  void* b[MT_MAX];        // pack: bf16 dst ; unpack: fp32 dst ; adagrad: param (fp32)
  void* c[MT_MAX];        // adagrad: state sum (fp32)
  int64_t numel[MT_MAX];
  int32_t chunk0[MT_MAX + 1];  // first block of tensor t (prefix sum of ceil(numel / CHUNK))
  float lr[MT_MAX];
  int n;
};

__device__ __forceinline__ int find_tensor(const Table& t, int block) {
  int lo = 0, hi = t.n - 1;
  while (lo < hi) {  // last tensor whose first block is <= block
    const int mid = (lo + hi + 1) >> 1;
    if (t.chunk0[mid] <= block) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(THREADS) pack_bf16_kernel(const __grid_constant__ Table t) {
  const int ti = find_tensor(t, blockIdx.x);
  const int64_t beg = (int64_t)(blockIdx.x - t.chunk0[ti]) * CHUNK;
  const int64_t n = t.numel[ti];
  const int64_t end = beg + CHUNK < n ? beg + CHUNK : n;
  const float* src = reinterpret_cast<const float*>(t.a[ti]);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(t.b[ti]);
  const int64_t end8 = beg + ((end - beg) & ~(int64_t)7);
  for (int64_t i = beg + threadIdx.x * 8; i < end8; i += THREADS * 8) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(src + i));
    const float4 y = __ldg(reinterpret_cast<const float4*>(src + i + 4));
    uint4 o;
    o.x = pack_bf16x2(x.x, x.y); o.y = pack_bf16x2(x.z, x.w);
    o.z = pack_bf16x2(y.x, y.y); o.w = pack_bf16x2(y.z, y.w);
    *reinterpret_cast<uint4*>(dst + i) = o;
  }
  for (int64_t i = end8 + threadIdx.x; i < end; i += THREADS) dst[i] = __float2bfloat16(src[i]);
}

__global__ void __launch_bounds__(THREADS) unpack_bf16_kernel(const __grid_constant__ Table t) {
  const int ti = find_tensor(t, blockIdx.x);
  const int64_t beg = (int64_t)(blockIdx.x - t.chunk0[ti]) * CHUNK;
  const int64_t n = t.numel[ti];
  const int64_t end = beg + CHUNK < n ? beg + CHUNK : n;
  const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(t.a[ti]);
  float* dst = reinterpret_cast<float*>(t.b[ti]);
  const int64_t end8 = beg + ((end - beg) & ~(int64_t)7);
  for (int64_t i = beg + threadIdx.x * 8; i < end8; i += THREADS * 8) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + i));
    float f[8];
    unpack8(v, f);
    *reinterpret_cast<float4*>(dst + i) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(dst + i + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
  for (int64_t i = end8 + threadIdx.x; i < end; i += THREADS) dst[i] = __bfloat162float(src[i]);
}

// g = grad * scale + wd * p ; state += g * g ; p -= lr * g / (sqrt(state) + eps)
template <bool GRAD_BF16>
__global__ void __launch_bounds__(THREADS)
adagrad_kernel(const __grid_constant__ Table t, float wd, float eps, float grad_scale, const float* grad_scale_dev) {
  const int ti = find_tensor(t, blockIdx.x);
  const int64_t beg = (int64_t)(blockIdx.x - t.chunk0[ti]) * CHUNK;
  const int64_t n = t.numel[ti];
  const int64_t end = beg + CHUNK < n ? beg + CHUNK : n;
  float* p = reinterpret_cast<float*>(t.b[ti]);
  float* st = reinterpret_cast<float*>(t.c[ti]);
  const float lr = t.lr[ti];
  const float gs = grad_scale * (grad_scale_dev != nullptr ? __ldg(grad_scale_dev) : 1.0f);
  auto upd = [&](float& pv, float& sv, float gv) {
    const float g = fmaf(wd, pv, gv * gs);
    sv = fmaf(g, g, sv);
    pv -= lr * g / (sqrtf(sv) + eps);
  };
  const int64_t end4 = beg + ((end - beg) & ~(int64_t)3);
  for (int64_t i = beg + threadIdx.x * 4; i < end4; i += THREADS * 4) {
    float4 pv = *reinterpret_cast<const float4*>(p + i);
    float4 sv = *reinterpret_cast<const float4*>(st + i);
    float g[4];
    if (GRAD_BF16) {
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(t.a[ti]) + i));
      g[0] = bf16lo_to_f32(v.x); g[1] = bf16hi_to_f32(v.x); g[2] = bf16lo_to_f32(v.y); g[3] = bf16hi_to_f32(v.y);
    } else {
      const float4 v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(t.a[ti]) + i));
      g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w;
    }
    upd(pv.x, sv.x, g[0]); upd(pv.y, sv.y, g[1]); upd(pv.z, sv.z, g[2]); upd(pv.w, sv.w, g[3]);
    *reinterpret_cast<float4*>(p + i) = pv;
    *reinterpret_cast<float4*>(st + i) = sv;
  }
  for (int64_t i = end4 + threadIdx.x; i < end; i += THREADS) {
    const float gv = GRAD_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(t.a[ti])[i])
                               : reinterpret_cast<const float*>(t.a[ti])[i];
    float pv = p[i], sv = st[i];
    upd(pv, sv, gv);
    p[i] = pv;
    st[i] = sv;
  }
}

// fills one table from entries [first, first + count) and returns the number of blocks
template <typename FillFn>
static int fill_table(Table& t, int first, int count, const int64_t* numel, FillFn fill) {
  t.n = 0;
  int blocks = 0;
  for (int k = 0; k < count; ++k) {
    const int i = first + k;
    if (numel[i] <= 0) continue;
    const int s = t.n++;
    fill(s, i);
    t.numel[s] = numel[i];
    t.chunk0[s] = blocks;
    blocks += (int)((numel[i] + CHUNK - 1) / CHUNK);
  }
  t.chunk0[t.n] = blocks;
  return blocks;
}

}  // namespace mt
}  // namespace lstc

using namespace lstc;

extern "C" int lstc_multi_pack_bf16(const void* const* src_f32, const int64_t* numel, int n, void* const* dst_bf16,
                                    void* stream) {
  LSTC_CHECK_ARG(n >= 0 && (n == 0 || (src_f32 && numel && dst_bf16)), "lstc_multi_pack_bf16: null table");
  for (int i = 0; i < n; ++i)
    LSTC_CHECK_ARG(numel[i] == 0 || (src_f32[i] && dst_bf16[i] && (uintptr_t)src_f32[i] % 16 == 0 &&
                                     (uintptr_t)dst_bf16[i] % 16 == 0),
                   "lstc_multi_pack_bf16: tensor %d: null or not 16-byte aligned", i);
  for (int first = 0; first < n; first += mt::MT_MAX) {
    mt::Table t;
    const int count = n - first < mt::MT_MAX ? n - first : mt::MT_MAX;
    const int blocks = mt::fill_table(t, first, count, numel, [&](int s, int i) {
      t.a[s] = src_f32[i];
      t.b[s] = dst_bf16[i];
    });
    if (blocks == 0) continue;
    mt::pack_bf16_kernel<<<blocks, mt::THREADS, 0, (cudaStream_t)stream>>>(t);
    LSTC_CHECK_LAUNCH();
  }
  return LSTC_OK;
}

extern "C" int lstc_multi_unpack_bf16(const void* const* src_bf16, const int64_t* numel, int n, void* const* dst_f32,
                                      void* stream) {
  LSTC_CHECK_ARG(n >= 0 && (n == 0 || (src_bf16 && numel && dst_f32)), "lstc_multi_unpack_bf16: null table");
  for (int i = 0; i < n; ++i)
    LSTC_CHECK_ARG(numel[i] == 0 || (src_bf16[i] && dst_f32[i] && (uintptr_t)src_bf16[i] % 16 == 0 &&
                                     (uintptr_t)dst_f32[i] % 16 == 0),
                   "lstc_multi_unpack_bf16: tensor %d: null or not 16-byte aligned", i);
  for (int first = 0; first < n; first += mt::MT_MAX) {
    mt::Table t;
    const int count = n - first < mt::MT_MAX ? n - first : mt::MT_MAX;
    const int blocks = mt::fill_table(t, first, count, numel, [&](int s, int i) {
      t.a[s] = src_bf16[i];
      t.b[s] = dst_f32[i];
    });
    if (blocks == 0) continue;
    mt::unpack_bf16_kernel<<<blocks, mt::THREADS, 0, (cudaStream_t)stream>>>(t);
    LSTC_CHECK_LAUNCH();
  }
  return LSTC_OK;
}

extern "C" int lstc_multi_adagrad(void* const* params, const void* const* grads, int grad_is_bf16, void* const* states,
                                  const int64_t* numel, const float* lrs, int n, float weight_decay, float eps,
                                  float grad_scale, const float* grad_scale_dev, void* stream) {
  LSTC_CHECK_ARG(n >= 0 && (n == 0 || (params && grads && states && numel && lrs)), "lstc_multi_adagrad: null table");
  for (int i = 0; i < n; ++i)
    LSTC_CHECK_ARG(numel[i] == 0 || (params[i] && grads[i] && states[i] && (uintptr_t)params[i] % 16 == 0 &&
                                     (uintptr_t)states[i] % 16 == 0 &&
                                     (uintptr_t)grads[i] % (grad_is_bf16 ? 8 : 16) == 0),
                   "lstc_multi_adagrad: tensor %d: null or misaligned", i);
  for (int first = 0; first < n; first += mt::MT_MAX) {
    mt::Table t;
    const int count = n - first < mt::MT_MAX ? n - first : mt::MT_MAX;
    const int blocks = mt::fill_table(t, first, count, numel, [&](int s, int i) {
      t.a[s] = grads[i];
      t.b[s] = params[i];
      t.c[s] = states[i];
      t.lr[s] = lrs[i];
    });
    if (blocks == 0) continue;
    if (grad_is_bf16)
      mt::adagrad_kernel<true><<<blocks, mt::THREADS, 0, (cudaStream_t)stream>>>(t, weight_decay, eps, grad_scale,
                                                                                 grad_scale_dev);
    else
      mt::adagrad_kernel<false><<<blocks, mt::THREADS, 0, (cudaStream_t)stream>>>(t, weight_decay, eps, grad_scale,
                                                                                  grad_scale_dev);
    LSTC_CHECK_LAUNCH();
  }
  return LSTC_OK;
}
