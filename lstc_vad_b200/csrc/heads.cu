// Classifier / Regressor tail: layers 2 and 3 of the 3-layer MLP heads plus the final softmax / sigmoid.
// Replaces `nn.Linear(512,32), nn.Dropout, nn.Linear(32,C), nn.Softmax|nn.Sigmoid` of
// models/Classifier.py:9-10 and models/Regressor.py:8-9 (layer 1 = Linear+ReLU+Dropout runs on the
// tcgen05 GEMM with a fused epilogue).  One warp per row; W2^T lives in shared memory, lane j owns
// hidden unit j of the 32-wide layer.
#include "common.cuh"
#include "../../include/lstc_vad_b200.h"

namespace lstc {
namespace heads {

constexpr int HID = 32;
constexpr int WARPS = 8;
constexpr int MAX_C = 4;

__global__ void __launch_bounds__(WARPS * 32)
head_tail_fwd_kernel(const __nv_bfloat16* __restrict__ h1, int64_t n, int K1, const float* __restrict__ W2,
                     const float* __restrict__ b2, const float* __restrict__ W3, const float* __restrict__ b3, int C,
                     int sigmoid, float drop_p, float drop_scale, uint32_t thr16, uint64_t seed, uint64_t offset,
                     __nv_bfloat16* __restrict__ h2, float* __restrict__ out) {
  offset += rng_step();
  extern __shared__ float smem_f[];
  float* w2t = smem_f;                 // [K1][32]
  float* rowbuf = smem_f + K1 * HID;   // [WARPS][K1]
  for (int i = threadIdx.x; i < HID * K1; i += blockDim.x) {
    const int j = i / K1, k = i % K1;  // coalesced read of W2[j][k]
    w2t[k * HID + j] = W2[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* xr = rowbuf + warp * K1;
  const float bias2 = b2[lane];
  float w3[MAX_C];
#pragma unroll
  for (int c = 0; c < MAX_C; ++c) w3[c] = c < C ? W3[c * HID + lane] : 0.f;

  for (int64_t row = (int64_t)blockIdx.x * WARPS + warp; row < n; row += (int64_t)gridDim.x * WARPS) {
    for (int k = lane * 8; k < K1; k += 256) {
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(h1 + row * K1 + k)), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) xr[k + j] = f[j];
    }
    __syncwarp();
    float acc = bias2;
#pragma unroll 8
    for (int k = 0; k < K1; ++k) acc = fmaf(xr[k], w2t[k * HID + lane], acc);
    __syncwarp();
    if (drop_p > 0.f) {
      const uint32_t keep = dropout_keep8(seed, offset, (uint64_t)(row * (HID / 8) + (lane >> 3)), thr16);
      acc = ((keep >> (lane & 7)) & 1u) ? acc * drop_scale : 0.f;
    }
    // the saved activation is rounded to bf16; use the rounded value downstream so backward is consistent
    const __nv_bfloat16 hb = __float2bfloat16(acc);
    h2[row * HID + lane] = hb;
    const float hv = __bfloat162float(hb);
    float z[MAX_C];
#pragma unroll
    for (int c = 0; c < MAX_C; ++c) z[c] = c < C ? warp_sum(hv * w3[c]) + b3[c] : -INFINITY;
    if (lane == 0) {
      if (sigmoid) {
        for (int c = 0; c < C; ++c) out[row * C + c] = 1.f / (1.f + __expf(-z[c]));
      } else {
        float mx = z[0];
        for (int c = 1; c < C; ++c) mx = fmaxf(mx, z[c]);
        float e[MAX_C], s = 0.f;
        for (int c = 0; c < C; ++c) {
          e[c] = __expf(z[c] - mx);
          s += e[c];
        }
        for (int c = 0; c < C; ++c) out[row * C + c] = e[c] / s;
      }
    }
  }
}

__global__ void __launch_bounds__(WARPS * 32)
head_tail_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, const __nv_bfloat16* __restrict__ h2,
                     const float* __restrict__ W3, int64_t n, int C, int sigmoid, float drop_p, float drop_scale,
                     uint32_t thr16, uint64_t seed, uint64_t offset, __nv_bfloat16* __restrict__ dh2,
                     float* __restrict__ dW3, float* __restrict__ db3) {
  offset += rng_step();
  __shared__ float red_w[WARPS][MAX_C][HID];
  __shared__ float red_b[WARPS][MAX_C];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float w3[MAX_C], gw[MAX_C], gb[MAX_C];
#pragma unroll
  for (int c = 0; c < MAX_C; ++c) {
    w3[c] = c < C ? W3[c * HID + lane] : 0.f;
    gw[c] = 0.f;
    gb[c] = 0.f;
  }
  for (int64_t row = (int64_t)blockIdx.x * WARPS + warp; row < n; row += (int64_t)gridDim.x * WARPS) {
    float dz[MAX_C];
    if (sigmoid) {
#pragma unroll
      for (int c = 0; c < MAX_C; ++c) {
        const float o = c < C ? out[row * C + c] : 0.f;
        dz[c] = c < C ? dout[row * C + c] * o * (1.f - o) : 0.f;
      }
    } else {
      float dot = 0.f;
      for (int c = 0; c < C; ++c) dot += out[row * C + c] * dout[row * C + c];
#pragma unroll
      for (int c = 0; c < MAX_C; ++c) dz[c] = c < C ? out[row * C + c] * (dout[row * C + c] - dot) : 0.f;
    }
    const float hv = __bfloat162float(h2[row * HID + lane]);
    float g = 0.f;
#pragma unroll
    for (int c = 0; c < MAX_C; ++c) {
      gw[c] += dz[c] * hv;
      gb[c] += dz[c];
      g += dz[c] * w3[c];
    }
    if (drop_p > 0.f) {
      const uint32_t keep = dropout_keep8(seed, offset, (uint64_t)(row * (HID / 8) + (lane >> 3)), thr16);
      g = ((keep >> (lane & 7)) & 1u) ? g * drop_scale : 0.f;
    }
    dh2[row * HID + lane] = __float2bfloat16(g);
  }
#pragma unroll
  for (int c = 0; c < MAX_C; ++c) {
    red_w[warp][c][lane] = gw[c];
    if (lane == 0) red_b[warp][c] = gb[c];
  }
  __syncthreads();
  if (warp == 0) {
    for (int c = 0; c < C; ++c) {
      float s = 0.f, sb = 0.f;
      for (int w = 0; w < WARPS; ++w) {
        s += red_w[w][c][lane];
        sb += red_b[w][c];
      }
      atomicAdd(dW3 + c * HID + lane, s);
      if (lane == 0) atomicAdd(db3 + c, sb);
    }
  }
}

}  // namespace heads
}  // namespace lstc

using namespace lstc;

extern "C" int lstc_head_tail_fwd(const void* h1, int64_t n, int K1, const float* W2, const float* b2,
                                  const float* W3, const float* b3, int C, int sigmoid, float drop_p, uint64_t seed,
                                  uint64_t offset, void* h2, float* out, void* stream) {
  LSTC_CHECK_ARG(h1 && W2 && b2 && W3 && b3 && h2 && out, "lstc_head_tail_fwd: null pointer");
  LSTC_CHECK_ARG(K1 > 0 && K1 % 8 == 0 && K1 <= 1024, "lstc_head_tail_fwd: K1=%d must be a multiple of 8, <= 1024", K1);
  LSTC_CHECK_ARG(C >= 1 && C <= heads::MAX_C, "lstc_head_tail_fwd: C=%d out of range", C);
  LSTC_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "lstc_head_tail_fwd: drop_p out of range");
  if (n == 0) return LSTC_OK;
  const int smem = (K1 * heads::HID + heads::WARPS * K1) * (int)sizeof(float);
  LSTC_CHECK_CUDA(cudaFuncSetAttribute(heads::head_tail_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int64_t grid = (n + heads::WARPS - 1) / heads::WARPS;
  if (grid > num_sms()) grid = num_sms();
  heads::head_tail_fwd_kernel<<<(unsigned)grid, heads::WARPS * 32, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)h1, n, K1, W2, b2, W3, b3, C, sigmoid, drop_p, drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f,
      dropout_threshold16(drop_p), seed, offset, (__nv_bfloat16*)h2, out);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

extern "C" int lstc_head_tail_bwd(const float* dout, const float* out, const void* h2, const float* W3, int64_t n,
                                  int C, int sigmoid, float drop_p, uint64_t seed, uint64_t offset, void* dh2,
                                  float* dW3, float* db3, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LSTC_CHECK_ARG(dout && out && h2 && W3 && dh2 && dW3 && db3, "lstc_head_tail_bwd: null pointer");
  LSTC_CHECK_ARG(C >= 1 && C <= heads::MAX_C, "lstc_head_tail_bwd: C=%d out of range", C);
  LSTC_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "lstc_head_tail_bwd: drop_p out of range");
  LSTC_CHECK_CUDA(cudaMemsetAsync(dW3, 0, sizeof(float) * C * heads::HID, stream));
  LSTC_CHECK_CUDA(cudaMemsetAsync(db3, 0, sizeof(float) * C, stream));
  if (n == 0) return LSTC_OK;
  int64_t grid = (n + heads::WARPS - 1) / heads::WARPS;
  if (grid > num_sms()) grid = num_sms();
  heads::head_tail_bwd_kernel<<<(unsigned)grid, heads::WARPS * 32, 0, stream>>>(
      dout, out, (const __nv_bfloat16*)h2, W3, n, C, sigmoid, drop_p, drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f,
      dropout_threshold16(drop_p), seed, offset, (__nv_bfloat16*)dh2, dW3, db3);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

LSTC_DEFINE_RNG_STEP_SETTER(set_rng_step_heads)
