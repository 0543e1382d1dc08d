// MIL ranking + sparsity loss, soft-label cross-entropy, weighted BCE — value and gradient in one launch.
// The reference computes these with a Python loop of B tiny kernels per step
// (Train/temporal_transformer_shanghaitech.py:21-36, Train/spatio_transformer_shanghaitech.py:21-32,
//  Train/spatio_transformer_MIL_CE.py:23-44).  Everything here is a few thousand floats: one CTA,
// fixed-order reductions (bit-reproducible run to run), latency-bound by design.
#include "common.cuh"
#include "../../include/lstc_vad_b200.h"

namespace lstc {
namespace loss {

constexpr int THREADS = 1024;

__device__ __forceinline__ float block_sum(float v, float* red /*[32]*/) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // protect `red` from the previous call's readers
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) s += red[i];
  return s;
}

// dynamic smem: part[2B*P] | bag[2B] | dbag[2B] | sel[2B*k] (int)
__global__ void __launch_bounds__(THREADS)
mil_loss_kernel(const float* __restrict__ scores, int64_t stride, int B, int P, int T, int topk, float lambda1,
                int64_t spar_start, float* __restrict__ out3, int32_t* __restrict__ top_idx,
                float* __restrict__ dscores) {
  extern __shared__ float sm[];
  __shared__ float red[32];
  const int nb = 2 * B;
  const int64_t n = (int64_t)nb * P * T;
  float* part = sm;
  float* bag = part + (int64_t)nb * P;
  float* dbag = bag + nb;
  int* sel = reinterpret_cast<int*>(dbag + nb);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  const float invT = 1.0f / (float)T;

  // 1. part scores = mean over T consecutive scores
  for (int i = tid; i < nb * P; i += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += scores[((int64_t)i * T + t) * stride];
    part[i] = s * invT;
  }
  __syncthreads();
  // 2. per-bag top-k (one warp per bag); ties -> lowest index
  for (int b = warp; b < nb; b += nw) {
    float acc = 0.f;
    for (int k = 0; k < topk; ++k) {
      float best = -INFINITY;
      int besti = 0x7fffffff;
      for (int p = lane; p < P; p += 32) {
        bool taken = false;
        for (int kk = 0; kk < k; ++kk) taken |= (sel[b * topk + kk] == p);
        const float v = part[b * P + p];
        if (!taken && (v > best || (v == best && p < besti))) {
          best = v;
          besti = p;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ob > best || (ob == best && oi < besti)) {
          best = ob;
          besti = oi;
        }
      }
      if (lane == 0) sel[b * topk + k] = besti;
      __syncwarp();
      acc += best;
    }
    if (lane == 0) bag[b] = acc / (float)topk;
  }
  __syncthreads();
  // 3. hinge over all (normal i, abnormal j) pairs
  float e = 0.f;
  for (int i = tid; i < B * B; i += blockDim.x) {
    const int ni = i / B, aj = i % B;
    e += fmaxf(1.f - bag[B + aj] + bag[ni], 0.f);
  }
  const float invB2 = 1.0f / ((float)B * (float)B);
  const float err = block_sum(e, red) * invB2;
  for (int b = tid; b < nb; b += blockDim.x) {
    float gsum = 0.f;
    if (b < B) {
      for (int j = 0; j < B; ++j) gsum += (1.f - bag[B + j] + bag[b] > 0.f) ? 1.f : 0.f;
    } else {
      for (int i = 0; i < B; ++i) gsum -= (1.f - bag[b] + bag[i] > 0.f) ? 1.f : 0.f;
    }
    dbag[b] = gsum * invB2;
  }
  // 4. sparsity term: mean of the flat tail scores[spar_start:]
  float sp = 0.f;
  for (int64_t i = spar_start + tid; i < n; i += blockDim.x) sp += scores[i * stride];
  const int64_t nsp = n - spar_start;
  const float spar = nsp > 0 ? block_sum(sp, red) / (float)nsp : 0.f;
  __syncthreads();
  if (tid == 0) {
    out3[0] = err + lambda1 * spar;
    out3[1] = err;
    out3[2] = spar;
  }
  if (top_idx != nullptr)
    for (int i = tid; i < nb * topk; i += blockDim.x) top_idx[i] = sel[i];
  // 5. gradient
  if (dscores != nullptr) {
    const float gsp = nsp > 0 ? lambda1 / (float)nsp : 0.f;
    for (int64_t i = tid; i < n; i += blockDim.x) dscores[i * stride] = i >= spar_start ? gsp : 0.f;
    __syncthreads();
    for (int i = tid; i < nb * topk; i += blockDim.x) {
      const int b = i / topk;
      const int p = sel[i];
      const float gpart = dbag[b] / (float)topk * invT;
      for (int t = 0; t < T; ++t) dscores[(((int64_t)b * P + p) * T + t) * stride] += gpart;
    }
  }
}

// mean_r( -sum_c lab[r,c] * log_softmax(x[r,:])[c] )
__global__ void __launch_bounds__(THREADS)
soft_ce_kernel(const float* __restrict__ x, const float* __restrict__ lab, int64_t n, int C, float* __restrict__ out1,
               float* __restrict__ dx) {
  __shared__ float red[32];
  float acc = 0.f;
  const float invn = 1.0f / (float)n;
  for (int64_t r = threadIdx.x; r < n; r += blockDim.x) {
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, x[r * C + c]);
    float se = 0.f, sl = 0.f;
    for (int c = 0; c < C; ++c) {
      se += expf(x[r * C + c] - mx);
      sl += lab[r * C + c];
    }
    const float lse = logf(se) + mx;
    for (int c = 0; c < C; ++c) {
      const float lp = x[r * C + c] - lse;
      acc -= lab[r * C + c] * lp;
      if (dx != nullptr) dx[r * C + c] = (expf(lp) * sl - lab[r * C + c]) * invn;
    }
  }
  const float total = block_sum(acc, red);
  if (threadIdx.x == 0) out1[0] = total * invn;
}

// mean over parts of -wn*lab0*log(1-o+1e-8) - wa*lab1*log(o+1e-8),  o = mean of T scores
__global__ void __launch_bounds__(THREADS)
bce_kernel(const float* __restrict__ scores, const float* __restrict__ lab, int64_t n_parts, int T, float wn, float wa,
           float* __restrict__ out1, float* __restrict__ dscores) {
  __shared__ float red[32];
  float acc = 0.f;
  const float invn = 1.0f / (float)n_parts, invT = 1.0f / (float)T;
  for (int64_t r = threadIdx.x; r < n_parts; r += blockDim.x) {
    float o = 0.f;
    for (int t = 0; t < T; ++t) o += scores[r * T + t];
    o *= invT;
    const float l0 = lab[r * 2], l1 = lab[r * 2 + 1];
    acc += -wn * l0 * logf(1.f - o + 1e-8f) - wa * l1 * logf(o + 1e-8f);
    if (dscores != nullptr) {
      const float g = (wn * l0 / (1.f - o + 1e-8f) - wa * l1 / (o + 1e-8f)) * invn * invT;
      for (int t = 0; t < T; ++t) dscores[r * T + t] = g;
    }
  }
  const float total = block_sum(acc, red);
  if (threadIdx.x == 0) out1[0] = total * invn;
}

// ------------------------------------------------------------------------------------------------------
// Frame-level ROC-AUC from per-window scores (utils/eval_utils.py:21-24 = sklearn roc_curve + auc on per-frame
// scores; every frame of a window carries the window's score, Test/evaluation_shanghaitech_ubnormal.py:92-94, so the
// frame-level curve only has one threshold per distinct WINDOW score: sort the windows by score (descending), walk
// the groups of equal scores accumulating positive / negative frame counts and add one trapezoid per group).
// One CTA: bitonic sort of (score, index) in shared memory, then a sequential fp64 trapezoid walk.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
weighted_auc_kernel(const float* __restrict__ scores, const float* __restrict__ pos_w, const float* __restrict__ neg_w,
                    int n, int npad, double* __restrict__ out /* auc, total_pos, total_neg */) {
  extern __shared__ float sm[];
  float* key = sm;                                   // [npad]
  int* idx = reinterpret_cast<int*>(sm + npad);      // [npad]
  for (int i = threadIdx.x; i < npad; i += blockDim.x) {
    key[i] = i < n ? scores[i] : -INFINITY;
    idx[i] = i < n ? i : -1;
  }
  __syncthreads();
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const bool desc = (i & k) == 0;  // overall order: descending
          const float a = key[i], b = key[l];
          // ties broken by index so the sort is deterministic (the AUC itself does not depend on the tie order)
          const bool a_first = (a > b) || (a == b && idx[i] < idx[l]);
          if (desc ? !a_first : a_first) {
            key[i] = b; key[l] = a;
            const int t = idx[i]; idx[i] = idx[l]; idx[l] = t;
          }
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    double tp = 0, fp = 0, tp_prev = 0, fp_prev = 0, area = 0;
    int i = 0;
    while (i < n) {
      const float s = key[i];
      while (i < n && key[i] == s) {
        tp += (double)pos_w[idx[i]];
        fp += (double)neg_w[idx[i]];
        ++i;
      }
      area += (fp - fp_prev) * (tp + tp_prev) * 0.5;
      tp_prev = tp;
      fp_prev = fp;
    }
    out[0] = (tp > 0 && fp > 0) ? area / (tp * fp) : nan("");
    out[1] = tp;
    out[2] = fp;
  }
}

}  // namespace loss
}  // namespace lstc

using namespace lstc;

extern "C" int lstc_weighted_auc(const float* scores, const float* pos_w, const float* neg_w, int64_t n, double* out3,
                                 void* stream) {
  LSTC_CHECK_ARG(scores && pos_w && neg_w && out3, "lstc_weighted_auc: null pointer");
  LSTC_CHECK_ARG(n >= 1 && n <= 16384, "lstc_weighted_auc: n=%lld windows outside [1, 16384] (single-CTA sort)", (long long)n);
  int npad = 1;
  while (npad < n) npad <<= 1;
  const size_t smem = (size_t)npad * 8;
  LSTC_CHECK_CUDA(cudaFuncSetAttribute(loss::weighted_auc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  loss::weighted_auc_kernel<<<1, loss::THREADS, smem, (cudaStream_t)stream>>>(scores, pos_w, neg_w, (int)n, npad, out3);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

extern "C" int lstc_mil_loss(const float* scores, int64_t stride, int B, int P, int T, int topk, float lambda1,
                             int64_t spar_start, float* out3, int32_t* top_idx, float* dscores, void* stream) {
  LSTC_CHECK_ARG(scores && out3, "lstc_mil_loss: null pointer");
  LSTC_CHECK_ARG(B >= 1 && P >= 1 && T >= 1 && stride >= 1, "lstc_mil_loss: bad sizes B=%d P=%d T=%d", B, P, T);
  LSTC_CHECK_ARG(topk >= 1 && topk <= P && topk <= 16, "lstc_mil_loss: topk=%d must be in [1, min(P,16)]", topk);
  const int64_t n = 2ll * B * P * T;
  LSTC_CHECK_ARG(spar_start >= 0 && spar_start <= n, "lstc_mil_loss: spar_start out of range");
  const size_t smem = sizeof(float) * ((size_t)2 * B * P + 4 * (size_t)B) + sizeof(int) * (size_t)2 * B * topk;
  LSTC_CHECK_ARG(smem <= 200 * 1024, "lstc_mil_loss: 2*B*P=%lld parts exceed the single-CTA capacity",
                 (long long)(2ll * B * P));
  LSTC_CHECK_CUDA(cudaFuncSetAttribute(loss::mil_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  loss::mil_loss_kernel<<<1, loss::THREADS, smem, (cudaStream_t)stream>>>(scores, stride, B, P, T, topk, lambda1,
                                                                         spar_start, out3, top_idx, dscores);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

extern "C" int lstc_soft_ce_loss(const float* probs, const float* labels, int64_t n, int C, float* out1,
                                 float* dprobs, void* stream) {
  LSTC_CHECK_ARG(probs && labels && out1, "lstc_soft_ce_loss: null pointer");
  LSTC_CHECK_ARG(n >= 1 && C >= 1, "lstc_soft_ce_loss: empty input");
  loss::soft_ce_kernel<<<1, loss::THREADS, 0, (cudaStream_t)stream>>>(probs, labels, n, C, out1, dprobs);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

extern "C" int lstc_bce_loss(const float* scores, const float* labels, int64_t n_parts, int T, float w_normal,
                             float w_abnormal, float* out1, float* dscores, void* stream) {
  LSTC_CHECK_ARG(scores && labels && out1, "lstc_bce_loss: null pointer");
  LSTC_CHECK_ARG(n_parts >= 1 && T >= 1, "lstc_bce_loss: empty input");
  loss::bce_kernel<<<1, loss::THREADS, 0, (cudaStream_t)stream>>>(scores, labels, n_parts, T, w_normal, w_abnormal,
                                                                 out1, dscores);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}
