// CLS-query attention for the LAST encoder layer.
//
// Every caller of the reference consumes only the CLS row of the encoder output (`feats[:, 0, :]`,
// Train/temporal_transformer_shanghaitech.py:123,181; Test/evaluation_shanghaitech_ubnormal.py:88), so in the last
// layer only the CLS query attends: o_cls = softmax(q_cls K^T / sqrt(dk)) V per (window, head).  The CLS row of the
// relative-position bias is zero (models/MultiHeadAttention.py:111 adds the bias to attn[:, :, 1:, 1:] only), so no
// bias enters.  The dropout mask is the row-0 slice of the mask the full attention kernel draws for the same
// (seed, offset).  This is exact dead-work elimination; it is used only by the opt-in `Encoder.forward_cls` path of
// the harness, never by the drop-in `Encoder.forward`.
//
// One warp per (window, head); lane owns dk/32 consecutive features; HBM-bound (reads K and V once, fp32 math).
#include "common.cuh"
#include "../../include/lstc_vad_b200.h"

namespace lstc {
namespace attn_cls {

constexpr int WARPS = 8;
constexpr int MAXJ = 3;  // L <= 96 keys: lane (j & 31) keeps key j in slot j >> 5

template <int EPL>
__device__ __forceinline__ void load_vec(const __nv_bfloat16* p, float (&f)[EPL]) {
  if (EPL == 8) {
    float t[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(p)), t);
#pragma unroll
    for (int i = 0; i < EPL; ++i) f[i] = t[i];
  } else if (EPL == 4) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    f[0] = bf16lo_to_f32(v.x); f[1] = bf16hi_to_f32(v.x); f[2] = bf16lo_to_f32(v.y); f[3] = bf16hi_to_f32(v.y);
  } else {
    const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(p));
    f[0] = bf16lo_to_f32(v); f[1] = bf16hi_to_f32(v);
  }
}
template <int EPL>
__device__ __forceinline__ void store_vec(__nv_bfloat16* p, const float (&f)[EPL]) {
  if (EPL == 8) {
    float t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = f[i < EPL ? i : 0];
    *reinterpret_cast<uint4*>(p) = pack8(t);
  } else if (EPL == 4) {
    uint2 v;
    v.x = pack_bf16x2(f[0], f[1]);
    v.y = pack_bf16x2(f[2], f[3]);
    *reinterpret_cast<uint2*>(p) = v;
  } else {
    *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(f[0], f[1]);
  }
}

struct Params {
  const __nv_bfloat16* q;   // [W, ld_q]   head h at cols h*dk
  int64_t ld_q;
  const __nv_bfloat16* k;   // [W*L, ld_kv] head h at cols h*dk
  const __nv_bfloat16* v;   // [W*L, ld_kv]
  int64_t ld_kv;
  const __nv_bfloat16* dout;  // bwd: [W, ld_do]
  int64_t ld_do;
  int64_t W;
  int L, H;
  float scale;
  float drop_p, drop_scale;
  uint32_t drop_thr16;
  uint64_t seed, offset;
  __nv_bfloat16* out;   // fwd: o [W, ld_out]; bwd: dq [W, ld_out]
  int64_t ld_out;
  __nv_bfloat16* dk;    // bwd: [W*L, ld_dkv]
  __nv_bfloat16* dv;
  int64_t ld_dkv;
};

// scores of the CLS query against every key, softmax, dropout.  On return lane (j & 31), slot (j >> 5) holds
// p_j (pre-dropout) in pr[] and the post-dropout probability in pd[].
template <int EPL>
__device__ __forceinline__ void cls_probs(const Params& p, int64_t w, int h, int lane, const float (&q)[EPL],
                                          float (&pr)[MAXJ], float (&pd)[MAXJ]) {
  const int dk = EPL * 32;
  const __nv_bfloat16* kbase = p.k + (w * p.L) * p.ld_kv + h * dk + lane * EPL;
#pragma unroll
  for (int s = 0; s < MAXJ; ++s) pr[s] = -INFINITY;
  for (int j0 = 0; j0 < p.L; j0 += 4) {
    float part[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      part[u] = 0.f;
      if (j0 + u < p.L) {
        float kv[EPL];
        load_vec<EPL>(kbase + (int64_t)(j0 + u) * p.ld_kv, kv);
#pragma unroll
        for (int i = 0; i < EPL; ++i) part[u] = fmaf(q[i], kv[i], part[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float sc = warp_sum(part[u]) * p.scale;
      const int j = j0 + u;
      if (j < p.L && (j & 31) == lane) {
#pragma unroll
        for (int s = 0; s < MAXJ; ++s)
          if ((j >> 5) == s) pr[s] = sc;
      }
    }
  }
  float mx = fmaxf(fmaxf(pr[0], pr[1]), pr[2]);
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int s = 0; s < MAXJ; ++s) {
    pr[s] = (pr[s] == -INFINITY) ? 0.f : __expf(pr[s] - mx);
    sum += pr[s];
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  const int64_t grow = (w * p.H + h) * (int64_t)p.L;  // row 0 (the CLS query) of this (window, head)
  const int64_t ld8 = (p.L + 7) >> 3;
#pragma unroll
  for (int s = 0; s < MAXJ; ++s) {
    pr[s] *= inv;
    pd[s] = pr[s];
    const int j = s * 32 + lane;
    if (p.drop_p > 0.f && j < p.L) {
      const uint32_t keep = dropout_keep8(p.seed, p.offset, (uint64_t)(grow * ld8 + (j >> 3)), p.drop_thr16);
      pd[s] = ((keep >> (j & 7)) & 1u) ? pr[s] * p.drop_scale : 0.f;
    }
  }
}

__device__ __forceinline__ float pick(const float (&a)[MAXJ], int j) {
  // value of key j (held by lane j & 31, slot j >> 5), broadcast to the warp
  float v = (j >> 5) == 0 ? a[0] : ((j >> 5) == 1 ? a[1] : a[2]);
  return __shfl_sync(0xffffffffu, v, j & 31);
}

template <int EPL>
__global__ void __launch_bounds__(WARPS * 32) attn_cls_fwd_kernel(const Params p_in) {
  Params p = p_in;
  p.offset += rng_step();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int dk = EPL * 32;
  const int64_t total = p.W * p.H;
  for (int64_t wh = (int64_t)blockIdx.x * WARPS + warp; wh < total; wh += (int64_t)gridDim.x * WARPS) {
    const int64_t w = wh / p.H;
    const int h = (int)(wh % p.H);
    float q[EPL], pr[MAXJ], pd[MAXJ];
    load_vec<EPL>(p.q + w * p.ld_q + h * dk + lane * EPL, q);
    cls_probs<EPL>(p, w, h, lane, q, pr, pd);
    const __nv_bfloat16* vbase = p.v + (w * p.L) * p.ld_kv + h * dk + lane * EPL;
    float o[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) o[i] = 0.f;
    for (int j0 = 0; j0 < p.L; j0 += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        if (j < p.L) {  // warp-uniform
          const float pj = pick(pd, j);
          float vv[EPL];
          load_vec<EPL>(vbase + (int64_t)j * p.ld_kv, vv);
#pragma unroll
          for (int i = 0; i < EPL; ++i) o[i] = fmaf(pj, vv[i], o[i]);
        }
      }
    }
    store_vec<EPL>(p.out + w * p.ld_out + h * dk + lane * EPL, o);
  }
}

template <int EPL>
__global__ void __launch_bounds__(WARPS * 32) attn_cls_bwd_kernel(const Params p_in) {
  Params p = p_in;
  p.offset += rng_step();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int dk = EPL * 32;
  const int64_t total = p.W * p.H;
  for (int64_t wh = (int64_t)blockIdx.x * WARPS + warp; wh < total; wh += (int64_t)gridDim.x * WARPS) {
    const int64_t w = wh / p.H;
    const int h = (int)(wh % p.H);
    float q[EPL], d_o[EPL], pr[MAXJ], pd[MAXJ], ds[MAXJ];
    load_vec<EPL>(p.q + w * p.ld_q + h * dk + lane * EPL, q);
    load_vec<EPL>(p.dout + w * p.ld_do + h * dk + lane * EPL, d_o);
    cls_probs<EPL>(p, w, h, lane, q, pr, pd);
    const int64_t col = h * dk + lane * EPL;
    const __nv_bfloat16* kbase = p.k + (w * p.L) * p.ld_kv + col;
    const __nv_bfloat16* vbase = p.v + (w * p.L) * p.ld_kv + col;
    // dP_j = dO . v_j  (then through the dropout mask), delta = sum_j P_j dP_j
#pragma unroll
    for (int s = 0; s < MAXJ; ++s) ds[s] = 0.f;
    for (int j0 = 0; j0 < p.L; j0 += 4) {
      float part[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        part[u] = 0.f;
        if (j0 + u < p.L) {
          float vv[EPL];
          load_vec<EPL>(vbase + (int64_t)(j0 + u) * p.ld_kv, vv);
#pragma unroll
          for (int i = 0; i < EPL; ++i) part[u] = fmaf(d_o[i], vv[i], part[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float dpj = warp_sum(part[u]);
        const int j = j0 + u;
        if (j < p.L && (j & 31) == lane) {
#pragma unroll
          for (int s = 0; s < MAXJ; ++s)
            if ((j >> 5) == s) ds[s] = dpj;
        }
      }
    }
    float delta = 0.f;
#pragma unroll
    for (int s = 0; s < MAXJ; ++s) {
      // d/dP of the dropped probability is keep/(1-p) = pd/pr where pr > 0
      const float dpm = (p.drop_p > 0.f) ? (pd[s] != 0.f ? ds[s] * p.drop_scale : 0.f) : ds[s];
      ds[s] = dpm;
      delta += pr[s] * dpm;
    }
    delta = warp_sum(delta);
#pragma unroll
    for (int s = 0; s < MAXJ; ++s) ds[s] = pr[s] * (ds[s] - delta);
    // dq = scale * sum_j dS_j k_j ; dk_j = scale * dS_j q ; dv_j = Pd_j dO
    float dq[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) dq[i] = 0.f;
    __nv_bfloat16* dkbase = p.dk + (w * p.L) * p.ld_dkv + col;
    __nv_bfloat16* dvbase = p.dv + (w * p.L) * p.ld_dkv + col;
    for (int j = 0; j < p.L; ++j) {
      const float dsj = pick(ds, j) * p.scale;
      const float pdj = pick(pd, j);
      float kk[EPL], o1[EPL], o2[EPL];
      load_vec<EPL>(kbase + (int64_t)j * p.ld_kv, kk);
#pragma unroll
      for (int i = 0; i < EPL; ++i) {
        dq[i] = fmaf(dsj, kk[i], dq[i]);
        o1[i] = dsj * q[i];
        o2[i] = pdj * d_o[i];
      }
      store_vec<EPL>(dkbase + (int64_t)j * p.ld_dkv, o1);
      store_vec<EPL>(dvbase + (int64_t)j * p.ld_dkv, o2);
    }
    store_vec<EPL>(p.out + w * p.ld_out + col, dq);
  }
}

// dst[r, c] += src[r, c]   (bf16, row-strided views)
__global__ void add_rows_kernel(__nv_bfloat16* __restrict__ dst, int64_t ld_dst, const __nv_bfloat16* __restrict__ src,
                                int64_t ld_src, int64_t rows, int64_t cols8) {
  const int64_t total = rows * cols8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols8, c = (i % cols8) * 8;
    float a[8], b[8];
    unpack8(*reinterpret_cast<const uint4*>(dst + r * ld_dst + c), a);
    unpack8(__ldg(reinterpret_cast<const uint4*>(src + r * ld_src + c)), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    *reinterpret_cast<uint4*>(dst + r * ld_dst + c) = pack8(a);
  }
}

static int fill(Params& p, const void* q, int64_t ld_q, const void* k, const void* v, int64_t ld_kv, int64_t W, int L,
                int H, int dk, float scale, float drop_p, uint64_t seed, uint64_t offset) {
  LSTC_CHECK_ARG(q && k && v, "attn_cls: null pointer");
  LSTC_CHECK_ARG(W >= 0 && L >= 1 && L <= 96 && H >= 1, "attn_cls: bad sizes W=%lld L=%d H=%d (L <= 96)", (long long)W, L, H);
  LSTC_CHECK_ARG(dk == 64 || dk == 128 || dk == 256, "attn_cls: d_k=%d unsupported (64, 128 or 256)", dk);
  LSTC_CHECK_ARG(ld_q % 8 == 0 && ld_kv % 8 == 0, "attn_cls: leading dims must be multiples of 8");
  LSTC_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "attn_cls: dropout_p out of range");
  LSTC_CHECK_ARG(LSTC_ALIGNED16(q) && LSTC_ALIGNED16(k) && LSTC_ALIGNED16(v), "attn_cls: q, k, v must be 16-byte aligned");
  p.q = (const __nv_bfloat16*)q; p.ld_q = ld_q; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v;
  p.ld_kv = ld_kv; p.W = W; p.L = L; p.H = H; p.scale = scale; p.drop_p = drop_p;
  p.drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f; p.drop_thr16 = dropout_threshold16(drop_p);
  p.seed = seed; p.offset = offset;
  return LSTC_OK;
}

static unsigned grid_for(int64_t W, int H) {
  int64_t g = (W * H + WARPS - 1) / WARPS;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace attn_cls
}  // namespace lstc

using namespace lstc;

extern "C" int lstc_attn_cls_fwd(const void* q, int64_t ld_q, const void* k, const void* v, int64_t ld_kv, int64_t W,
                                 int L, int H, int dk, float scale, float dropout_p, uint64_t seed, uint64_t offset,
                                 void* out, int64_t ld_out, void* stream) {
  attn_cls::Params p{};
  int rc = attn_cls::fill(p, q, ld_q, k, v, ld_kv, W, L, H, dk, scale, dropout_p, seed, offset);
  if (rc != LSTC_OK) return rc;
  LSTC_CHECK_ARG(out != nullptr && ld_out % 8 == 0 && LSTC_ALIGNED16(out), "lstc_attn_cls_fwd: bad output");
  if (W == 0) return LSTC_OK;
  p.out = (__nv_bfloat16*)out; p.ld_out = ld_out;
  const unsigned g = attn_cls::grid_for(W, H);
  cudaStream_t st = (cudaStream_t)stream;
  if (dk == 256) attn_cls::attn_cls_fwd_kernel<8><<<g, attn_cls::WARPS * 32, 0, st>>>(p);
  else if (dk == 128) attn_cls::attn_cls_fwd_kernel<4><<<g, attn_cls::WARPS * 32, 0, st>>>(p);
  else attn_cls::attn_cls_fwd_kernel<2><<<g, attn_cls::WARPS * 32, 0, st>>>(p);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

extern "C" int lstc_attn_cls_bwd(const void* q, int64_t ld_q, const void* k, const void* v, int64_t ld_kv,
                                 const void* dout, int64_t ld_do, int64_t W, int L, int H, int dk, float scale,
                                 float dropout_p, uint64_t seed, uint64_t offset, void* dq, int64_t ld_dq, void* dk_out,
                                 void* dv_out, int64_t ld_dkv, void* stream) {
  attn_cls::Params p{};
  int rc = attn_cls::fill(p, q, ld_q, k, v, ld_kv, W, L, H, dk, scale, dropout_p, seed, offset);
  if (rc != LSTC_OK) return rc;
  LSTC_CHECK_ARG(dout && dq && dk_out && dv_out, "lstc_attn_cls_bwd: null pointer");
  LSTC_CHECK_ARG(ld_do % 8 == 0 && ld_dq % 8 == 0 && ld_dkv % 8 == 0, "lstc_attn_cls_bwd: leading dims must be multiples of 8");
  LSTC_CHECK_ARG(LSTC_ALIGNED16(dout) && LSTC_ALIGNED16(dq) && LSTC_ALIGNED16(dk_out) && LSTC_ALIGNED16(dv_out),
                 "lstc_attn_cls_bwd: dout, dq, dk, dv must be 16-byte aligned");
  if (W == 0) return LSTC_OK;
  p.dout = (const __nv_bfloat16*)dout; p.ld_do = ld_do; p.out = (__nv_bfloat16*)dq; p.ld_out = ld_dq;
  p.dk = (__nv_bfloat16*)dk_out; p.dv = (__nv_bfloat16*)dv_out; p.ld_dkv = ld_dkv;
  const unsigned g = attn_cls::grid_for(W, H);
  cudaStream_t st = (cudaStream_t)stream;
  if (dk == 256) attn_cls::attn_cls_bwd_kernel<8><<<g, attn_cls::WARPS * 32, 0, st>>>(p);
  else if (dk == 128) attn_cls::attn_cls_bwd_kernel<4><<<g, attn_cls::WARPS * 32, 0, st>>>(p);
  else attn_cls::attn_cls_bwd_kernel<2><<<g, attn_cls::WARPS * 32, 0, st>>>(p);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

extern "C" int lstc_add_rows_bf16(void* dst, int64_t ld_dst, const void* src, int64_t ld_src, int64_t rows,
                                  int64_t cols, void* stream) {
  LSTC_CHECK_ARG(dst && src, "lstc_add_rows_bf16: null pointer");
  LSTC_CHECK_ARG(cols % 8 == 0 && ld_dst % 8 == 0 && ld_src % 8 == 0, "lstc_add_rows_bf16: cols / lds must be multiples of 8");
  LSTC_CHECK_ARG(LSTC_ALIGNED16(dst) && LSTC_ALIGNED16(src), "lstc_add_rows_bf16: dst, src must be 16-byte aligned");
  if (rows * cols == 0) return LSTC_OK;
  int64_t g = (rows * (cols / 8) + 255) / 256;
  if (g > (int64_t)num_sms() * 8) g = (int64_t)num_sms() * 8;
  attn_cls::add_rows_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)dst, ld_dst,
                                                                         (const __nv_bfloat16*)src, ld_src, rows, cols / 8);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

LSTC_DEFINE_RNG_STEP_SETTER(set_rng_step_attention_cls)
