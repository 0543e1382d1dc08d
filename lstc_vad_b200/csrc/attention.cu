// Fused short-sequence multi-head attention (forward + backward) for windows of L <= 96 tokens.
//
// Replaces models/MultiHeadAttention.py:100-122 of the reference:
//   q,k,v -> [W,H,L,dk] ; S = (q/sqrt(dk)) k^T ; S[:,:,1:,1:] += bias ; P = dropout(softmax(S)) ;
//   O = P v ; O.transpose(1,2).contiguous().view(W, L, H*dv)
// and its autograd backward.
//
// Grid = (window groups, heads).  A CTA owns one head and walks over windows.  Q/K/V (and dO) are
// streamed through shared memory in [L x 64-column] chunk tiles by a multi-stage TMA ring (cp.async.bulk.tensor,
// one mbarrier per stage) that runs ahead across window boundaries, so HBM stays busy while the tensor cores work
// (the CTA never holds a whole [L x dk] tile: ~49-114 KB of smem, 2-4 CTAs per SM):
//   forward : S  += Q_c K_c^T (c = 0..dk/64-1) -> softmax / dropout in registers -> O_c = P V_c
//   backward: S  += Q_c K_c^T ; dP += dO_c V_c^T -> P, dS in registers, bf16 copies in smem ->
//             dQ_c = dS K_c ; dK_c = dS^T Q_c ; dV_c = P^T dO_c      (K, Q, dO chunks re-read through L2)
// Math runs on mma.sync.m16n8k16 bf16 (tiles of <= 96 rows are too small for a 128-row tcgen05 tile), fp32
// accumulation, softmax by quad shuffles, dropout regenerated from Philox, rel-pos bias read from a dense
// [H,L,L] table, its gradient accumulated in registers across the windows of the CTA (one atomic flush).
// Warp w of the CTA owns query rows [16w, 16w+16) (and key rows [16w, 16w+16) of dK / dV).
#include <cuda.h>
#include <string.h>

#include <stdlib.h>

#include "common.cuh"
#include "attention_params.cuh"
#include "../../include/lstc_vad_b200.h"

namespace lstc {
namespace attn {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr)
               : "memory");
}
// D(16x8, fp32) += A(16x16, bf16, row) * B(16x8, bf16, col)
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void st_shared_b32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// ---- mbarrier + TMA (tile loads of the stage ring) ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a broken pipeline traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("lstc attention: mbarrier wait timed out (block %d,%d thread %d)\n", (int)blockIdx.x, (int)blockIdx.y,
             (int)threadIdx.x);
      __trap();
    }
  }
}
// [LP x 64] box of window c2 starting at column c0 (row c1 = 0; rows >= L are out of bounds -> zero-filled)
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int32_t c0,
                                            int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// byte offset of 16-byte chunk c16 (0..7) of row r inside a [rows][64] bf16 chunk tile (128-byte rows); the chunk
// index is XOR-swizzled with the row so the 8 rows of one ldmatrix phase hit 8 distinct bank groups
__device__ __forceinline__ uint32_t ct_off(int r, int c16) { return (uint32_t)(r * 128 + ((c16 ^ (r & 7)) << 4)); }


// acc (this warp's 16 rows x LP) += A_c[m0.., 0:64] * B_c[:, 0:64]^T, both chunk tiles [rows][64] (k contiguous)
template <int LP>
__device__ __forceinline__ void mma_acc_rows_x_rowsT(float (&acc)[LP / 8][4], uint32_t sA, uint32_t sB, int m0, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    ldsm_x4(sA + ct_off(m0 + (lane & 7) + ((lane >> 3) & 1) * 8, kk * 2 + (lane >> 4)), a);
#pragma unroll
    for (int np = 0; np < LP / 16; ++np) {
      uint32_t b[4];
      ldsm_x4(sB + ct_off(np * 16 + (lane & 7) + (lane >> 4) * 8, kk * 2 + ((lane >> 3) & 1)), b);
      mma16816(acc[2 * np], a, b[0], b[1]);
      mma16816(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
}

// acc[8 n-tiles][4] (16 rows x 64 cols) = sum over k-tiles of A-frags * B_c[k rows][64 cols]  (B rows = k, cols
// contiguous => ldmatrix.trans)
template <int KT>
__device__ __forceinline__ void mma_frag_x_rows(float (&acc)[8][4], const uint32_t (&afr)[KT][4], uint32_t sB, int lane) {
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    acc[n][0] = 0.f; acc[n][1] = 0.f; acc[n][2] = 0.f; acc[n][3] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < KT; ++j) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4_t(sB + ct_off(j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, np * 2 + (lane >> 4)), b);
      mma16816(acc[2 * np], afr[j], b[0], b[1]);
      mma16816(acc[2 * np + 1], afr[j], b[2], b[3]);
    }
  }
}

// Where the rel-pos bias enters the scores.  Windows longer than 64 tokens: the score accumulators start at
// bias / scale instead of zero, so that after S += Q K^T the softmax's single multiply (S * scale * log2e) yields
// (q k^T * scale + bias) * log2e and the 2 * LP/4 bias loads of a thread are issued before the first stage wait of the
// window (measured on B200, L = 81: backward 1.65 -> 1.50 ms).  Shorter windows: the bias is added inside the softmax,
// where its loads overlap the other row's exponentials (L = 49: 0.258 / 0.610 ms against 0.267 / 0.630 ms).
template <int LP>
struct BiasInAcc {
  static constexpr bool value = LP > 64;
};
template <int LP>
__device__ __forceinline__ void init_scores(float (&s)[LP / 8][4], const Params& p, int h, int m0, int lane) {
  const int g = lane >> 2, t = lane & 3;
  if (!BiasInAcc<LP>::value || p.bias == nullptr) {
#pragma unroll
    for (int n = 0; n < LP / 8; ++n) {
      s[n][0] = 0.f; s[n][1] = 0.f; s[n][2] = 0.f; s[n][3] = 0.f;
    }
    return;
  }
  const float inv_scale = 1.0f / p.scale;
  const int nfull = p.L >> 3;  // n-tiles whose 8 columns are all real keys (warp-uniform)
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = m0 + g + 8 * r;
    // rows >= L are padding (never stored): they read the last real row's bias so the address stays valid
    const float* brow = p.bias + ((int64_t)h * p.L + (row < p.L ? row : p.L - 1)) * p.L + 2 * t;
#pragma unroll
    for (int n = 0; n < LP / 8; ++n) {
      if (n < nfull) {
        s[n][2 * r] = __ldg(brow + n * 8) * inv_scale;
        s[n][2 * r + 1] = __ldg(brow + n * 8 + 1) * inv_scale;
      } else {
#pragma unroll
        for (int e = 0; e < 2; ++e)
          s[n][2 * r + e] = (n * 8 + 2 * t + e < p.L) ? __ldg(brow + n * 8 + e) * inv_scale : 0.f;
      }
    }
  }
}

// Scale, add the bias (unless it is already in the accumulators, see init_scores), mask, softmax over the key axis of the warp's score fragment (in place) and draw the dropout
// mask.  On return s = softmax probabilities (pre-dropout); keep[r] bit (2n+e) tells whether element (row r, n-tile
// n, e) survives dropout (all ones when dropout is off).  dropped(s, keep) gives the post-dropout probability.
template <int LP>
__device__ __forceinline__ void softmax_dropout(float (&s)[LP / 8][4], uint32_t (&keep)[2], const Params& p, int64_t w,
                                                int h, int m0, int lane) {
  const int g = lane >> 2, t = lane & 3;
  constexpr float LOG2E = 1.4426950408889634f;
  const float sl2 = p.scale * LOG2E;  // scores are taken to the log2 domain in the same FFMA that scales them
  const int nfull = p.L >> 3;         // n-tiles whose 8 columns are all real keys (warp-uniform)
  const bool has_bias = !BiasInAcc<LP>::value && p.bias != nullptr;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = m0 + g + 8 * r;
    // rows >= L are padding (never stored): they read the last real row's bias so the address stays valid
    const float* brow = has_bias ? p.bias + ((int64_t)h * p.L + (row < p.L ? row : p.L - 1)) * p.L + 2 * t : nullptr;
    float mx = -INFINITY;
#pragma unroll
    for (int n = 0; n < LP / 8; ++n) {
      if (n < nfull) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float v = s[n][2 * r + e] * sl2;
          if (has_bias) v = fmaf(__ldg(brow + n * 8 + e), LOG2E, v);
          s[n][2 * r + e] = v;
          mx = fmaxf(mx, v);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = n * 8 + 2 * t + e;
          float v = s[n][2 * r + e] * sl2;
          if (has_bias && col < p.L) v = fmaf(__ldg(brow + n * 8 + e), LOG2E, v);
          v = col < p.L ? v : -INFINITY;
          s[n][2 * r + e] = v;
          mx = fmaxf(mx, v);
        }
      }
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float sum = 0.f;
#pragma unroll
    for (int n = 0; n < LP / 8; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float pr = ex2_approx(s[n][2 * r + e] - mx);
        s[n][2 * r + e] = pr;
        sum += pr;
      }
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = 1.0f / sum;
    const int64_t grow = (w * p.H + h) * (int64_t)p.L + row;
    const int64_t ld8 = (p.L + 7) >> 3;
    uint32_t kb = 0xffffffffu;
#pragma unroll
    for (int n = 0; n < LP / 8; ++n) {
      s[n][2 * r] *= inv;
      s[n][2 * r + 1] *= inv;
    }
    if (p.drop_p > 0.f) {  // warp-uniform
      // the 4 lanes of a quad share a row: lane t draws the masks of n-tiles t, t+4, ... and the quad exchanges them
      constexpr int NK = (LP / 8 + 3) / 4;
      uint32_t mine[NK];
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const int n = t + 4 * k;
        mine[k] = 0xffu;
        if (n < LP / 8 && row < p.L && n * 8 < p.L)
          mine[k] = dropout_keep8(p.seed, p.offset, (uint64_t)(grow * ld8 + n), p.drop_thr16);
      }
#pragma unroll
      for (int n = 0; n < LP / 8; ++n) {
        const uint32_t k8 = __shfl_sync(0xffffffffu, mine[n >> 2], (lane & ~3) | (n & 3));
        const uint32_t two = (k8 >> (2 * t)) & 3u;
        kb = (kb & ~(3u << (2 * n))) | (two << (2 * n));
      }
    }
    keep[r] = kb;
  }
}
// post-dropout probability of fragment element (n, i) with i = 2r+e
__device__ __forceinline__ float dropped(float pr, const uint32_t (&keep)[2], int n, int i, float scale) {
  return ((keep[i >> 1] >> (2 * n + (i & 1))) & 1u) ? pr * scale : 0.f;
}

// ------------------------------------------------------------------------------------------
// cp.async stage ring shared by both kernels.  A stage holds two [LP x 64] chunk tiles.
//   PHASES x NC stages per window; phase ph, chunk c loads tile0 / tile1 from the tensors listed below.
// ------------------------------------------------------------------------------------------
// Stage ring filled by TMA: thread 0 issues one cp.async.bulk.tensor per tile (3-D map [W][L][cols], box
// [1][LP][64], SWIZZLE_128B = the ct_off() pattern, rows >= L zero-filled by the hardware), completion on one mbarrier
// per stage; the other threads execute no load instructions at all.  A stage is recycled behind the CTA-wide barrier
// of acquire(), which also publishes the bf16 P / dS tiles of the backward.
template <int LP, int DK, int NS, bool BWD>
struct Ring {
  static constexpr int NC = DK / 64;
  static constexpr int CT = LP * 128;
  static constexpr int STAGE = 2 * CT;
  static constexpr int NPH = BWD ? 4 : 2;  // phases per window, NC stages each

  const Params& p;
  uint32_t base;
  int n_it;  // windows this CTA processes
  // producer cursor (next stage to issue; meaningful in thread 0 only) and consumer slot, maintained incrementally
  int w_p, ph_p, c_p;
  uint32_t off_p, off_c;  // byte offsets of the producer / consumer stage
  const CUtensorMap *tq, *td;
  uint32_t bars;          // NS mbarriers (8 bytes each)
  uint32_t bar_p, bar_c;  // barrier offsets of the producer / consumer stage
  uint32_t par_c;         // phase parity the consumer waits for
  int col0, HD;

  __device__ Ring(const Params& p_, uint32_t base_, uint32_t bars_, int h, const CUtensorMap* tq_,
                  const CUtensorMap* td_)
      : p(p_), base(base_), tq(tq_), td(td_), bars(bars_) {
    const int W = (int)p.W;
    n_it = ((int)blockIdx.x < W) ? (W - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    w_p = blockIdx.x;
    ph_p = 0; c_p = 0;
    off_p = 0; off_c = (NS - 1) * STAGE;         // off_c is advanced to 0 by the first acquire()
    bar_p = 0; bar_c = (NS - 1) * 8; par_c = 1;  // parity flips to 0 when the consumer wraps to slot 0
    HD = p.H * DK;
    col0 = h * DK;
    if (threadIdx.x == 0) {
#pragma unroll
      for (int s = 0; s < NS; ++s) mbar_init(bars + s * 8, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  __device__ __forceinline__ uint32_t cur(int which) const { return base + off_c + which * CT; }

  // one stage: tile 0 (and tile 1 unless `single`) from tensor a / b (0 = qkv, 1 = dout) at column offsets ca / cb
  __device__ __forceinline__ void issue_tma(bool single, int ta, int ca, int tb, int cb) {
    const uint32_t bar = bars + bar_p, t0 = base + off_p;
    mbar_arrive_expect_tx(bar, single ? CT : 2 * CT);
    tma_load_3d(t0, ta ? td : tq, bar, ca, 0, w_p);
    if (!single) tma_load_3d(t0 + CT, tb ? td : tq, bar, cb, 0, w_p);
  }

  __device__ __forceinline__ void issue_next() {  // thread 0 only
    if (w_p < (int)p.W) {
      const int cq = col0 + c_p * 64;
      if (!BWD) {
        if (ph_p == 0) issue_tma(false, 0, cq, 0, cq + HD);  // Q_c, K_c
        else issue_tma(true, 0, cq + 2 * HD, 0, 0);          // V_c
      } else {
        if (ph_p == 0) issue_tma(false, 0, cq, 0, cq + HD);           // Q_c, K_c
        else if (ph_p == 1) issue_tma(false, 1, cq, 0, cq + 2 * HD);  // dO_c, V_c
        else if (ph_p == 2) issue_tma(false, 0, cq + HD, 0, cq);      // K_c, Q_c
        else issue_tma(true, 1, cq, 0, 0);                            // dO_c
      }
      if (++c_p == NC) {
        c_p = 0;
        if (++ph_p == NPH) {
          ph_p = 0;
          w_p += gridDim.x;
        }
      }
      off_p = (off_p + STAGE == NS * STAGE) ? 0u : off_p + STAGE;
      bar_p = (bar_p + 8 == NS * 8) ? 0u : bar_p + 8;
    }
  }
  __device__ __forceinline__ void prologue() {
    if (threadIdx.x == 0) {
#pragma unroll
      for (int i = 0; i < NS - 1; ++i) issue_next();
    }
  }
  // Makes the next stage resident and visible to all warps (-> cur()), then refills the slot consumed one step earlier.
  __device__ __forceinline__ void acquire() {
    off_c = (off_c + STAGE == NS * STAGE) ? 0u : off_c + STAGE;
    bar_c = (bar_c + 8 == NS * 8) ? 0u : bar_c + 8;
    if (bar_c == 0) par_c ^= 1u;
    mbar_wait(bars + bar_c, par_c);
    __syncthreads();
    if (threadIdx.x == 0) issue_next();
  }
};

// writes a finished 16x64 fp32 fragment block (rows row0.., 64 columns at gbase) as bf16: each quad stores 16
// contiguous bytes per row and n-tile; L2 merges the two halves of a 32-byte sector before it reaches HBM
__device__ __forceinline__ void store_chunk(const float (&acc)[8][4], __nv_bfloat16* gbase, int64_t ld_out, int row0,
                                            int L, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const int r0 = row0 + g;
  uint32_t* p0 = reinterpret_cast<uint32_t*>(gbase + (int64_t)r0 * ld_out + 2 * t);
  uint32_t* p1 = reinterpret_cast<uint32_t*>(gbase + (int64_t)(r0 + 8) * ld_out + 2 * t);
  if (r0 < L) {
#pragma unroll
    for (int n = 0; n < 8; ++n) p0[n * 4] = pack_bf16x2(acc[n][0], acc[n][1]);
  }
  if (r0 + 8 < L) {
#pragma unroll
    for (int n = 0; n < 8; ++n) p1[n * 4] = pack_bf16x2(acc[n][2], acc[n][3]);
  }
}

// ==========================================================================================
// Forward
// ==========================================================================================
template <int LP, int DK, int NS>
__global__ void __launch_bounds__(LP / 16 * 32, LP > 64 ? 3 : 4)
attn_fwd_kernel(const Params p_in, const __grid_constant__ CUtensorMap tm_qkv) {
  Params p = p_in;
  p.offset += rng_step();
  constexpr int NT = LP / 8, KT = LP / 16;
  using R = Ring<LP, DK, NS, false>;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s0 = (smem_u32(smem) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  const int h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = warp * 16;
  R ring(p, s0, s0 + NS * R::STAGE, h, &tm_qkv, nullptr);
  ring.prologue();
  for (int it = 0; it < ring.n_it; ++it) {
    const int64_t w = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
    float s[NT][4];
    uint32_t keep[2];
    init_scores<LP>(s, p, h, m0, lane);
#pragma unroll 1
    for (int c = 0; c < R::NC; ++c) {
      ring.acquire();
      mma_acc_rows_x_rowsT<LP>(s, ring.cur(0), ring.cur(1), m0, lane);
    }
    softmax_dropout<LP>(s, keep, p, w, h, m0, lane);
#pragma unroll
    for (int n = 0; n < NT; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) s[n][e] = dropped(s[n][e], keep, n, e, p.drop_scale);  // s = post-dropout P
    }
    if (p.probs != nullptr) {
#pragma unroll
      for (int n = 0; n < NT; ++n) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int row = m0 + g + ((e >> 1) << 3);
          const int col = n * 8 + 2 * t + (e & 1);
          if (row < p.L && col < p.L) p.probs[((w * p.H + h) * (int64_t)p.L + row) * p.L + col] = s[n][e];
        }
      }
    }
    uint32_t pa[KT][4];
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      pa[j][0] = pack_bf16x2(s[2 * j][0], s[2 * j][1]);
      pa[j][1] = pack_bf16x2(s[2 * j][2], s[2 * j][3]);
      pa[j][2] = pack_bf16x2(s[2 * j + 1][0], s[2 * j + 1][1]);
      pa[j][3] = pack_bf16x2(s[2 * j + 1][2], s[2 * j + 1][3]);
    }
    __nv_bfloat16* obase = p.out + (w * p.L) * p.ld_out + h * DK;
#pragma unroll 1
    for (int c = 0; c < R::NC; ++c) {
      ring.acquire();
      float o[8][4];
      mma_frag_x_rows<KT>(o, pa, ring.cur(0), lane);
      store_chunk(o, obase + c * 64, p.ld_out, m0, p.L, lane);
    }
  }
}

// ==========================================================================================
// Backward
// ==========================================================================================
template <int LP, int DK, int NS>
__global__ void __launch_bounds__(LP / 16 * 32, LP <= 64 ? 3 : 2)
attn_bwd_kernel(const Params p_in, const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do) {
  Params p = p_in;
  p.offset += rng_step();
  constexpr int NT = LP / 8, KT = LP / 16;
  constexpr int PP = (LP + 8) * 2;  // padded row pitch (bytes) of the bf16 Pd / dS tiles
  using R = Ring<LP, DK, NS, true>;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s0 = (smem_u32(smem) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  const int h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = warp * 16;
  const int HD = p.H * DK;
  const uint32_t sPd = s0 + NS * R::STAGE, sDS = sPd + LP * PP;

  float dbacc[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    dbacc[n][0] = 0.f; dbacc[n][1] = 0.f; dbacc[n][2] = 0.f; dbacc[n][3] = 0.f;
  }
  R ring(p, s0, sDS + LP * PP, h, &tm_qkv, &tm_do);
  ring.prologue();
  for (int it = 0; it < ring.n_it; ++it) {
    const int64_t w = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
    {
      // ---- phase A: S = Q K^T, dP = dO V^T for this warp's query rows ----
      float s[NT][4], dp[NT][4];
      uint32_t keep[2];
      init_scores<LP>(s, p, h, m0, lane);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        dp[n][0] = 0.f; dp[n][1] = 0.f; dp[n][2] = 0.f; dp[n][3] = 0.f;
      }
#pragma unroll 1
      for (int c = 0; c < R::NC; ++c) {
        ring.acquire();
        mma_acc_rows_x_rowsT<LP>(s, ring.cur(0), ring.cur(1), m0, lane);
      }
#pragma unroll 1
      for (int c = 0; c < R::NC; ++c) {
        ring.acquire();
        mma_acc_rows_x_rowsT<LP>(dp, ring.cur(0), ring.cur(1), m0, lane);
      }
      softmax_dropout<LP>(s, keep, p, w, h, m0, lane);  // s = P, keep = dropout mask bits
      // dropout backward + softmax backward:  dS = P * (dP - sum_j P dP)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float delta = 0.f;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int i = 2 * r + e;
            const float dpm = dropped(dp[n][i], keep, n, i, p.drop_scale);
            dp[n][i] = dpm;
            delta += s[n][i] * dpm;
          }
        }
        delta += __shfl_xor_sync(0xffffffffu, delta, 1);
        delta += __shfl_xor_sync(0xffffffffu, delta, 2);
#pragma unroll
        for (int n = 0; n < NT; ++n) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int i = 2 * r + e;
            const int row = m0 + g + 8 * r, col = n * 8 + 2 * t + e;
            float ds = s[n][i] * (dp[n][i] - delta);
            ds = (row < p.L && col < p.L) ? ds : 0.f;
            dp[n][i] = ds;  // dp now holds dS
            if (row >= 1 && col >= 1) dbacc[n][i] += ds;
          }
        }
      }
      // bf16 copies for the transposed products of phase B (previous window's readers are >= 8 barriers behind)
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const uint32_t cb = (n * 8 + 2 * t) * 2;
        st_shared_b32(sPd + (m0 + g) * PP + cb, pack_bf16x2(dropped(s[n][0], keep, n, 0, p.drop_scale),
                                                            dropped(s[n][1], keep, n, 1, p.drop_scale)));
        st_shared_b32(sPd + (m0 + g + 8) * PP + cb, pack_bf16x2(dropped(s[n][2], keep, n, 2, p.drop_scale),
                                                                dropped(s[n][3], keep, n, 3, p.drop_scale)));
        // the 1/sqrt(dk) of dQ = scale * dS K and dK = scale * dS^T Q is folded into the bf16 copy of dS
        st_shared_b32(sDS + (m0 + g) * PP + cb, pack_bf16x2(dp[n][0] * p.scale, dp[n][1] * p.scale));
        st_shared_b32(sDS + (m0 + g + 8) * PP + cb, pack_bf16x2(dp[n][2] * p.scale, dp[n][3] * p.scale));
      }
    }
    // ---- phase B ----
    __nv_bfloat16* dq_base = p.out + (w * p.L) * p.ld_out + h * DK;
    __nv_bfloat16* dk_base = dq_base + HD;
    __nv_bfloat16* dv_base = dq_base + 2 * HD;
    {
      uint32_t a_ds[KT][4], a_dst[KT][4];
#pragma unroll 1
      for (int c = 0; c < R::NC; ++c) {
        ring.acquire();  // its barrier also publishes Pd / dS of every warp
        if (c == 0) {
#pragma unroll
          for (int j = 0; j < KT; ++j) {
            // A = dS rows of this warp (row i, k = j)
            ldsm_x4(sDS + (m0 + (lane & 7) + ((lane >> 3) & 1) * 8) * PP + (j * 16 + (lane >> 4) * 8) * 2, a_ds[j]);
            // A = dS^T rows of this warp (row j, k = i): transposed ldmatrix of dS[i][j]
            ldsm_x4_t(sDS + (j * 16 + (lane & 7) + (lane >> 4) * 8) * PP + (m0 + ((lane >> 3) & 1) * 8) * 2, a_dst[j]);
          }
        }
        float acc[8][4];
        mma_frag_x_rows<KT>(acc, a_ds, ring.cur(0), lane);   // dQ_c = (scale dS) K_c
        store_chunk(acc, dq_base + c * 64, p.ld_out, m0, p.L, lane);
        mma_frag_x_rows<KT>(acc, a_dst, ring.cur(1), lane);  // dK_c = (scale dS)^T Q_c
        store_chunk(acc, dk_base + c * 64, p.ld_out, m0, p.L, lane);
      }
    }
    {
      uint32_t a_pdt[KT][4];
#pragma unroll
      for (int i = 0; i < KT; ++i)
        ldsm_x4_t(sPd + (i * 16 + (lane & 7) + (lane >> 4) * 8) * PP + (m0 + ((lane >> 3) & 1) * 8) * 2, a_pdt[i]);
#pragma unroll 1
      for (int c = 0; c < R::NC; ++c) {
        ring.acquire();
        float acc[8][4];
        mma_frag_x_rows<KT>(acc, a_pdt, ring.cur(0), lane);  // dV_c = Pd^T dO_c
        store_chunk(acc, dv_base + c * 64, p.ld_out, m0, p.L, lane);
      }
    }
  }
  if (p.dbias != nullptr) {
#pragma unroll
    for (int n = 0; n < NT; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = m0 + g + ((e >> 1) << 3);
        const int col = n * 8 + 2 * t + (e & 1);
        if (row >= 1 && col >= 1 && row < p.L && col < p.L)
          atomicAdd(p.dbias + ((int64_t)h * p.L + row) * p.L + col, dbacc[n][e]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// 3-D bf16 map over [W][L][cols] (row pitch ld elements, window pitch L*ld), box [1][box_rows][64], SWIZZLE_128B,
// zero fill out of bounds (rows L..box_rows-1 of a window)
static int make_tmap3d(CUtensorMap* tm, const void* ptr, int64_t cols, int64_t L, int64_t W, int64_t ld, uint32_t box_rows) {
  const uint64_t gdim[3] = {(uint64_t)cols, (uint64_t)L, (uint64_t)W};
  const uint64_t gstride[2] = {(uint64_t)ld * 2, (uint64_t)L * (uint64_t)ld * 2};
  const uint32_t box[3] = {64u, box_rows, 1u};
  return encode_tmap_bf16_sw128(tm, ptr, 3, gdim, gstride, box, 128);
}

template <int LP, int DK, bool BWD>
static int launch(const Params& p, cudaStream_t stream) {
  constexpr int NS = 3;
  constexpr int NW = LP / 16;
  constexpr int STAGE = 2 * LP * 128;
  // + 1024: alignment slack of the swizzled tiles, + NS mbarriers
  constexpr int SMEM = NS * STAGE + (BWD ? 2 * LP * (LP + 8) * 2 : 0) + 1024 + NS * 8;
  static bool attr_set[64] = {false};
  int dev = 0;
  LSTC_CHECK_CUDA(cudaGetDevice(&dev));
  static int ctas_per_sm[64] = {0};  // resident CTAs per SM for this instantiation and device (occupancy API)
  int& CTAS_PER_SM = ctas_per_sm[(dev >= 0 && dev < 64) ? dev : 0];
  const int HD = p.H * DK;
  CUtensorMap tq, td;
  memset(&tq, 0, sizeof(tq));
  memset(&td, 0, sizeof(td));
  int rc = make_tmap3d(&tq, p.qkv, 3 * (int64_t)HD, p.L, p.W, p.ld, LP);
  if (rc != LSTC_OK) return rc;
  if (BWD) {
    rc = make_tmap3d(&td, p.dout, HD, p.L, p.W, p.ld_dout, LP);
    if (rc != LSTC_OK) return rc;
  }
  if (BWD) {
    auto kern = attn_bwd_kernel<LP, DK, NS>;
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      LSTC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    if (CTAS_PER_SM == 0) LSTC_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&CTAS_PER_SM, kern, NW * 32, SMEM));
    // whole grid co-resident (a partial second wave would double the kernel time): round DOWN
    int64_t gx = ((int64_t)num_sms() * (CTAS_PER_SM > 0 ? CTAS_PER_SM : 1)) / p.H;
    if (gx < 1) gx = 1;
    if (gx > p.W) gx = p.W;
    kern<<<dim3((unsigned)gx, (unsigned)p.H), NW * 32, SMEM, stream>>>(p, tq, td);
  } else {
    auto kern = attn_fwd_kernel<LP, DK, NS>;
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      LSTC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    if (CTAS_PER_SM == 0) LSTC_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&CTAS_PER_SM, kern, NW * 32, SMEM));
    // whole grid co-resident (a partial second wave would double the kernel time): round DOWN
    int64_t gx = ((int64_t)num_sms() * (CTAS_PER_SM > 0 ? CTAS_PER_SM : 1)) / p.H;
    if (gx < 1) gx = 1;
    if (gx > p.W) gx = p.W;
    kern<<<dim3((unsigned)gx, (unsigned)p.H), NW * 32, SMEM, stream>>>(p, tq);
  }
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

template <int DK>
static int dispatch_lp(bool bwd, const Params& p, cudaStream_t stream) {
  const int L = p.L;
#define LSTC_ATTN_CASE(LPV) \
  if (L <= LPV) return bwd ? launch<LPV, DK, true>(p, stream) : launch<LPV, DK, false>(p, stream);
  LSTC_ATTN_CASE(16)
  LSTC_ATTN_CASE(32)
  LSTC_ATTN_CASE(48)
  LSTC_ATTN_CASE(64)
  LSTC_ATTN_CASE(80)
  LSTC_ATTN_CASE(96)
#undef LSTC_ATTN_CASE
  set_last_error("attention: L=%d exceeds the supported maximum of 96 tokens per window", L);
  return LSTC_ERR_UNSUPPORTED;
}

// Implementation choice.  Default: the tcgen05 / TMEM kernels of attention_tc.cu for every supported length (measured
// on B200, dropout + bias, fwd / bwd per layer: L = 49 0.18 / 0.40 ms against 0.26 / 0.61 ms of the mma.sync kernels of
// this file, L = 81 0.42 / 0.93 ms against 0.55 / 1.50 ms).  LSTC_ATTN_IMPL=mma selects the mma.sync kernels
// (A/B measurements); LSTC_ATTN_IMPL=tc is the default spelled out.
static int attn_impl_override() {
  static const int v = [] {
    const char* e = getenv("LSTC_ATTN_IMPL");
    if (e == nullptr) return 0;
    if (strcmp(e, "mma") == 0) return 1;
    if (strcmp(e, "tc") == 0) return 2;
    return 0;
  }();
  return v;
}

static int dispatch(bool bwd, const Params& p, int dk, cudaStream_t stream) {
  const int force = attn_impl_override();
  if (force != 1) return attn_tc::run(bwd, p, dk, stream);
  switch (dk) {
    case 64: return dispatch_lp<64>(bwd, p, stream);
    case 128: return dispatch_lp<128>(bwd, p, stream);
    case 256: return dispatch_lp<256>(bwd, p, stream);
    default:
      set_last_error("attention: d_k=%d unsupported (64, 128 or 256)", dk);
      return LSTC_ERR_UNSUPPORTED;
  }
}

}  // namespace attn
}  // namespace lstc

using namespace lstc;

extern "C" int lstc_attn_fwd(const void* qkv, int64_t ld, int64_t W, int L, int H, int dk, const float* bias,
                             float scale, float dropout_p, uint64_t seed, uint64_t offset, void* out,
                             int64_t ld_out, float* probs, void* stream) {
  LSTC_CHECK_ARG(qkv && out, "lstc_attn_fwd: null pointer");
  LSTC_CHECK_ARG(W >= 0 && W < (int64_t)2000000000 && L >= 1 && H >= 1, "lstc_attn_fwd: bad sizes W=%lld L=%d H=%d",
                 (long long)W, L, H);
  LSTC_CHECK_ARG(ld % 8 == 0 && ld_out % 8 == 0 && ld >= 3 * (int64_t)H * dk && ld_out >= (int64_t)H * dk,
                 "lstc_attn_fwd: leading dims must be multiples of 8 and cover all heads");
  LSTC_CHECK_ARG(((uintptr_t)qkv % 16 == 0) && ((uintptr_t)out % 16 == 0), "lstc_attn_fwd: 16-byte alignment");
  LSTC_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "lstc_attn_fwd: dropout_p out of range");
  if (W == 0) return LSTC_OK;
  attn::Params p{};
  p.qkv = (const __nv_bfloat16*)qkv; p.ld = ld; p.W = W; p.L = L; p.H = H; p.bias = bias; p.scale = scale;
  p.drop_p = dropout_p; p.drop_scale = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
  p.drop_thr16 = dropout_threshold16(dropout_p); p.seed = seed; p.offset = offset;
  p.out = (__nv_bfloat16*)out; p.ld_out = ld_out; p.probs = probs;
  return attn::dispatch(false, p, dk, (cudaStream_t)stream);
}

extern "C" int lstc_attn_bwd(const void* qkv, int64_t ld, const void* dout, int64_t ld_dout, int64_t W, int L,
                             int H, int dk, const float* bias, float scale, float dropout_p, uint64_t seed,
                             uint64_t offset, void* dqkv, int64_t ld_dqkv, float* dbias, void* stream) {
  LSTC_CHECK_ARG(qkv && dout && dqkv, "lstc_attn_bwd: null pointer");
  LSTC_CHECK_ARG(W >= 0 && W < (int64_t)2000000000 && L >= 1 && H >= 1, "lstc_attn_bwd: bad sizes");
  LSTC_CHECK_ARG(ld % 8 == 0 && ld_dout % 8 == 0 && ld_dqkv % 8 == 0 && ld >= 3 * (int64_t)H * dk &&
                     ld_dqkv >= 3 * (int64_t)H * dk && ld_dout >= (int64_t)H * dk,
                 "lstc_attn_bwd: leading dims must be multiples of 8 and cover all heads");
  LSTC_CHECK_ARG(((uintptr_t)qkv % 16 == 0) && ((uintptr_t)dout % 16 == 0) && ((uintptr_t)dqkv % 16 == 0),
                 "lstc_attn_bwd: 16-byte alignment");
  LSTC_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "lstc_attn_bwd: dropout_p out of range");
  if (dbias != nullptr)
    LSTC_CHECK_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)H * L * L, (cudaStream_t)stream));
  if (W == 0) return LSTC_OK;
  attn::Params p{};
  p.qkv = (const __nv_bfloat16*)qkv; p.ld = ld; p.dout = (const __nv_bfloat16*)dout; p.ld_dout = ld_dout;
  p.W = W; p.L = L; p.H = H; p.bias = bias; p.scale = scale;
  p.drop_p = dropout_p; p.drop_scale = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
  p.drop_thr16 = dropout_threshold16(dropout_p); p.seed = seed; p.offset = offset;
  p.out = (__nv_bfloat16*)dqkv; p.ld_out = ld_dqkv; p.dbias = dbias;
  return attn::dispatch(true, p, dk, (cudaStream_t)stream);
}

LSTC_DEFINE_RNG_STEP_SETTER(set_rng_step_attention)
