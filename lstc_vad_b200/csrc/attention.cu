// Fused short-sequence multi-head attention (forward + backward) for windows of L <= 96 tokens.
//
// Replaces models/MultiHeadAttention.py:100-122 of the reference:
//   q,k,v -> [W,H,L,dk] ; S = (q/sqrt(dk)) k^T ; S[:,:,1:,1:] += bias ; P = dropout(softmax(S)) ;
//   O = P v ; O.transpose(1,2).contiguous().view(W, L, H*dv)
// and its autograd backward.
//
// Grid = (window groups, heads).  A CTA owns one head and walks over windows; the whole
// Q/K/V (and dO) tile of a (window, head) pair sits in shared memory (XOR-swizzled 16-byte
// chunks, conflict-free for ldmatrix), the [L,L] score tile lives in registers, softmax uses
// quad shuffles, dropout is regenerated from Philox, and the rel-pos bias / its gradient stay
// in registers across the windows of the CTA (one atomic flush per CTA for dbias).
// Warp w of the CTA owns query rows [16w, 16w+16).
#include "common.cuh"
#include "../../include/lstc_vad_b200.h"

namespace lstc {
namespace attn {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

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
__device__ __forceinline__ void st_shared_b32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// byte offset of 16-byte chunk c16 of row r inside a [rows][DK] bf16 tile (row pitch DK*2 bytes),
// chunk index XOR-swizzled with the row so 8 consecutive rows of one chunk hit 8 distinct bank groups
template <int DK>
__device__ __forceinline__ uint32_t tile_off(int r, int c16) {
  return (uint32_t)(r * (DK * 2) + ((c16 ^ (r & 7)) << 4));
}

struct Params {
  const __nv_bfloat16* qkv;
  int64_t ld;
  const __nv_bfloat16* dout;  // bwd
  int64_t ld_dout;
  int64_t W;
  int L, H;
  const float* bias;  // [H,L,L] or null
  float scale;
  float drop_p, drop_scale;
  uint32_t drop_thr16;
  uint64_t seed, offset;
  __nv_bfloat16* out;  // fwd: O ; bwd: dqkv
  int64_t ld_out;
  float* probs;  // fwd optional
  float* dbias;  // bwd optional
};

// Loads the [L, DK] slab starting at `src` (row pitch ld elements) into a swizzled smem tile, zero-filling
// rows L..LP-1.
template <int LP, int DK>
__device__ __forceinline__ void load_tile(uint32_t s_tile, const __nv_bfloat16* src, int64_t ld, int L) {
  constexpr int CH = DK / 8;
  for (int idx = threadIdx.x; idx < LP * CH; idx += blockDim.x) {
    const int r = idx / CH, c = idx % CH;
    const bool valid = r < L;
    cp_async16(s_tile + tile_off<DK>(r, c), valid ? (const void*)(src + (int64_t)r * ld + c * 8) : (const void*)src,
               valid ? 16 : 0);
  }
}

// S (this warp's 16 query rows x LP keys) = A_tile[m0.., :] * B_tile^T, both tiles [rows][DK] with DK contiguous
template <int LP, int DK>
__device__ __forceinline__ void mma_rows_x_rowsT(float (&s)[LP / 8][4], uint32_t sA, uint32_t sB, int m0, int lane) {
#pragma unroll
  for (int n = 0; n < LP / 8; ++n) {
    s[n][0] = 0.f; s[n][1] = 0.f; s[n][2] = 0.f; s[n][3] = 0.f;
  }
#pragma unroll 4
  for (int kk = 0; kk < DK / 16; ++kk) {
    uint32_t a[4];
    ldsm_x4(sA + tile_off<DK>(m0 + (lane & 7) + ((lane >> 3) & 1) * 8, kk * 2 + (lane >> 4)), a);
#pragma unroll
    for (int np = 0; np < LP / 16; ++np) {
      uint32_t b[4];
      ldsm_x4(sB + tile_off<DK>(np * 16 + (lane & 7) + (lane >> 4) * 8, kk * 2 + ((lane >> 3) & 1)), b);
      mma16816(s[2 * np], a, b[0], b[1]);
      mma16816(s[2 * np + 1], a, b[2], b[3]);
    }
  }
}

// Softmax over the key axis of the warp's score fragment (in place), then dropout.
//   on return: p  = softmax probabilities (pre-dropout)
//              pd = post-dropout probabilities (p * keep / (1-p_drop))
template <int LP>
__device__ __forceinline__ void softmax_dropout(float (&s)[LP / 8][4], float (&pd)[LP / 8][4],
                                                const float (&bfr)[LP / 8][4], const Params& p, int64_t w, int h,
                                                int m0, int lane) {
  const int g = lane >> 2, t = lane & 3;
  constexpr float LOG2E = 1.4426950408889634f;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float mx = -INFINITY;
#pragma unroll
    for (int n = 0; n < LP / 8; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = n * 8 + 2 * t + e;
        float v = s[n][2 * r + e] * p.scale + bfr[n][2 * r + e];
        v = col < p.L ? v : -INFINITY;
        s[n][2 * r + e] = v;
        mx = fmaxf(mx, v);
      }
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float sum = 0.f;
#pragma unroll
    for (int n = 0; n < LP / 8; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float pr = exp2f((s[n][2 * r + e] - mx) * LOG2E);
        s[n][2 * r + e] = pr;
        sum += pr;
      }
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = 1.0f / sum;
    const int row = m0 + g + 8 * r;
    const int64_t grow = (w * p.H + h) * (int64_t)p.L + row;
    const int64_t ld8 = (p.L + 7) >> 3;
#pragma unroll
    for (int n = 0; n < LP / 8; ++n) {
      const float p0 = s[n][2 * r] * inv, p1 = s[n][2 * r + 1] * inv;
      s[n][2 * r] = p0;
      s[n][2 * r + 1] = p1;
      float d0 = p0, d1 = p1;
      if (p.drop_p > 0.f && row < p.L && n * 8 < p.L) {
        const uint32_t keep = dropout_keep8(p.seed, p.offset, (uint64_t)(grow * ld8 + n), p.drop_thr16);
        d0 = ((keep >> (2 * t)) & 1u) ? p0 * p.drop_scale : 0.f;
        d1 = ((keep >> (2 * t + 1)) & 1u) ? p1 * p.drop_scale : 0.f;
      }
      pd[n][2 * r] = d0;
      pd[n][2 * r + 1] = d1;
    }
  }
}

template <int LP>
__device__ __forceinline__ void load_bias_frag(float (&bfr)[LP / 8][4], const float* bias, int h, int L, int m0,
                                               int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int n = 0; n < LP / 8; ++n) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int row = m0 + g + ((e >> 1) << 3);
      const int col = n * 8 + 2 * t + (e & 1);
      bfr[n][e] = (bias != nullptr && row < L && col < L) ? __ldg(bias + ((int64_t)h * L + row) * L + col) : 0.f;
    }
  }
}

// acc[8 n-tiles][4] (16 rows x 64 cols) = sum over k-tiles of A-frags * B_tile[k rows][chunk cols] (B rows = k,
// cols contiguous => ldmatrix.trans).
template <int KT, int DK>
__device__ __forceinline__ void mma_frag_x_rows(float (&acc)[8][4], const uint32_t (&afr)[KT][4], uint32_t sB,
                                                int chunk, int lane) {
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    acc[n][0] = 0.f; acc[n][1] = 0.f; acc[n][2] = 0.f; acc[n][3] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < KT; ++j) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4_t(sB + tile_off<DK>(j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, chunk * 8 + np * 2 + (lane >> 4)), b);
      mma16816(acc[2 * np], afr[j], b[0], b[1]);
      mma16816(acc[2 * np + 1], afr[j], b[2], b[3]);
    }
  }
}

// ==========================================================================================
// Forward
// ==========================================================================================
template <int LP, int DK>
__global__ void __launch_bounds__(LP / 16 * 32) attn_fwd_kernel(const Params p) {
  constexpr int NT = LP / 8, KT = LP / 16, TILE = LP * DK * 2, CH = DK / 8;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem), sK = sQ + TILE, sV = sK + TILE;
  const int h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = warp * 16;
  const int HD = p.H * DK;

  float bfr[NT][4];
  load_bias_frag<LP>(bfr, p.bias, h, p.L, m0, lane);

  for (int64_t w = blockIdx.x; w < p.W; w += gridDim.x) {
    const __nv_bfloat16* base = p.qkv + (w * p.L) * p.ld + h * DK;
    load_tile<LP, DK>(sQ, base, p.ld, p.L);
    load_tile<LP, DK>(sK, base + HD, p.ld, p.L);
    cp_async_commit();
    load_tile<LP, DK>(sV, base + 2 * HD, p.ld, p.L);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    float s[NT][4], pd[NT][4];
    mma_rows_x_rowsT<LP, DK>(s, sQ, sK, m0, lane);
    softmax_dropout<LP>(s, pd, bfr, p, w, h, m0, lane);

    if (p.probs != nullptr) {
#pragma unroll
      for (int n = 0; n < NT; ++n) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int row = m0 + g + ((e >> 1) << 3);
          const int col = n * 8 + 2 * t + (e & 1);
          if (row < p.L && col < p.L) p.probs[((w * p.H + h) * (int64_t)p.L + row) * p.L + col] = pd[n][e];
        }
      }
    }
    uint32_t pa[KT][4];
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      pa[j][0] = pack_bf16x2(pd[2 * j][0], pd[2 * j][1]);
      pa[j][1] = pack_bf16x2(pd[2 * j][2], pd[2 * j][3]);
      pa[j][2] = pack_bf16x2(pd[2 * j + 1][0], pd[2 * j + 1][1]);
      pa[j][3] = pack_bf16x2(pd[2 * j + 1][2], pd[2 * j + 1][3]);
    }
    cp_async_wait<0>();
    __syncthreads();

#pragma unroll 1
    for (int chunk = 0; chunk < DK / 64; ++chunk) {
      float o[8][4];
      mma_frag_x_rows<KT, DK>(o, pa, sV, chunk, lane);
      // stage into this warp's (now dead) Q rows
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        st_shared_b32(sQ + tile_off<DK>(m0 + g, chunk * 8 + n) + 4 * t, pack_bf16x2(o[n][0], o[n][1]));
        st_shared_b32(sQ + tile_off<DK>(m0 + g + 8, chunk * 8 + n) + 4 * t, pack_bf16x2(o[n][2], o[n][3]));
      }
    }
    __syncwarp();
    for (int idx = lane; idx < 16 * CH; idx += 32) {
      const int r = m0 + idx / CH, c = idx % CH;
      if (r < p.L) {
        const uint4 v = ld_shared_v4(sQ + tile_off<DK>(r, c));
        *reinterpret_cast<uint4*>(p.out + (w * p.L + r) * p.ld_out + h * DK + c * 8) = v;
      }
    }
    __syncthreads();  // tiles are overwritten by the next window's loads
  }
}

// ==========================================================================================
// Backward
// ==========================================================================================
// smem: Q | K | V | dO tiles, then per-warp 16x64 staging.  After phase 1 the V tile is dead and is
// re-used for the bf16 copies of Pd (post-dropout probs) and dS, [LP][LP+8] each.
template <int LP, int DK>
__global__ void __launch_bounds__(LP / 16 * 32) attn_bwd_kernel(const Params p) {
  constexpr int NT = LP / 8, KT = LP / 16, TILE = LP * DK * 2, NW = LP / 16;
  constexpr int PP = (LP + 8) * 2;  // padded row pitch (bytes) of the Pd / dS tiles
  constexpr bool ALIAS_V = (2 * LP * PP <= TILE);  // else Pd/dS get their own region after the staging
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem), sK = sQ + TILE, sV = sK + TILE, sDO = sV + TILE;
  const uint32_t sStage = sDO + TILE + (threadIdx.x >> 5) * 2048;  // 16 rows x 128 B per warp
  const uint32_t sPd = ALIAS_V ? sV : (sDO + TILE + NW * 2048), sDS = sPd + LP * PP;
  const int h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = warp * 16;
  const int HD = p.H * DK;

  float bfr[NT][4], dbacc[NT][4];
  load_bias_frag<LP>(bfr, p.bias, h, p.L, m0, lane);
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    dbacc[n][0] = 0.f; dbacc[n][1] = 0.f; dbacc[n][2] = 0.f; dbacc[n][3] = 0.f;
  }

  // writes a finished 16x64 fp32 fragment block (rows row0.., cols chunk*64..) to global through the staging tile
  auto store_chunk = [&](const float (&acc)[8][4], float mul, __nv_bfloat16* gbase, int row0, int chunk) {
    __syncwarp();
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      st_shared_b32(sStage + g * 128 + ((n ^ (g & 7)) << 4) + 4 * t, pack_bf16x2(acc[n][0] * mul, acc[n][1] * mul));
      st_shared_b32(sStage + (g + 8) * 128 + ((n ^ (g & 7)) << 4) + 4 * t,
                    pack_bf16x2(acc[n][2] * mul, acc[n][3] * mul));
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = lane + 32 * i;
      const int r = idx >> 3, c = idx & 7;
      if (row0 + r < p.L) {
        const uint4 v = ld_shared_v4(sStage + r * 128 + ((c ^ (r & 7)) << 4));
        *reinterpret_cast<uint4*>(gbase + (int64_t)(row0 + r) * p.ld_out + chunk * 64 + c * 8) = v;
      }
    }
  };

  for (int64_t w = blockIdx.x; w < p.W; w += gridDim.x) {
    const __nv_bfloat16* base = p.qkv + (w * p.L) * p.ld + h * DK;
    load_tile<LP, DK>(sQ, base, p.ld, p.L);
    load_tile<LP, DK>(sK, base + HD, p.ld, p.L);
    cp_async_commit();
    load_tile<LP, DK>(sV, base + 2 * HD, p.ld, p.L);
    load_tile<LP, DK>(sDO, p.dout + (w * p.L) * p.ld_dout + h * DK, p.ld_dout, p.L);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    // ---- phase 1: P, dP, dS for this warp's query rows ----
    float s[NT][4], pd[NT][4];
    mma_rows_x_rowsT<LP, DK>(s, sQ, sK, m0, lane);
    softmax_dropout<LP>(s, pd, bfr, p, w, h, m0, lane);  // s = P, pd = dropped P
    cp_async_wait<0>();
    __syncthreads();
    float dp[NT][4];
    mma_rows_x_rowsT<LP, DK>(dp, sDO, sV, m0, lane);  // dP_d = dO V^T
    // dropout backward + softmax backward:  dS = P * (dP - sum_j P dP)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float delta = 0.f;
#pragma unroll
      for (int n = 0; n < NT; ++n) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int i = 2 * r + e;
          // d/dP of the dropped prob: keep/(1-p) == pd/P when P > 0 ; use the mask implied by pd
          const float dpm = (p.drop_p > 0.f) ? (pd[n][i] != 0.f ? dp[n][i] * p.drop_scale : 0.f) : dp[n][i];
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
    __syncthreads();  // every warp is done reading V before it is overwritten by Pd / dS
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const uint32_t c = (n * 8 + 2 * t) * 2;
      st_shared_b32(sPd + (m0 + g) * PP + c, pack_bf16x2(pd[n][0], pd[n][1]));
      st_shared_b32(sPd + (m0 + g + 8) * PP + c, pack_bf16x2(pd[n][2], pd[n][3]));
      st_shared_b32(sDS + (m0 + g) * PP + c, pack_bf16x2(dp[n][0], dp[n][1]));
      st_shared_b32(sDS + (m0 + g + 8) * PP + c, pack_bf16x2(dp[n][2], dp[n][3]));
    }
    __syncthreads();

    // ---- phase 2 ----
    __nv_bfloat16* dq_base = p.out + (w * p.L) * p.ld_out + h * DK;
    __nv_bfloat16* dk_base = dq_base + HD;
    __nv_bfloat16* dv_base = dq_base + 2 * HD;
    {
      // dQ[i,:] = scale * sum_j dS[i,j] K[j,:]        (A = dS rows i, non-transposed)
      uint32_t afr[KT][4];
#pragma unroll
      for (int j = 0; j < KT; ++j)
        ldsm_x4(sDS + (m0 + (lane & 7) + ((lane >> 3) & 1) * 8) * PP + (j * 16 + (lane >> 4) * 8) * 2, afr[j]);
#pragma unroll 1
      for (int chunk = 0; chunk < DK / 64; ++chunk) {
        float acc[8][4];
        mma_frag_x_rows<KT, DK>(acc, afr, sK, chunk, lane);
        store_chunk(acc, p.scale, dq_base, m0, chunk);
      }
    }
    {
      // dK[j,:] = scale * sum_i dS[i,j] Q[i,:]        (A = dS^T: transposed ldmatrix of dS[i][j])
      uint32_t afr[KT][4];
#pragma unroll
      for (int i = 0; i < KT; ++i)
        ldsm_x4_t(sDS + (i * 16 + (lane & 7) + (lane >> 4) * 8) * PP + (m0 + ((lane >> 3) & 1) * 8) * 2, afr[i]);
#pragma unroll 1
      for (int chunk = 0; chunk < DK / 64; ++chunk) {
        float acc[8][4];
        mma_frag_x_rows<KT, DK>(acc, afr, sQ, chunk, lane);
        store_chunk(acc, p.scale, dk_base, m0, chunk);
      }
    }
    {
      // dV[j,:] = sum_i Pd[i,j] dO[i,:]
      uint32_t afr[KT][4];
#pragma unroll
      for (int i = 0; i < KT; ++i)
        ldsm_x4_t(sPd + (i * 16 + (lane & 7) + (lane >> 4) * 8) * PP + (m0 + ((lane >> 3) & 1) * 8) * 2, afr[i]);
#pragma unroll 1
      for (int chunk = 0; chunk < DK / 64; ++chunk) {
        float acc[8][4];
        mma_frag_x_rows<KT, DK>(acc, afr, sDO, chunk, lane);
        store_chunk(acc, 1.0f, dv_base, m0, chunk);
      }
    }
    __syncthreads();  // tiles are overwritten by the next window's loads
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
  (void)NW;
}

// ------------------------------------------------------------------------------------------
template <int LP, int DK>
static int launch_fwd(const Params& p, cudaStream_t stream) {
  constexpr int SMEM = 3 * LP * DK * 2;
  auto kern = attn_fwd_kernel<LP, DK>;
  LSTC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  int per_sm = (227 * 1024) / (SMEM + 1024);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int64_t gx = ((int64_t)num_sms() * per_sm + p.H - 1) / p.H;
  if (gx > p.W) gx = p.W;
  if (gx < 1) gx = 1;
  kern<<<dim3((unsigned)gx, (unsigned)p.H), LP / 16 * 32, SMEM, stream>>>(p);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}
template <int LP, int DK>
static int launch_bwd(const Params& p, cudaStream_t stream) {
  constexpr int PDS = 2 * LP * (LP + 8) * 2;
  constexpr int SMEM = 4 * LP * DK * 2 + (LP / 16) * 2048 + (PDS <= LP * DK * 2 ? 0 : PDS);
  auto kern = attn_bwd_kernel<LP, DK>;
  LSTC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  int per_sm = (227 * 1024) / (SMEM + 1024);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int64_t gx = ((int64_t)num_sms() * per_sm + p.H - 1) / p.H;
  if (gx > p.W) gx = p.W;
  if (gx < 1) gx = 1;
  kern<<<dim3((unsigned)gx, (unsigned)p.H), LP / 16 * 32, SMEM, stream>>>(p);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

template <int DK>
static int dispatch_lp(bool bwd, const Params& p, cudaStream_t stream) {
  const int L = p.L;
#define LSTC_ATTN_CASE(LPV)                                                                  \
  if (L <= LPV) return bwd ? launch_bwd<LPV, DK>(p, stream) : launch_fwd<LPV, DK>(p, stream);
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

static int dispatch(bool bwd, const Params& p, int dk, cudaStream_t stream) {
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
  LSTC_CHECK_ARG(W >= 0 && L >= 1 && H >= 1, "lstc_attn_fwd: bad sizes W=%lld L=%d H=%d", (long long)W, L, H);
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
  LSTC_CHECK_ARG(W >= 0 && L >= 1 && H >= 1, "lstc_attn_bwd: bad sizes");
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
