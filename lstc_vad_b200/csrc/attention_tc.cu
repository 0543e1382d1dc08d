// Fused short-sequence multi-head attention (forward + backward) on the sm_100a tensor cores: tcgen05.mma with the
// score / gradient accumulators in TMEM, operands streamed by TMA, results written back by TMA stores.
//
// Replaces models/MultiHeadAttention.py:100-122 of the reference:
//   q,k,v -> [W,H,L,dk] ; S = (q/sqrt(dk)) k^T ; S[:,:,1:,1:] += bias ; P = dropout(softmax(S)) ;
//   O = P v ; O.transpose(1,2).contiguous().view(W, L, H*dv)
// and its autograd backward.
//
// A window has L <= 96 tokens, far below the 128 rows of a tcgen05 tile, so a TILE stacks G = 128 / LP windows of one
// head (LP = 32, 64 or 128 rows per window, rows >= L zero-filled by TMA): one [G][LP][64] box of q / k / v / dO is a
// [128 x 64] SWIZZLE_128B operand tile.  Scores of the G windows are the diagonal LP x LP blocks of one 128 x 128
// product; probabilities go back to shared memory as a block-diagonal bf16 matrix P (off-diagonal blocks stay zero), so
// that P V, dS K, dS^T Q and P^T dO of all G windows are again single 128-row products:
//   forward : S = sum_c Q_c K_c^T (c = 64-column chunks of dk)  -> softmax / dropout -> O_c = P V_c
//   backward: S = sum_c Q_c K_c^T ; dP = sum_c dO_c V_c^T       -> P, dS            -> dQ_c = dS K_c ;
//             dK_c = dS^T Q_c ; dV_c = P^T dO_c                     (K, Q, dO chunks re-read through L2)
// The same shared-memory tile serves as K-major operand (row = token: Q K^T, dO V^T), as MN-major B operand (token =
// reduction index: P V, dS K, ...) and P / dS as K-major or MN-major A operand (dS vs dS^T) purely through the UMMA
// descriptors - nothing is ever transposed in memory.
//
// One persistent CTA per SM (grid = (CTAs per head, heads)), 10 warps, no CTA-wide barrier after start-up:
//   warp 0     TMA producer: [128 x 64] tiles into an n-stage ring (full / empty mbarriers), runs ahead across tiles
//   warp 1     tcgen05.mma issuer (one thread), owns the 512 TMEM columns
//   warps 2-5  softmax: thread = one score row (tcgen05.ld 32x32b: no shuffles), online max / sum, Philox dropout,
//              dS = P (dP - sum P dP); bf16 P / dS into the block-diagonal smem operands; the rel-pos bias gradient is
//              accumulated in TMEM columns over all tiles of the CTA and flushed once with atomics
//   warps 6-9  epilogue: accumulator chunk (TMEM) -> bf16 -> swizzled smem staging tile -> TMA store (the tensor map
//              clips rows >= L and windows >= W)
// TMEM columns: S [0,128) ; forward: 4 accumulator slots [128,384) ; backward: dP [128,256), 2 accumulator slots
// [256,384), bias gradient [384,512).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "attention_params.cuh"
#include "ptx_sm100.cuh"
#include "../../include/lstc_vad_b200.h"

namespace lstc {
namespace attn_tc {

using attn::Params;
using namespace ptx;

constexpr int NUM_THREADS = 320;    // M = 128 kernel (L > 64)
constexpr int NUM_THREADS64 = 384;  // sub-tile kernel (L <= 64)
constexpr uint32_t TILE_B = 16384;  // [128 rows][64 bf16]
constexpr uint32_t PANEL_B = 16384; // one 64-column panel of the [128 x 128] P / dS operand
constexpr int MAX_STAGES = 12;
constexpr uint32_t SMEM_LIMIT = 232448;  // 227 KB opt-in maximum per CTA
constexpr uint32_t BAR_BYTES = 512;

struct TcParams {
  Params p;
  int tiles;       // tiles per head = ceil(W / G)
  int nkeys;       // N of the score product (multiple of 16): 128 when G > 1, round_up(L, 16) when G == 1
  int ks_tok;      // UMMA_K steps over the token (reduction) dim of P V, dS K, dS^T Q, P^T dO
  int npiece;      // 32-column pieces of one score row
  int n_stages;    // ring depth
  int bias_pitch;  // floats per row of the shared-memory bias copy (odd -> conflict-free column reads)
  uint32_t bias_bytes;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// byte offset of the 16-byte chunk c16 (0..7) of row r in a [rows][128 B] SWIZZLE_128B tile (1024-byte aligned base)
__device__ __forceinline__ uint32_t sw128(int r, int c16) { return (uint32_t)(r * 128 + ((c16 ^ (r & 7)) << 4)); }

// writes 32 consecutive bf16 values of row r starting at column cb (multiple of 32) of a [128 x 128] two-panel operand
__device__ __forceinline__ void store_piece_bf16(uint32_t base, int r, int cb, const float (&v)[32]) {
  const uint32_t pb = base + (uint32_t)(cb >> 6) * PANEL_B;
  const int c0 = (cb & 63) >> 3;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * c + 0], v[8 * c + 1]);
    u.y = pack_bf16x2(v[8 * c + 2], v[8 * c + 3]);
    u.z = pack_bf16x2(v[8 * c + 4], v[8 * c + 5]);
    u.w = pack_bf16x2(v[8 * c + 6], v[8 * c + 7]);
    st_shared_v4(pb + sw128(r, c0 + c), u);
  }
}

template <int LP, int DK, bool BWD>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_tc_kernel(const TcParams tp, const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
               const __grid_constant__ CUtensorMap tm_out) {
  constexpr int G = 128 / LP;
  constexpr int NC = DK / 64;
  constexpr int NACC = BWD ? 2 : 4;
  constexpr uint32_t ACC_COL0 = BWD ? 256u : 128u;
  constexpr uint32_t DP_COL0 = 128u, DB_COL0 = 384u;
  constexpr int MAXP = (G > 1) ? LP / 32 : 3;
  constexpr int NPROD = BWD ? 3 * NC : NC;
  constexpr float LOG2E = 1.4426950408889634f;

  Params p = tp.p;
  p.offset += rng_step();
  const int NS = tp.n_stages;
  const int L = p.L;
  const int h = blockIdx.y;
  const int HD = p.H * DK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s0 = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t sRing = s0;
  const uint32_t sP = sRing + (uint32_t)NS * TILE_B;
  const uint32_t sDS = sP + 2 * PANEL_B;  // backward only
  const uint32_t sStg = BWD ? sDS + 2 * PANEL_B : sDS;
  const uint32_t sBias = sStg + 2 * TILE_B;
  const uint32_t sBar = sBias + tp.bias_bytes;
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (MAX_STAGES + s); };
  const uint32_t s_full = sBar + 8u * (2 * MAX_STAGES), p_full = s_full + 8u;
  auto acc_full = [&](int a) { return sBar + 8u * (2 * MAX_STAGES + 2 + a); };
  auto acc_empty = [&](int a) { return sBar + 8u * (2 * MAX_STAGES + 6 + a); };
  const uint32_t tmem_slot = sBar + 8u * (2 * MAX_STAGES + 10);
  float* bias_s = reinterpret_cast<float*>(smem + (sBias - smem_u32(smem)));

  // ---------------- start-up ----------------
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_out);
    if (BWD) tma_prefetch_desc(&tm_do);
    for (int s = 0; s < NS; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    for (int a = 0; a < NACC; ++a) {
      mbar_init(acc_full(a), 1);
      mbar_init(acc_empty(a), 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  {
    // the block-diagonal operands start (and, off the diagonal blocks, stay) all zero
    const uint32_t nbytes = (BWD ? 4u : 2u) * PANEL_B;
    for (uint32_t o = threadIdx.x * 16u; o < nbytes; o += NUM_THREADS * 16u) st_shared_v4(sP + o, make_uint4(0, 0, 0, 0));
    if (p.bias != nullptr) {
      const float* bsrc = p.bias + (int64_t)h * L * L;
      for (int idx = threadIdx.x; idx < L * L; idx += NUM_THREADS) {
        const int r = idx / L;
        bias_s[r * tp.bias_pitch + (idx - r * L)] = __ldg(bsrc + idx);
      }
    }
  }
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      auto load = [&](const CUtensorMap* tm, int col, int w0) {
        mbar_wait(empty_bar(s), ph ^ 1u, 1);
        mbar_arrive_expect_tx(full_bar(s), TILE_B);
        tma_load_3d(sRing + (uint32_t)s * TILE_B, tm, full_bar(s), col, 0, w0);
        if (++s == NS) { s = 0; ph ^= 1u; }
      };
      const int cq = h * DK, ck = HD + h * DK, cv = 2 * HD + h * DK;
      for (int t = blockIdx.x; t < tp.tiles; t += gridDim.x) {
        const int w0 = t * G;
        for (int c = 0; c < NC; ++c) {
          load(&tm_qkv, cq + 64 * c, w0);
          load(&tm_qkv, ck + 64 * c, w0);
        }
        if (BWD) {
          for (int c = 0; c < NC; ++c) {
            load(&tm_do, cq + 64 * c, w0);
            load(&tm_qkv, cv + 64 * c, w0);
          }
          for (int c = 0; c < NC; ++c) {
            load(&tm_qkv, ck + 64 * c, w0);
            load(&tm_qkv, cq + 64 * c, w0);
            load(&tm_do, cq + 64 * c, w0);
          }
        } else {
          for (int c = 0; c < NC; ++c) load(&tm_qkv, cv + 64 * c, w0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t id_s = idesc_bf16_m128((uint32_t)tp.nkeys, false, false);  // Q K^T, dO V^T
      const uint32_t id_pk = idesc_bf16_m128(64u, false, true);                   // P V, dS K
      const uint32_t id_ptk = idesc_bf16_m128(64u, true, true);                   // dS^T Q, P^T dO
      const int KS = tp.ks_tok;
      int s = 0;
      uint32_t ph = 0, nacc = 0, it = 0;
      auto acquire = [&]() -> uint32_t {
        mbar_wait(full_bar(s), ph, 2);
        const uint32_t a = sRing + (uint32_t)s * TILE_B;
        return a;
      };
      auto release = [&]() {  // the stage acquired last is free once the MMAs issued so far retire
        umma_commit(empty_bar(s));
        if (++s == NS) { s = 0; ph ^= 1u; }
      };
      // acc (128 x nkeys) (+)= A_tile[128 x 64] * B_tile[nkeys x 64]^T over two consecutive ring stages
      auto score_product = [&](uint32_t tmem_d, bool first) {
        const uint32_t a = acquire();
        const int sa = s;
        if (++s == NS) { s = 0; ph ^= 1u; }
        const uint32_t b = acquire();
        tcgen05_fence_after();
        const uint64_t da = desc_kmajor(a), db = desc_kmajor(b);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, da + 2u * k, db + 2u * k, id_s, (!first || k > 0) ? 1u : 0u);
        umma_commit(empty_bar(sa));
        release();
      };
      // acc slot (128 x 64) = A (two-panel smem operand, K- or MN-major) * ring tile (MN-major B: token rows = k)
      auto chunk_product = [&](uint32_t a_base, bool a_mn) {
        const uint32_t b = acquire();
        const uint32_t slot = nacc % NACC;
        mbar_wait(acc_empty(slot), ((nacc / NACC) & 1u) ^ 1u, 3);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + ACC_COL0 + 64u * slot;
        const uint64_t db = desc_mnmajor(b, 8192u);
        if (a_mn) {
          const uint64_t da = desc_mnmajor(a_base, PANEL_B);
          for (int j = 0; j < KS; ++j) umma_bf16(tmem_d, da + 128u * j, db + 128u * j, id_ptk, j > 0 ? 1u : 0u);
        } else {
          for (int j = 0; j < KS; ++j) {
            const uint64_t da = desc_kmajor(a_base + (uint32_t)(j >> 2) * PANEL_B) + 2u * (j & 3);
            umma_bf16(tmem_d, da, db + 128u * j, id_pk, j > 0 ? 1u : 0u);
          }
        }
        release();
        umma_commit(acc_full(slot));
        ++nacc;
      };
      for (int t = blockIdx.x; t < tp.tiles; t += gridDim.x, ++it) {
        for (int c = 0; c < NC; ++c) score_product(tmem_base, c == 0);
        if (BWD)
          for (int c = 0; c < NC; ++c) score_product(tmem_base + DP_COL0, c == 0);
        umma_commit(s_full);
        mbar_wait(p_full, it & 1u, 4);
        tcgen05_fence_after();
        for (int c = 0; c < NC; ++c) {
          if (BWD) {
            chunk_product(sDS, false);  // dQ_c = dS K_c
            chunk_product(sDS, true);   // dK_c = dS^T Q_c
            chunk_product(sP, true);    // dV_c = P^T dO_c
          } else {
            chunk_product(sP, false);   // O_c = P V_c
          }
        }
      }
    }
  } else if (warp < 6) {
    // ===================== softmax warps: thread = score row =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row of the tile = TMEM lane
    const int g = r / LP, i = r - g * LP;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t cb = (G > 1) ? (uint32_t)(g * LP) : 0u;  // first column of this row's diagonal block
    const int npiece = (G > 1) ? MAXP : tp.npiece;
    const float sl2 = p.scale * LOG2E;
    const bool has_bias = p.bias != nullptr;
    const bool drop = p.drop_p > 0.f;
    const int ii = i < L ? i : L - 1;
    const float* brow = bias_s + ii * tp.bias_pitch;
    const int64_t ld8 = (L + 7) >> 3;
    if (BWD && p.dbias != nullptr) {
      uint32_t z[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) z[j] = 0u;
#pragma unroll
      for (int k = 0; k < MAXP; ++k)
        if (k < npiece) tmem_st_32x32b_x32(trow + DB_COL0 + 32u * k, z);
      tmem_st_wait();
    }
    uint32_t it = 0;
    for (int t = blockIdx.x; t < tp.tiles; t += gridDim.x, ++it) {
      const int64_t w = (int64_t)t * G + g;
      const bool valid = (i < L) && (w < p.W);
      const int64_t grow = (w * p.H + h) * (int64_t)L + i;
      mbar_wait(s_full, it & 1u, 5);
      tcgen05_fence_after();
      // ---- pass 1: v = (s * scale + bias) * log2e back into TMEM, online row max / sum ----
      float m = -INFINITY, l = 0.f;
#pragma unroll
      for (int k = 0; k < MAXP; ++k) {
        if (k < npiece) {
          uint32_t rs[32];
          tmem_ld_32x32b_x32(trow + cb + 32u * k, rs);
          tmem_ld_wait();
          float pm = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = 32 * k + j;
            float x = __uint_as_float(rs[j]) * sl2;
            if (has_bias && col < L) x = fmaf(brow[col], LOG2E, x);
            x = col < L ? x : -INFINITY;
            rs[j] = __float_as_uint(x);
            pm = fmaxf(pm, x);
          }
          const float mn = fmaxf(m, pm);
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) acc += ex2_approx(__uint_as_float(rs[j]) - mn);
          l = l * ex2_approx(m - mn) + acc;
          m = mn;
          tmem_st_32x32b_x32(trow + cb + 32u * k, rs);
        }
      }
      tmem_st_wait();
      const float inv = 1.0f / l;
      uint32_t keepb[MAXP];
#pragma unroll
      for (int k = 0; k < MAXP; ++k) {
        uint32_t kb = 0xffffffffu;
        if (drop && k < npiece) {
          kb = 0u;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int n = 4 * k + u;
            uint32_t m8 = 0xffu;
            if (valid && n * 8 < L) m8 = dropout_keep8(p.seed, p.offset, (uint64_t)(grow * ld8 + n), p.drop_thr16);
            kb |= m8 << (8 * u);
          }
        }
        keepb[k] = kb;
      }
      if (!BWD) {
        // ---- pass 2 (forward): P = dropout(softmax) -> bf16 block-diagonal operand ----
#pragma unroll
        for (int k = 0; k < MAXP; ++k) {
          if (k < npiece) {
            uint32_t rs[32];
            tmem_ld_32x32b_x32(trow + cb + 32u * k, rs);
            tmem_ld_wait();
            float pd[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float pr = ex2_approx(__uint_as_float(rs[j]) - m) * inv;
              pd[j] = (valid && ((keepb[k] >> j) & 1u)) ? pr * p.drop_scale : 0.f;
            }
            if (p.probs != nullptr && valid) {
              float* prow = p.probs + grow * (int64_t)L;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (32 * k + j < L) prow[32 * k + j] = pd[j];
            }
            store_piece_bf16(sP, r, (int)cb + 32 * k, pd);
          }
        }
      } else {
        // ---- pass 2 (backward): p and the dropout-masked dP back into TMEM, delta = sum_j p dP ----
        float delta = 0.f;
#pragma unroll
        for (int k = 0; k < MAXP; ++k) {
          if (k < npiece) {
            uint32_t rs[32], rd[32];
            tmem_ld_32x32b_x32(trow + cb + 32u * k, rs);
            tmem_ld_32x32b_x32(trow + DP_COL0 + cb + 32u * k, rd);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float pr = ex2_approx(__uint_as_float(rs[j]) - m) * inv;
              const float d = (((keepb[k] >> j) & 1u) && (32 * k + j < L)) ? __uint_as_float(rd[j]) * p.drop_scale : 0.f;
              delta = fmaf(pr, d, delta);
              rs[j] = __float_as_uint(pr);
              rd[j] = __float_as_uint(d);
            }
            tmem_st_32x32b_x32(trow + cb + 32u * k, rs);
            tmem_st_32x32b_x32(trow + DP_COL0 + cb + 32u * k, rd);
          }
        }
        tmem_st_wait();
        // ---- pass 3: dS = p (dP - delta) ; bias gradient ; bf16 P (post-dropout) and scale * dS operands ----
        const bool want_db = p.dbias != nullptr;
#pragma unroll
        for (int k = 0; k < MAXP; ++k) {
          if (k < npiece) {
            uint32_t rs[32], rd[32], rb[32];
            tmem_ld_32x32b_x32(trow + cb + 32u * k, rs);
            tmem_ld_32x32b_x32(trow + DP_COL0 + cb + 32u * k, rd);
            if (want_db) tmem_ld_32x32b_x32(trow + DB_COL0 + 32u * k, rb);
            tmem_ld_wait();
            float pd[32], ds[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = 32 * k + j;
              const float pr = __uint_as_float(rs[j]);
              float x = pr * (__uint_as_float(rd[j]) - delta);
              x = (valid && col < L) ? x : 0.f;
              if (want_db) rb[j] = __float_as_uint(__uint_as_float(rb[j]) + x);
              ds[j] = x * p.scale;  // the 1/sqrt(dk) of dQ = scale dS K and dK = scale dS^T Q
              pd[j] = (valid && ((keepb[k] >> j) & 1u)) ? pr * p.drop_scale : 0.f;
            }
            if (want_db) tmem_st_32x32b_x32(trow + DB_COL0 + 32u * k, rb);
            store_piece_bf16(sP, r, (int)cb + 32 * k, pd);
            store_piece_bf16(sDS, r, (int)cb + 32 * k, ds);
          }
        }
        tmem_st_wait();
      }
      fence_proxy_async();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    if (BWD && p.dbias != nullptr) {
      // one atomic flush of the bias gradient this CTA accumulated (the CLS row / column carries no bias)
#pragma unroll
      for (int k = 0; k < MAXP; ++k) {
        if (k < npiece) {
          uint32_t rb[32];
          tmem_ld_32x32b_x32(trow + DB_COL0 + 32u * k, rb);
          tmem_ld_wait();
          if (i >= 1 && i < L) {
            float* drow = p.dbias + ((int64_t)h * L + i) * L;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = 32 * k + j;
              if (col >= 1 && col < L) atomicAdd(drow + col, __uint_as_float(rb[j]));
            }
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps: accumulator chunk -> bf16 -> staging tile -> TMA store =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool leader = threadIdx.x == 6 * 32;
    uint32_t n = 0;
    for (int t = blockIdx.x; t < tp.tiles; t += gridDim.x) {
      const int w0 = t * G;
#pragma unroll 1
      for (int prod = 0; prod < NPROD; ++prod, ++n) {
        const uint32_t slot = n % NACC;
        mbar_wait(acc_full(slot), (n / NACC) & 1u, 6);
        tcgen05_fence_after();
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(trow + ACC_COL0 + 64u * slot, r0);
        tmem_ld_32x32b_x32(trow + ACC_COL0 + 64u * slot + 32u, r1);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(slot));
        const uint32_t stg = sStg + (n & 1u) * TILE_B;
        if (leader) tma_wait_group_read<1>();  // the store issued two products ago no longer reads this buffer
        named_bar_sync(1, 128);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(r0[8 * c + 0]), __uint_as_float(r0[8 * c + 1]));
          u.y = pack_bf16x2(__uint_as_float(r0[8 * c + 2]), __uint_as_float(r0[8 * c + 3]));
          u.z = pack_bf16x2(__uint_as_float(r0[8 * c + 4]), __uint_as_float(r0[8 * c + 5]));
          u.w = pack_bf16x2(__uint_as_float(r0[8 * c + 6]), __uint_as_float(r0[8 * c + 7]));
          st_shared_v4(stg + sw128(r, c), u);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(r1[8 * c + 0]), __uint_as_float(r1[8 * c + 1]));
          u.y = pack_bf16x2(__uint_as_float(r1[8 * c + 2]), __uint_as_float(r1[8 * c + 3]));
          u.z = pack_bf16x2(__uint_as_float(r1[8 * c + 4]), __uint_as_float(r1[8 * c + 5]));
          u.w = pack_bf16x2(__uint_as_float(r1[8 * c + 6]), __uint_as_float(r1[8 * c + 7]));
          st_shared_v4(stg + sw128(r, 4 + c), u);
        }
        fence_proxy_async();
        named_bar_sync(1, 128);
        if (leader) {
          const int c = BWD ? prod / 3 : prod;
          const int which = BWD ? prod - 3 * c : 0;  // 0: dQ (or O), 1: dK, 2: dV
          tma_store_3d(&tm_out, stg, which * HD + h * DK + 64 * c, 0, w0);
          tma_commit_group();
        }
      }
    }
    if (leader) tma_wait_group<0>();
  }

  // ---------------- teardown ----------------
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}


// ==========================================================================================================
// Windows of L <= 64 tokens: M = 64 products, two per tile, software-pipelined across tiles.
//
// A tile still holds 128 rows = G = 128 / LP windows (LP = 32 or 64), but every product is issued per 64-row SUB-TILE
// (one window of LP = 64 or two of LP = 32) as tcgen05.mma with M = 64.  An M = 64 accumulator occupies only 16 of
// the 32 lanes of each TMEM quarter (row m -> lane 32 (m / 16) + m % 16), so the second sub-tile goes to lane offset
// 16 of the SAME columns: scores of a tile take 64 columns instead of 128 and nothing in TMEM is wasted.  That leaves
// room to double-buffer S (and dP): the MMA warp issues the score products of tile t+1 BEFORE it waits for the softmax
// of tile t, so tensor pipe, TMA ring and softmax warps no longer take turns:
//   MMA order   : A(0) ; for t: A(t+1), [wait P(t)], B(t)       (A = S, dP ; B = P V or dQ, dK, dV)
//   ring order  : the same (the producer issues loads in exactly the order the MMA warp consumes them)
// P / dS are [64 x 64] bf16 operands per sub-tile (8 KB each, block-diagonal when LP = 32), single-buffered: the softmax
// of tile t+1 runs its TMEM passes while B(t) still reads them and waits on pds_empty only before writing.
// TMEM columns: forward S[2] 0,64 ; accumulator slots 128 + 64 a (a < 4)
//               backward S[2] 0,64 ; dP[2] 128,192 ; slots 256, 320, 448 ; bias gradient 384
// Thread (quarter q, lane l) of a softmax / epilogue warp owns tile row (l / 16) * 64 + 16 q + l % 16.
// ==========================================================================================================
template <int LP, int DK, bool BWD>
__global__ void __launch_bounds__(NUM_THREADS64, 1)
attn_tc64_kernel(const TcParams tp, const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                 const __grid_constant__ CUtensorMap tm_out) {
  static_assert(LP == 32 || LP == 64, "sub-tile kernel: 32 or 64 rows per window");
  constexpr int G = 128 / LP;
  constexpr int NC = DK / 64;
  constexpr int NACC = BWD ? 3 : 4;
  constexpr int MAXP = LP / 32;
  constexpr int NPROD = BWD ? 3 * NC : NC;
  constexpr uint32_t SUB_B = 8192;  // 64 rows x 128 B
  constexpr float LOG2E = 1.4426950408889634f;
  constexpr uint32_t DP_COL0 = 128u, DB_COL0 = 384u;
  auto acc_col = [](uint32_t a) -> uint32_t { return BWD ? (a < 2 ? 256u + 64u * a : 448u) : 128u + 64u * a; };

  Params p = tp.p;
  p.offset += rng_step();
  const int NS = tp.n_stages;
  const int L = p.L;
  const int h = blockIdx.y;
  const int HD = p.H * DK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s0 = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t sRing = s0;
  const uint32_t sP = sRing + (uint32_t)NS * TILE_B;  // P of sub-tile g at + g * 8192
  const uint32_t sDS = sP + 2 * SUB_B;                // backward only
  const uint32_t sStg = BWD ? sDS + 2 * SUB_B : sDS;
  const uint32_t sBias = sStg + 2 * TILE_B;
  const uint32_t sBar = sBias + tp.bias_bytes;
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (MAX_STAGES + s); };
  auto s_full = [&](uint32_t b) { return sBar + 8u * (2 * MAX_STAGES + b); };
  const uint32_t p_full = sBar + 8u * (2 * MAX_STAGES + 2), p_empty = p_full + 8u;
  auto acc_full = [&](int a) { return sBar + 8u * (2 * MAX_STAGES + 4 + a); };
  auto acc_empty = [&](int a) { return sBar + 8u * (2 * MAX_STAGES + 8 + a); };
  const uint32_t tmem_slot = sBar + 8u * (2 * MAX_STAGES + 12);
  float* bias_s = reinterpret_cast<float*>(smem + (sBias - smem_u32(smem)));

  // ---------------- start-up ----------------
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_out);
    if (BWD) tma_prefetch_desc(&tm_do);
    for (int s = 0; s < NS; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(s_full(0), 1);
    mbar_init(s_full(1), 1);
    mbar_init(p_full, 4);
    mbar_init(p_empty, 1);
    for (int a = 0; a < NACC; ++a) {
      mbar_init(acc_full(a), 1);
      mbar_init(acc_empty(a), 4);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  {
    const uint32_t nbytes = (BWD ? 4u : 2u) * SUB_B;
    for (uint32_t o = threadIdx.x * 16u; o < nbytes; o += NUM_THREADS64 * 16u) st_shared_v4(sP + o, make_uint4(0, 0, 0, 0));
    // shared-memory bias rows: bias * log2e (0 without a bias) in the columns < L, -inf in the padding columns, so the
    // softmax needs no column mask
    const float* bsrc = p.bias != nullptr ? p.bias + (int64_t)h * L * L : nullptr;
    for (int idx = threadIdx.x; idx < L * LP; idx += NUM_THREADS64) {
      const int r = idx / LP, c = idx - r * LP;
      bias_s[r * tp.bias_pitch + c] = c < L ? (bsrc != nullptr ? __ldg(bsrc + r * L + c) * LOG2E : 0.f) : -INFINITY;
    }
  }
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  const int n_my = ((int)blockIdx.x < tp.tiles) ? (tp.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  // register re-allocation per warpgroup (the kernel launches with 168 per thread): softmax threads hold a whole score
  // row + its dS / bias-gradient pieces
  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 96;" ::: "memory");
  }
  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      auto load = [&](const CUtensorMap* tm, int col, int w0, uint64_t policy) {
        mbar_wait(empty_bar(s), ph ^ 1u, 1);
        mbar_arrive_expect_tx(full_bar(s), TILE_B);
        tma_load_3d_hint(sRing + (uint32_t)s * TILE_B, tm, full_bar(s), col, 0, w0, policy);
        if (++s == NS) { s = 0; ph ^= 1u; }
      };
      // backward: q, k, dO are read again by the chunk products of the same tile -> keep them in L2 until then;
      // everything that is read for the last time leaves L2 first
      const uint64_t pol_again = BWD ? L2_EVICT_LAST : L2_EVICT_FIRST, pol_once = L2_EVICT_FIRST;
      const int cq = h * DK, ck = HD + h * DK, cv = 2 * HD + h * DK;
      // load order = the order the MMA warp consumes: A(0), A(1), then per tile and 64-column chunk c the operands of
      // the chunk products of tile `it` followed by the score operands of tile `it + 2`
      auto load_a = [&](int it, int c) {
        const int w0 = ((int)blockIdx.x + it * (int)gridDim.x) * G;
        load(&tm_qkv, cq + 64 * c, w0, pol_again);
        load(&tm_qkv, ck + 64 * c, w0, pol_again);
        if (BWD) {
          load(&tm_do, cq + 64 * c, w0, pol_again);
          load(&tm_qkv, cv + 64 * c, w0, pol_once);
        }
      };
      auto load_b = [&](int it, int c) {
        const int w0 = ((int)blockIdx.x + it * (int)gridDim.x) * G;
        if (BWD) {
          load(&tm_qkv, ck + 64 * c, w0, pol_once);
          load(&tm_qkv, cq + 64 * c, w0, pol_once);
          load(&tm_do, cq + 64 * c, w0, pol_once);
        } else {
          load(&tm_qkv, cv + 64 * c, w0, pol_once);
        }
      };
      for (int it = 0; it < 2 && it < n_my; ++it)
        for (int c = 0; c < NC; ++c) load_a(it, c);
      for (int it = 0; it < n_my; ++it) {
        for (int c = 0; c < NC; ++c) {
          load_b(it, c);
          if (it + 2 < n_my) load_a(it + 2, c);
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t M64 = (64u >> 4) << 24, M128 = (128u >> 4) << 24;
      const uint32_t id_s = (idesc_bf16_m128(64u, false, false) & ~M128) | M64;
      const uint32_t id_pk = (idesc_bf16_m128(64u, false, true) & ~M128) | M64;
      const uint32_t id_ptk = (idesc_bf16_m128(64u, true, true) & ~M128) | M64;
      int s = 0;
      uint32_t ph = 0, nacc = 0;
      auto acquire = [&]() -> uint32_t {
        mbar_wait(full_bar(s), ph, 2);
        return sRing + (uint32_t)s * TILE_B;
      };
      auto advance = [&]() { if (++s == NS) { s = 0; ph ^= 1u; } };
      // per sub-tile g: acc_g (64 x 64) (+)= A_tile[g] (64 x 64) * B_tile[g]^T, two consecutive ring stages
      auto score_product = [&](uint32_t col, bool first) {
        const uint32_t a = acquire();
        const int sa = s;
        advance();
        const uint32_t b = acquire();
        tcgen05_fence_after();
#pragma unroll
        for (uint32_t g = 0; g < 2; ++g) {
          const uint64_t da = desc_kmajor(a + g * SUB_B), db = desc_kmajor(b + g * SUB_B);
          const uint32_t d = tmem_base + col + ((16u * g) << 16);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d, da + 2u * k, db + 2u * k, id_s, (!first || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(sa));
        umma_commit(empty_bar(s));
        advance();
      };
      // per sub-tile g: slot_g (64 x 64) = A_g (P or dS, [64 x 64], K- or MN-major) * ring tile rows of g (MN-major B)
      auto chunk_product = [&](uint32_t a_base, bool a_mn) {
        const uint32_t b = acquire();
        const uint32_t slot = nacc % NACC;
        mbar_wait(acc_empty(slot), ((nacc / NACC) & 1u) ^ 1u, 3);
        tcgen05_fence_after();
#pragma unroll
        for (uint32_t g = 0; g < 2; ++g) {
          const uint32_t d = tmem_base + acc_col(slot) + ((16u * g) << 16);
          const uint64_t db = desc_mnmajor(b + g * SUB_B, 8192u);
          const uint64_t da = a_mn ? desc_mnmajor(a_base + g * SUB_B, 8192u) : desc_kmajor(a_base + g * SUB_B);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            umma_bf16(d, da + (a_mn ? 128u : 2u) * j, db + 128u * j, a_mn ? id_ptk : id_pk, j > 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(s));
        advance();
        umma_commit(acc_full(slot));
        ++nacc;
      };
      // score products of chunk c of tile `it` (buffer it & 1); the last chunk publishes the buffer to the softmax warps
      auto phase_a = [&](int it, int c) {
        const uint32_t b = (uint32_t)it & 1u;
        score_product(64u * b, c == 0);
        if (BWD) score_product(DP_COL0 + 64u * b, c == 0);
        if (c == NC - 1) umma_commit(s_full(b));
      };
      for (int it = 0; it < 2 && it < n_my; ++it)
        for (int c = 0; c < NC; ++c) phase_a(it, c);
      for (int it = 0; it < n_my; ++it) {
        mbar_wait(p_full, (uint32_t)it & 1u, 4);  // also: the softmax has finished reading S / dP buffer it & 1
        tcgen05_fence_after();
        for (int c = 0; c < NC; ++c) {
          if (BWD) {
            chunk_product(sDS, false);  // dQ_c = dS K_c
            chunk_product(sDS, true);   // dK_c = dS^T Q_c
            chunk_product(sP, true);    // dV_c = P^T dO_c
          } else {
            chunk_product(sP, false);   // O_c = P V_c
          }
          if (c == NC - 1) umma_commit(p_empty);  // P / dS may be overwritten once these products retire
          if (it + 2 < n_my) phase_a(it + 2, c);  // HBM-bound score operands interleave with the L2-resident re-reads
        }
      }
    }
  } else if (warp < 4) {
    // ===================== softmax warps =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 240;" ::: "memory");
    const int q = warp & 3;
    const int g = lane >> 4;                  // sub-tile
    const int m = 16 * q + (lane & 15);       // row inside the sub-tile
    const int g2 = m / LP, i = m - g2 * LP;   // window inside the sub-tile, token
    const int wt = g * (64 / LP) + g2;        // window inside the tile
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t cb = (uint32_t)(g2 * LP);  // first column of this row's diagonal block (warp-uniform)
    const uint32_t sPg = sP + (uint32_t)g * SUB_B, sDSg = sDS + (uint32_t)g * SUB_B;
    const float sl2 = p.scale * LOG2E;
    const bool drop = p.drop_p > 0.f;
    const bool want_db = BWD && p.dbias != nullptr;
    const float* brow = bias_s + (i < L ? i : L - 1) * tp.bias_pitch;  // bias * log2e ; -inf in the columns >= L
    const int64_t ld8 = (L + 7) >> 3;
    const uint32_t thr_hi = p.drop_thr16 << 16;
    const int c0 = (int)(cb >> 3);            // first 16-byte chunk of the diagonal block in this row of P / dS
    if (want_db) {
      uint32_t z[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) z[j] = 0u;
#pragma unroll
      for (int k = 0; k < MAXP; ++k) tmem_st_32x32b_x32(trow + DB_COL0 + 32u * k, z);
      tmem_st_wait();
    }
    for (int it = 0; it < n_my; ++it) {
      const int t = (int)blockIdx.x + it * (int)gridDim.x;
      const uint32_t b = (uint32_t)it & 1u;
      const int64_t w = (int64_t)t * G + wt;
      const bool valid = (i < L) && (w < p.W);
      const int64_t grow = (w * p.H + h) * (int64_t)L + i;
      mbar_wait(s_full(b), ((uint32_t)it >> 1) & 1u, 5);
      tcgen05_fence_after();
      // ---- the whole score row in registers: e = exp2((s scale + bias) log2e - max) ; columns >= L come out as 0 ----
      float e[32 * MAXP];
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < MAXP; ++k) {
        uint32_t rs[32];
        tmem_ld_32x32b_x32(trow + 64u * b + cb + 32u * k, rs);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = fmaf(__uint_as_float(rs[j]), sl2, brow[32 * k + j]);
          e[32 * k + j] = x;
          mx = fmaxf(mx, x);
        }
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 32 * MAXP; ++j) {
        e[j] = ex2_approx(e[j] - mx);
        sum += e[j];
      }
      // rows of padding (token >= L, window >= W) produce all-zero P / dS rows through this one factor
      const float inv = valid ? 1.0f / sum : 0.f;
      const float ic = inv * p.drop_scale;
      // post-dropout probabilities as packed bf16 (the operand of P V / P^T dO); backward: t = P_dropped * dP
      uint32_t pk[16 * MAXP];
      float tt[BWD ? 32 * MAXP : 1];
      float delta = 0.f;
#pragma unroll
      for (int k = 0; k < MAXP; ++k) {
        uint32_t rd[32];
        if (BWD) {
          tmem_ld_32x32b_x32(trow + DP_COL0 + 64u * b + cb + 32u * k, rd);
          tmem_ld_wait();
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int n = 4 * k + u;
          uint32_t rnd[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
          if (drop && n * 8 < L) dropout_rand8(p.seed, p.offset, (uint64_t)(grow * ld8 + n), rnd);
          float pd[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const uint32_t word = rnd[jj >> 1];
            const bool kept = (jj & 1) ? (word >= thr_hi) : ((word << 16) >= thr_hi);
            pd[jj] = kept ? e[32 * k + 8 * u + jj] * ic : 0.f;
            if (BWD) {
              const float x = pd[jj] * __uint_as_float(rd[8 * u + jj]);
              tt[BWD ? 32 * k + 8 * u + jj : 0] = x;
              delta += x;
            }
          }
          if (!BWD && p.probs != nullptr && valid) {
            float* prow = p.probs + grow * (int64_t)L;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
              if (8 * n + jj < L) prow[8 * n + jj] = pd[jj];
          }
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) pk[16 * k + 4 * u + jj] = pack_bf16x2(pd[2 * jj], pd[2 * jj + 1]);
        }
      }
      // ---- backward: scale * dS = scale (t - p delta) as packed bf16 ; bias gradient (kept scaled, unscaled at the
      // flush).  Everything is computed BEFORE waiting for the previous tile's products to release P / dS, so that only
      // the 16-byte stores remain on the MMA warp's critical path ----
      uint32_t dk[BWD ? 16 * MAXP : 1];
      if (BWD) {
        const float nk = -inv * delta * p.scale;
#pragma unroll
        for (int k = 0; k < MAXP; ++k) {
          uint32_t rb[32];
          if (want_db) {
            tmem_ld_32x32b_x32(trow + DB_COL0 + 32u * k, rb);
            tmem_ld_wait();
          }
#pragma unroll
          for (int j2 = 0; j2 < 16; ++j2) {
            const int j = 32 * k + 2 * j2;
            const float d0 = fmaf(e[j], nk, tt[BWD ? j : 0] * p.scale);
            const float d1 = fmaf(e[j + 1], nk, tt[BWD ? j + 1 : 0] * p.scale);
            if (want_db) {
              rb[2 * j2] = __float_as_uint(__uint_as_float(rb[2 * j2]) + d0);
              rb[2 * j2 + 1] = __float_as_uint(__uint_as_float(rb[2 * j2 + 1]) + d1);
            }
            dk[BWD ? 16 * k + j2 : 0] = pack_bf16x2(d0, d1);
          }
          if (want_db) tmem_st_32x32b_x32(trow + DB_COL0 + 32u * k, rb);
        }
        if (want_db) tmem_st_wait();
      }
      if (it > 0) mbar_wait(p_empty, ((uint32_t)it - 1u) & 1u, 7);  // the products of the previous tile have read P (dS)
#pragma unroll
      for (int c = 0; c < 4 * MAXP; ++c) {
        st_shared_v4(sPg + sw128(m, c0 + c), make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]));
        if (BWD)
          st_shared_v4(sDSg + sw128(m, c0 + c), make_uint4(dk[BWD ? 4 * c : 0], dk[BWD ? 4 * c + 1 : 0],
                                                           dk[BWD ? 4 * c + 2 : 0], dk[BWD ? 4 * c + 3 : 0]));
      }
      fence_proxy_async();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    if (want_db) {
      const float unscale = 1.0f / p.scale;
#pragma unroll
      for (int k = 0; k < MAXP; ++k) {
        uint32_t rb[32];
        tmem_ld_32x32b_x32(trow + DB_COL0 + 32u * k, rb);
        tmem_ld_wait();
        if (i >= 1 && i < L) {
          float* drow = p.dbias + ((int64_t)h * L + i) * L;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = 32 * k + j;
            if (col >= 1 && col < L) atomicAdd(drow + col, __uint_as_float(rb[j]) * unscale);
          }
        }
      }
    }
  } else if (warp < 8) {
    // ===================== epilogue warps =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 128;" ::: "memory");
    const int q = warp & 3;
    const int r = (lane >> 4) * 64 + 16 * q + (lane & 15);  // row of the 128-row staging tile
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool leader = threadIdx.x == 4 * 32;
    uint32_t n = 0;
    for (int it = 0; it < n_my; ++it) {
      const int w0 = ((int)blockIdx.x + it * (int)gridDim.x) * G;
#pragma unroll 1
      for (int prod = 0; prod < NPROD; ++prod, ++n) {
        const uint32_t slot = n % NACC;
        mbar_wait(acc_full(slot), (n / NACC) & 1u, 6);
        tcgen05_fence_after();
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(trow + acc_col(slot), r0);
        tmem_ld_32x32b_x32(trow + acc_col(slot) + 32u, r1);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(slot));
        const uint32_t stg = sStg + (n & 1u) * TILE_B;
        if (leader) tma_wait_group_read<1>();
        named_bar_sync(1, 128);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(r0[8 * c + 0]), __uint_as_float(r0[8 * c + 1]));
          u.y = pack_bf16x2(__uint_as_float(r0[8 * c + 2]), __uint_as_float(r0[8 * c + 3]));
          u.z = pack_bf16x2(__uint_as_float(r0[8 * c + 4]), __uint_as_float(r0[8 * c + 5]));
          u.w = pack_bf16x2(__uint_as_float(r0[8 * c + 6]), __uint_as_float(r0[8 * c + 7]));
          st_shared_v4(stg + sw128(r, c), u);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(r1[8 * c + 0]), __uint_as_float(r1[8 * c + 1]));
          u.y = pack_bf16x2(__uint_as_float(r1[8 * c + 2]), __uint_as_float(r1[8 * c + 3]));
          u.z = pack_bf16x2(__uint_as_float(r1[8 * c + 4]), __uint_as_float(r1[8 * c + 5]));
          u.w = pack_bf16x2(__uint_as_float(r1[8 * c + 6]), __uint_as_float(r1[8 * c + 7]));
          st_shared_v4(stg + sw128(r, 4 + c), u);
        }
        fence_proxy_async();
        named_bar_sync(1, 128);
        if (leader) {
          const int c = BWD ? prod / 3 : prod;
          const int which = BWD ? prod - 3 * c : 0;
          tma_store_3d_hint(&tm_out, stg, which * HD + h * DK + 64 * c, 0, w0, L2_EVICT_FIRST);
          tma_commit_group();
        }
      }
    }
    if (leader) tma_wait_group<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 9) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// 3-D bf16 map over [W][L][cols] (row pitch ld elements, window pitch L * ld), box [G][LP][64], SWIZZLE_128B; loads
// zero-fill rows L..LP-1 and windows >= W, stores skip them
static int make_tmap3d(CUtensorMap* tm, const void* ptr, int64_t cols, int64_t L, int64_t W, int64_t ld, int LP, int G) {
  const uint64_t gdim[3] = {(uint64_t)cols, (uint64_t)L, (uint64_t)W};
  const uint64_t gstride[2] = {(uint64_t)ld * 2, (uint64_t)L * (uint64_t)ld * 2};
  const uint32_t box[3] = {64u, (uint32_t)LP, (uint32_t)G};
  return encode_tmap_bf16_sw128(tm, ptr, 3, gdim, gstride, box, 128);
}

typedef void (*KernelFn)(const TcParams, const CUtensorMap, const CUtensorMap, const CUtensorMap);
template <int LP, int DK, bool BWD>
static KernelFn kernel_for() {
  if constexpr (LP == 128) return attn_tc_kernel<128, DK, BWD>;
  else return attn_tc64_kernel<LP, DK, BWD>;
}

template <int LP, int DK, bool BWD>
static int launch(const Params& p, cudaStream_t stream) {
  constexpr int G = 128 / LP;
  TcParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.p = p;
  tp.tiles = (int)((p.W + G - 1) / G);
  tp.nkeys = (G > 1) ? 128 : ((p.L + 15) / 16) * 16;
  tp.ks_tok = (G > 1) ? 8 : tp.nkeys / 16;
  tp.npiece = (G > 1) ? LP / 32 : (tp.nkeys + 31) / 32;
  if (G > 1) {  // sub-tile kernel: always staged, LP columns per row (padding columns hold -inf)
    tp.bias_pitch = LP + 1;
    tp.bias_bytes = (uint32_t)(((size_t)p.L * tp.bias_pitch * 4 + 15) / 16 * 16);
  } else {
    tp.bias_pitch = p.L | 1;
    tp.bias_bytes = p.bias != nullptr ? (uint32_t)(((size_t)p.L * tp.bias_pitch * 4 + 15) / 16 * 16) : 0u;
  }
  constexpr uint32_t OPERAND_B = (G > 1) ? 8192u : PANEL_B;  // P (and dS): 2 x [64 x 64] or 2 panels of [128 x 64]
  const uint32_t fixed = 1024u + (BWD ? 4u : 2u) * OPERAND_B + 2u * TILE_B + tp.bias_bytes + BAR_BYTES;
  int ns = (int)((SMEM_LIMIT - fixed) / TILE_B);
  if (ns > MAX_STAGES) ns = MAX_STAGES;
  {
    static const int cap = [] {  // LSTC_ATTN_STAGES=<n>: cap the ring depth (experiments on latency hiding)
      const char* e = getenv("LSTC_ATTN_STAGES");
      return e != nullptr ? atoi(e) : 0;
    }();
    if (cap >= 3 && ns > cap) ns = cap;
  }
  if (ns < 3) {
    set_last_error("attention: no shared memory left for the operand ring (L=%d)", p.L);
    return LSTC_ERR_UNSUPPORTED;
  }
  tp.n_stages = ns;
  const uint32_t smem = fixed + (uint32_t)ns * TILE_B;
  const int HD = p.H * DK;
  CUtensorMap tq, td, to;
  memset(&td, 0, sizeof(td));
  int rc = make_tmap3d(&tq, p.qkv, 3 * (int64_t)HD, p.L, p.W, p.ld, LP, G);
  if (rc != LSTC_OK) return rc;
  if (BWD) {
    rc = make_tmap3d(&td, p.dout, HD, p.L, p.W, p.ld_dout, LP, G);
    if (rc != LSTC_OK) return rc;
  }
  rc = make_tmap3d(&to, p.out, (BWD ? 3 : 1) * (int64_t)HD, p.L, p.W, p.ld_out, LP, G);
  if (rc != LSTC_OK) return rc;
  auto kern = kernel_for<LP, DK, BWD>();
  static bool attr_set[64] = {false};  // per instantiation and device; idempotent
  int dev = 0;
  LSTC_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    LSTC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  int gx = num_sms() / p.H;
  if (gx < 1) gx = 1;
  if (gx > tp.tiles) gx = tp.tiles;
  kern<<<dim3((unsigned)gx, (unsigned)p.H), LP == 128 ? NUM_THREADS : NUM_THREADS64, smem, stream>>>(tp, tq, td, to);
  LSTC_CHECK_LAUNCH();
  return LSTC_OK;
}

template <int DK>
static int dispatch_lp(bool bwd, const Params& p, cudaStream_t stream) {
  const int L = p.L;
  if (L <= 32) return bwd ? launch<32, DK, true>(p, stream) : launch<32, DK, false>(p, stream);
  if (L <= 64) return bwd ? launch<64, DK, true>(p, stream) : launch<64, DK, false>(p, stream);
  if (L <= 96) return bwd ? launch<128, DK, true>(p, stream) : launch<128, DK, false>(p, stream);
  set_last_error("attention: L=%d exceeds the supported maximum of 96 tokens per window", L);
  return LSTC_ERR_UNSUPPORTED;
}

int run(bool bwd, const Params& p, int dk, cudaStream_t stream) {
  switch (dk) {
    case 64: return dispatch_lp<64>(bwd, p, stream);
    case 128: return dispatch_lp<128>(bwd, p, stream);
    case 256: return dispatch_lp<256>(bwd, p, stream);
    default:
      set_last_error("attention: d_k=%d unsupported (64, 128 or 256)", dk);
      return LSTC_ERR_UNSUPPORTED;
  }
}

}  // namespace attn_tc
}  // namespace lstc

LSTC_DEFINE_RNG_STEP_SETTER(set_rng_step_attention_tc)
