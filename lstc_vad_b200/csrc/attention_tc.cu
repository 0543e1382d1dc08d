// Fused short-sequence multi-head attention (forward + backward) on the sm_100a tensor cores: tcgen05.mma with the
// score / gradient accumulators in TMEM, operands streamed by TMA, results written back by TMA stores.
//
// Replaces models/MultiHeadAttention.py:100-122 of the reference:
//   q,k,v -> [W,H,L,dk] ; S = (q/sqrt(dk)) k^T ; S[:,:,1:,1:] += bias ; P = dropout(softmax(S)) ;
//   O = P v ; O.transpose(1,2).contiguous().view(W, L, H*dv)
// and its autograd backward.
//
// A window has L <= 96 tokens, far below the 128 rows of a tcgen05 tile.  Two kernels (see the comment above each):
//   attn_tc64_kernel  (L <= 64): a TILE stacks G = 128 / LP windows of one head (LP = 32 or 64 rows per window, rows
//                     >= L zero-filled by TMA); M = 64 products per 64-row sub-tile, both sub-tiles share TMEM columns
//   attn_tc128_kernel (65..96) : one window per tile, M = 128 products
// Common design: one [G][LP][64] TMA box of q / k / v / dO is a SWIZZLE_128B operand tile; probabilities go back to
// shared memory as bf16 operands (block-diagonal when several windows share a product), so that
//   forward : S = sum_c Q_c K_c^T (c = 64-column chunks of dk)  -> softmax / dropout -> O_c = P V_c
//   backward: S = sum_c Q_c K_c^T ; dP = sum_c dO_c V_c^T       -> P, dS            -> dQ_c = dS K_c ;
//             dK_c = dS^T Q_c ; dV_c = P^T dO_c                     (K, Q, dO chunks re-read through L2, kept there
//                                                                    by evict-last / evict-first TMA cache hints)
// The same shared-memory tile serves as K-major operand (row = token: Q K^T, dO V^T), as MN-major B operand (token =
// reduction index: P V, dS K, ...) and P / dS as K-major or MN-major A operand (dS vs dS^T) purely through the UMMA
// descriptors - nothing is ever transposed in memory.
//
// One persistent CTA per SM (grid = (CTAs per head, heads)), warp-specialised, no CTA-wide barrier after start-up:
//   TMA producer   [rows x 64] tiles into an n-stage ring (full / empty mbarriers), runs ahead across tiles, in exactly
//                  the order the MMA warp consumes them
//   MMA issuer     one thread; S (and dP) are double-buffered in TMEM so the score products of later tiles are issued
//                  before the softmax of the current one has finished; owns the 512 TMEM columns
//   softmax warps  thread = one score row (tcgen05.ld 32x32b: no shuffles), Philox dropout, dS = P (dP - sum P dP);
//                  packed bf16 P / dS into the smem operands; the rel-pos bias gradient is accumulated in TMEM columns
//                  over all tiles of the CTA and flushed once with atomics
//   epilogue warps accumulator chunk (TMEM) -> bf16 -> swizzled smem staging tile -> TMA store (the tensor map clips
//                  rows >= L and windows >= W)
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

constexpr int NUM_THREADS64 = 384;   // 12 warps = 3 warpgroups (setmaxnreg re-allocates registers per warpgroup)
constexpr int NUM_THREADS128 = 512;  // 16 warps = 4 warpgroups (second epilogue group)
constexpr uint32_t TILE_B = 16384;  // [128 rows][64 bf16]
constexpr uint32_t PANEL_B = 16384; // one 64-column panel of the [128 x 128] P / dS operand
constexpr int MAX_STAGES = 16;
constexpr uint32_t SMEM_LIMIT = 232448;  // 227 KB opt-in maximum per CTA
constexpr uint32_t BAR_BYTES = 1024;

// Optional pipeline accounting (build with -DLSTC_ATTN_TIMING): CTA (0,0) prints, per role, the cycles it spent waiting
// on each barrier class - the quickest way to see which stage of the producer / MMA / softmax / epilogue pipeline is the
// critical one.  Off in normal builds (the macros compile to nothing).
#ifdef LSTC_ATTN_TIMING
#define TWAIT(acc, stmt)               \
  do {                                 \
    const long long _t0 = clock64();   \
    stmt;                              \
    (acc) += clock64() - _t0;          \
  } while (0)
#define TDECL(...) long long __VA_ARGS__
#define TPRINT(...)                                  \
  do {                                               \
    if (blockIdx.x == 0 && blockIdx.y == 0) printf(__VA_ARGS__); \
  } while (0)
#else
#define TWAIT(acc, stmt) stmt
#define TDECL(...)
#define TPRINT(...)
#endif

struct TcParams {
  Params p;
  int tiles;       // tiles per head = ceil(W / G)
  int nkeys;       // N of the score product (multiple of 16): 128 when G > 1, round_up(L, 16) when G == 1
  int ks_tok;      // UMMA_K steps over the token (reduction) dim of P V, dS K, dS^T Q, P^T dO
  int npiece;      // 32-column pieces of one score row
  int n_stages;    // total ring depth = n_stages_a + n_stages_b
  int n_stages_a;  // stages of the score-operand ring (q, k, dO, v of the S / dP products: HBM-bound)
  int n_stages_b;  // stages of the chunk-operand ring (forward: v ; backward: k, q, dO re-read through L2)
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

// writes 32 consecutive bf16 values (16 packed words) of row r starting at column cb of a two-panel operand
__device__ __forceinline__ void store_piece_packed(uint32_t base, int r, int cb, const uint32_t* w) {
  const uint32_t pb = base + (uint32_t)(cb >> 6) * PANEL_B;
  const int c0 = (cb & 63) >> 3;
#pragma unroll
  for (int c = 0; c < 4; ++c) st_shared_v4(pb + sw128(r, c0 + c), make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]));
}

// ==========================================================================================================
// Windows of 65..96 tokens: one window per tile, M = 128 products (an M = 64 instruction costs the same tensor-pipe
// time, so two sub-tile products per window would double it), software-pipelined like the sub-tile kernel below:
//   forward  MMA order: S(0), S(1) ; for t: [wait P(t)], { P V_c(t), S_c(t+2) } per chunk c
//   backward MMA order: S(0), dP(0), S(1) ; for t: [wait P(t)], dP(t+1), { dQ_c dK_c dV_c (t), S_c(t+2) } per chunk c
// S is double-buffered in TMEM (96 columns each); dP is not (there is no room), so dP(t+1) is issued first thing after
// the softmax of tile t has released it.  A score row has up to 96 columns - too many to keep e, dP and both packed
// operands in registers - so the softmax works in 32-column pieces and uses the S / dP columns themselves as scratch:
// e = exp2(x - max) is written back over S, t = P_dropped * dP over dP.
// Ring stages hold only nkeys = round_up(L, 16) rows (the 128-row A operand of a score product then reads the next
// stage's rows as its rows nkeys..127: finite garbage that only reaches accumulator rows nobody stores); rows of
// padding are forced to exact zeros in P / dS (they are reduction indices of dK = dS^T Q and dV = P^T dO).
// TMEM columns: S[2] 0, 96 ; forward: accumulator slots 192 + 64 a (a < 4)
//               backward: dP 192 ; slots 288, 352 ; bias gradient 416..511
// Warps: 0-3 softmax (thread = score row = TMEM lane), 4-7 epilogue, 8 TMA producer, 9 MMA issuer, 10-11 idle.
// ==========================================================================================================
template <int DK, bool BWD>
__global__ void __launch_bounds__(NUM_THREADS128, 1)
attn_tc128_kernel(const TcParams tp, const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                  const __grid_constant__ CUtensorMap tm_out) {
  constexpr int NC = DK / 64;
  constexpr int NACC = BWD ? 2 : 4;
  constexpr int NPIECE = 3;                 // 32-column pieces of a score row (96 columns)
  constexpr int NPROD = BWD ? 3 * NC : NC;
  constexpr uint32_t S_N = 96;              // N of the score products (keys nkeys..95: finite garbage, masked by the bias)
  constexpr uint32_t DP_COL0 = 192u, DB_COL0 = 416u;
  constexpr int BIAS_PITCH_W = 49;          // words (bf16 pairs) per bias row: odd -> conflict-free column reads
  constexpr float LOG2E = 1.4426950408889634f;
  auto acc_col = [](uint32_t a) -> uint32_t { return (BWD ? 288u : 192u) + 64u * a; };

  Params p = tp.p;
  p.offset += rng_step();
  const int NS = tp.n_stages, NSA = tp.n_stages_a, NSB = tp.n_stages_b;
  const int L = p.L;
  const int h = blockIdx.y;
  const int HD = p.H * DK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t STAGE_B = (uint32_t)tp.nkeys * 128u;
  const int KS = tp.ks_tok;

  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s0 = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t sRing = s0;
  const uint32_t sP = sRing + (uint32_t)NS * STAGE_B;
  const uint32_t sDS = sP + 2 * PANEL_B;
  const uint32_t sStg = BWD ? sDS + 2 * PANEL_B : sDS;
  const uint32_t sBias = sStg + 2 * TILE_B;
  const uint32_t sBar = sBias + tp.bias_bytes;
  // two operand rings with their own producer and issuer thread each: A = operands of the score products, B = operands of
  // the chunk products; ring B's stages follow ring A's in shared memory
  const uint32_t sRingB = sRing + (uint32_t)NSA * STAGE_B;
  auto full_a = [&](int s) { return sBar + 8u * s; };
  auto empty_a = [&](int s) { return sBar + 8u * (MAX_STAGES + s); };
  auto full_b = [&](int s) { return sBar + 8u * (2 * MAX_STAGES + s); };
  auto empty_b = [&](int s) { return sBar + 8u * (3 * MAX_STAGES + s); };
  auto s_full = [&](uint32_t b) { return sBar + 8u * (4 * MAX_STAGES + b); };
  const uint32_t p_full = sBar + 8u * (4 * MAX_STAGES + 2), p_empty = p_full + 8u, dp_full = p_full + 16u;
  auto acc_full = [&](int a) { return sBar + 8u * (4 * MAX_STAGES + 5 + a); };
  auto acc_empty = [&](int a) { return sBar + 8u * (4 * MAX_STAGES + 9 + a); };
  const uint32_t tmem_slot = sBar + 8u * (4 * MAX_STAGES + 13);
  uint32_t* bias_w = reinterpret_cast<uint32_t*>(smem + (sBias - smem_u32(smem)));

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_out);
    if (BWD) tma_prefetch_desc(&tm_do);
    for (int s = 0; s < NSA; ++s) {
      mbar_init(full_a(s), 1);
      mbar_init(empty_a(s), 1);
    }
    for (int s = 0; s < NSB; ++s) {
      mbar_init(full_b(s), 1);
      mbar_init(empty_b(s), 1);
    }
    mbar_init(s_full(0), 1);
    mbar_init(s_full(1), 1);
    mbar_init(p_full, 4);
    mbar_init(p_empty, 2);  // one commit per chunk-product issuer
    mbar_init(dp_full, 1);
    for (int a = 0; a < NACC; ++a) {
      mbar_init(acc_full(a), 1);
      mbar_init(acc_empty(a), 4);  // the 4 warps of the epilogue group that drains this slot (slot parity = group)
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  {
    // ring + block-diagonal operands start all zero: rows a product reads beyond a stage's nkeys rows (the next stage,
    // or the P buffer after the last one) are then always finite values, never stale NaN bit patterns
    const uint32_t nbytes = (uint32_t)NS * STAGE_B + (BWD ? 4u : 2u) * PANEL_B;  // NS = NSA + NSB: both rings
    for (uint32_t o = threadIdx.x * 16u; o < nbytes; o += NUM_THREADS128 * 16u) st_shared_v4(sRing + o, make_uint4(0, 0, 0, 0));
    // bias rows as bf16 pairs: bias * log2e (0 without a bias) in the columns < L, -inf in the padding columns
    const float* bsrc = p.bias != nullptr ? p.bias + (int64_t)h * L * L : nullptr;
    for (int idx = threadIdx.x; idx < L * 48; idx += NUM_THREADS128) {
      const int r = idx / 48, j = idx - r * 48;
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = 2 * j + e;
        v[e] = c < L ? (bsrc != nullptr ? __ldg(bsrc + r * L + c) * LOG2E : 0.f) : -INFINITY;
      }
      bias_w[r * BIAS_PITCH_W + j] = pack_bf16x2(v[0], v[1]);
    }
  }
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  const int n_my = ((int)blockIdx.x < tp.tiles) ? (tp.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp >= 8 && warp < 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;" ::: "memory");
  }
  if (warp == 8) {
    // ===================== TMA producer: ONE thread feeds both rings, each in its consumers' order; it polls the two
    // "stage free" barriers so that a full ring never holds up the other one =====================
    if (lane == 0) {
      const uint64_t pol_again = BWD ? L2_EVICT_LAST : L2_EVICT_FIRST, pol_once = L2_EVICT_FIRST;
      // backward: q, k, dO are read again by the chunk products of the same tile -> keep them in L2 until then;
      // everything that is read for the last time leaves L2 first
      const int cq = h * DK, ck = HD + h * DK, cv = 2 * HD + h * DK;
      const int na = n_my * (BWD ? 4 * NC : 2 * NC), nb = n_my * (BWD ? 3 * NC : NC);
      int ia = 0, ib = 0, sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      TDECL(t_begin = clock64());
      while (ia < na || ib < nb) {
        if (ia < na && mbar_try_wait(empty_a(sa), pha ^ 1u)) {
          const int per = BWD ? 4 * NC : 2 * NC;          // stages of ring A per tile: S(t) then dP(t)
          const int it = ia / per, j = ia - it * per;
          const bool dp = j >= 2 * NC;
          const int c = (dp ? j - 2 * NC : j) >> 1, second = j & 1;
          const CUtensorMap* tm = (dp && !second) ? &tm_do : &tm_qkv;
          const int col = dp ? (second ? cv : cq) : (second ? ck : cq);
          const uint64_t pol = (dp && second) ? pol_once : pol_again;
          mbar_arrive_expect_tx(full_a(sa), STAGE_B);
          tma_load_3d_hint(sRing + (uint32_t)sa * STAGE_B, tm, full_a(sa), col + 64 * c, 0, ((int)blockIdx.x + it * (int)gridDim.x), pol);
          ++ia;
          if (++sa == NSA) { sa = 0; pha ^= 1u; }
        }
        if (ib < nb && mbar_try_wait(empty_b(sb), phb ^ 1u)) {
          const int per = BWD ? 3 * NC : NC;              // stages of ring B per tile: per chunk k, q, dO (or v)
          const int it = ib / per, j = ib - it * per;
          const int c = BWD ? j / 3 : j, which = BWD ? j - 3 * c : 0;
          const CUtensorMap* tm = (BWD && which == 2) ? &tm_do : &tm_qkv;
          const int col = BWD ? (which == 0 ? ck : cq) : cv;
          mbar_arrive_expect_tx(full_b(sb), STAGE_B);
          tma_load_3d_hint(sRingB + (uint32_t)sb * STAGE_B, tm, full_b(sb), col + 64 * c, 0, ((int)blockIdx.x + it * (int)gridDim.x), pol_once);
          ++ib;
          if (++sb == NSB) { sb = 0; phb ^= 1u; }
        }
      }
      TPRINT("producer: total %lld cyc\n", clock64() - t_begin);
    }
  } else if (warp >= 9 && warp <= 11) {
    // ===================== MMA issuers =====================
    // Two issuing threads with their own operand ring each: warp 9 issues the score products (S, dP), warp 10 the chunk
    // products (P V, or dQ / dK / dV).  Every product is a handful of small MMAs behind two barrier waits and two
    // commits - ~1 us of single-thread issue work for ~0.1 us of tensor-pipe time - so ONE issuer was the critical
    // path of the whole kernel (17.5 k of 22 k cycles per tile at L = 81).
    if (lane == 0) {
      const bool do_a = warp == 9;
      const int n_s = do_a ? NSA : NSB;
      const uint32_t base = do_a ? sRing : sRingB;
      const uint32_t id_s = idesc_bf16_m128(S_N, false, false);
      const uint32_t id_pk = idesc_bf16_m128(64u, false, true);
      const uint32_t id_ptk = idesc_bf16_m128(64u, true, true);
      int s = 0;
      uint32_t ph = 0;
      TDECL(w_full = 0, w_acc = 0, w_p = 0, t_begin = clock64());
      auto acquire = [&]() -> uint32_t {
        TWAIT(w_full, mbar_wait(do_a ? full_a(s) : full_b(s), ph, 2));
        return base + (uint32_t)s * STAGE_B;
      };
      auto advance = [&]() { if (++s == n_s) { s = 0; ph ^= 1u; } };
      auto score_product = [&](uint32_t col, bool first) {
        const uint32_t a = acquire();
        const int sa = s;
        advance();
        const uint32_t b = acquire();
        tcgen05_fence_after();
        const uint64_t da = desc_kmajor(a), db = desc_kmajor(b);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + col, da + 2u * k, db + 2u * k, id_s, (!first || k > 0) ? 1u : 0u);
        umma_commit(empty_a(sa));
        umma_commit(empty_a(s));
        advance();
      };
      auto chunk_product = [&](uint32_t n, uint32_t me, uint32_t a_base, bool a_mn) {
        if ((n & 1u) != me) return;
        const uint32_t sb = n % (uint32_t)NSB, slot = n % NACC;
        TWAIT(w_full, mbar_wait(full_b(sb), (n / (uint32_t)NSB) & 1u, 2));
        const uint32_t b = sRingB + sb * STAGE_B;
        TWAIT(w_acc, mbar_wait(acc_empty(slot), ((n / NACC) & 1u) ^ 1u, 3));
        tcgen05_fence_after();
        const uint32_t d = tmem_base + acc_col(slot);
        const uint64_t db = desc_mnmajor(b, 8192u);
        if (a_mn) {
          const uint64_t da = desc_mnmajor(a_base, PANEL_B);
          for (int j = 0; j < KS; ++j) umma_bf16(d, da + 128u * j, db + 128u * j, id_ptk, j > 0 ? 1u : 0u);
        } else {
          for (int j = 0; j < KS; ++j)
            umma_bf16(d, desc_kmajor(a_base + (uint32_t)(j >> 2) * PANEL_B) + 2u * (j & 3), db + 128u * j, id_pk,
                      j > 0 ? 1u : 0u);
        }
        umma_commit(empty_b(sb));
        umma_commit(acc_full(slot));
      };
      if (do_a) {
        auto phase_s = [&](int it) {
          for (int c = 0; c < NC; ++c) score_product(S_N * ((uint32_t)it & 1u), c == 0);
          umma_commit(s_full((uint32_t)it & 1u));
        };
        auto phase_dp = [&](int it) {
          for (int c = 0; c < NC; ++c) score_product(DP_COL0, c == 0);
          umma_commit(dp_full);
        };
        if (n_my > 0) {
          phase_s(0);
          if (BWD) phase_dp(0);
        }
        if (n_my > 1) phase_s(1);
        for (int it = 0; it < n_my; ++it) {
          // softmax of tile `it` done: S[it & 1] and dP are released
          TWAIT(w_p, mbar_wait(p_full, (uint32_t)it & 1u, 4));
          tcgen05_fence_after();
          if (BWD && it + 1 < n_my) phase_dp(it + 1);
          if (it + 2 < n_my) phase_s(it + 2);
        }
      } else {
        // chunk products n = 0, 1, 2, ... of this CTA in order; issuer j takes those with n % 2 == j.  Product n reads
        // stage n % NSB of ring B (one stage per product) and writes accumulator slot n % NACC.
        const uint32_t me = (uint32_t)(warp - 10);
        uint32_t n = 0;
        for (int it = 0; it < n_my; ++it) {
          TWAIT(w_p, mbar_wait(p_full, (uint32_t)it & 1u, 4));  // P / dS of tile `it` written
          tcgen05_fence_after();
          for (int c = 0; c < NC; ++c) {
            if (BWD) {
              chunk_product(n++, me, sDS, false);  // dQ_c = dS K_c
              chunk_product(n++, me, sDS, true);   // dK_c = dS^T Q_c
              chunk_product(n++, me, sP, true);    // dV_c = P^T dO_c
            } else {
              chunk_product(n++, me, sP, false);   // O_c = P V_c
            }
          }
          umma_commit(p_empty);  // (one arrival per issuer) P / dS may be overwritten once these products retire
        }
      }
      TPRINT("mma%d    : total %lld cyc, waiting for operands %lld, for an accumulator slot %lld, for the softmax %lld (%d tiles)\n",
             warp - 9, clock64() - t_begin, w_full, w_acc, w_p, n_my);
    }
  } else if (warp < 4) {
    // ===================== softmax warps: thread = score row =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 240;" ::: "memory");
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const float sl2 = p.scale * LOG2E;
    const bool drop = p.drop_p > 0.f;
    const bool want_db = BWD && p.dbias != nullptr;
    const bool valid = r < L;
    const uint32_t* brow = bias_w + (valid ? r : L - 1) * BIAS_PITCH_W;
    const int64_t ld8 = (L + 7) >> 3;
    const uint32_t thr_hi = p.drop_thr16 << 16;
    if (want_db) {
      uint32_t z[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) z[j] = 0u;
#pragma unroll
      for (int k = 0; k < NPIECE; ++k) tmem_st_32x32b_x32(trow + DB_COL0 + 32u * k, z);
      tmem_st_wait();
    }
    // x = (s scale + bias) log2e of one 32-column piece
    auto scores_piece = [&](const uint32_t (&rs)[32], int k, float (&x)[32]) {
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        const uint32_t w = brow[16 * k + jj];
        x[2 * jj] = fmaf(__uint_as_float(rs[2 * jj]), sl2, bf16lo_to_f32(w));
        x[2 * jj + 1] = fmaf(__uint_as_float(rs[2 * jj + 1]), sl2, bf16hi_to_f32(w));
      }
    };
    TDECL(w_s = 0, w_dp = 0, w_pe = 0, t_begin = clock64());
    for (int it = 0; it < n_my; ++it) {
      const int t = (int)blockIdx.x + it * (int)gridDim.x;
      const uint32_t scol = S_N * ((uint32_t)it & 1u);
      const int64_t grow = ((int64_t)t * p.H + h) * (int64_t)L + r;
      TWAIT(w_s, mbar_wait(s_full((uint32_t)it & 1u), ((uint32_t)it >> 1) & 1u, 5));
      tcgen05_fence_after();
      // The four passes below run as ROLLED loops over the three 32-column pieces (a fully unrolled body is ~60 KB of
      // code: with one softmax warp per scheduler the instruction fetch, not the math, became the limit); the packed
      // results of a piece are moved into their slice of pk / dk by a k-dispatched copy so the arrays stay in registers.
      // ---- pass 1: row maximum ----
      float mx = -INFINITY;
#pragma unroll 1
      for (int k = 0; k < NPIECE; ++k) {
        uint32_t rs[32];
        tmem_ld_32x32b_x32(trow + scol + 32u * k, rs);
        tmem_ld_wait();
        float x[32];
        scores_piece(rs, k, x);
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, x[j]);
      }
      // ---- pass 2: e = exp2(x - max) (exact zeros in the rows of padding) back over S, row sum ----
      float sum = 0.f;
#pragma unroll 1
      for (int k = 0; k < NPIECE; ++k) {
        uint32_t rs[32];
        tmem_ld_32x32b_x32(trow + scol + 32u * k, rs);
        tmem_ld_wait();
        float x[32];
        scores_piece(rs, k, x);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float e = valid ? ex2_approx(x[j] - mx) : 0.f;
          sum += e;
          rs[j] = __float_as_uint(e);
        }
        tmem_st_32x32b_x32(trow + scol + 32u * k, rs);
      }
      tmem_st_wait();
      const float inv = valid ? 1.0f / sum : 0.f;
      const float ic = inv * p.drop_scale;
      if (BWD) {
        TWAIT(w_dp, mbar_wait(dp_full, (uint32_t)it & 1u, 8));
        tcgen05_fence_after();
      }
      // ---- pass 3: post-dropout probabilities (packed bf16) ; backward: t = P_dropped * dP back over dP, delta ----
      uint32_t pk[16 * NPIECE];
      float delta = 0.f;
#pragma unroll 1
      for (int k = 0; k < NPIECE; ++k) {
        uint32_t re[32], rd[32], tmp[16];
        tmem_ld_32x32b_x32(trow + scol + 32u * k, re);
        if (BWD) tmem_ld_32x32b_x32(trow + DP_COL0 + 32u * k, rd);
        tmem_ld_wait();
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
            pd[jj] = kept ? __uint_as_float(re[8 * u + jj]) * ic : 0.f;
            if (BWD) {
              const float x = valid ? pd[jj] * __uint_as_float(rd[8 * u + jj]) : 0.f;
              rd[8 * u + jj] = __float_as_uint(x);
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
          for (int jj = 0; jj < 4; ++jj) tmp[4 * u + jj] = pack_bf16x2(pd[2 * jj], pd[2 * jj + 1]);
        }
        if (BWD) tmem_st_32x32b_x32(trow + DP_COL0 + 32u * k, rd);
        if (k == 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = tmp[j];
        } else if (k == 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[16 + j] = tmp[j];
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[32 + j] = tmp[j];
        }
      }
      uint32_t dk[BWD ? 16 * NPIECE : 1];
      if (BWD) {
        tmem_st_wait();
        // ---- pass 4: scale * dS = scale (t - p delta) packed bf16 ; bias gradient (scaled; unscaled at the flush) ----
        const float nk = -inv * delta * p.scale;
#pragma unroll 1
        for (int k = 0; k < NPIECE; ++k) {
          uint32_t re[32], rt[32], rb[32], tmp[16];
          tmem_ld_32x32b_x32(trow + scol + 32u * k, re);
          tmem_ld_32x32b_x32(trow + DP_COL0 + 32u * k, rt);
          if (want_db) tmem_ld_32x32b_x32(trow + DB_COL0 + 32u * k, rb);
          tmem_ld_wait();
#pragma unroll
          for (int j2 = 0; j2 < 16; ++j2) {
            const float d0 = fmaf(__uint_as_float(re[2 * j2]), nk, __uint_as_float(rt[2 * j2]) * p.scale);
            const float d1 = fmaf(__uint_as_float(re[2 * j2 + 1]), nk, __uint_as_float(rt[2 * j2 + 1]) * p.scale);
            if (want_db) {
              rb[2 * j2] = __float_as_uint(__uint_as_float(rb[2 * j2]) + d0);
              rb[2 * j2 + 1] = __float_as_uint(__uint_as_float(rb[2 * j2 + 1]) + d1);
            }
            tmp[j2] = pack_bf16x2(d0, d1);
          }
          if (want_db) tmem_st_32x32b_x32(trow + DB_COL0 + 32u * k, rb);
          if (k == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) dk[BWD ? j : 0] = tmp[j];
          } else if (k == 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) dk[BWD ? 16 + j : 0] = tmp[j];
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) dk[BWD ? 32 + j : 0] = tmp[j];
          }
        }
        if (want_db) tmem_st_wait();
      }
      if (it > 0) TWAIT(w_pe, mbar_wait(p_empty, ((uint32_t)it - 1u) & 1u, 7));  // the products of the previous tile have read P (dS)
#pragma unroll
      for (int k = 0; k < NPIECE; ++k) {
        store_piece_packed(sP, r, 32 * k, pk + 16 * k);
        if (BWD) store_piece_packed(sDS, r, 32 * k, dk + (BWD ? 16 * k : 0));
      }
      fence_proxy_async();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    if (threadIdx.x == 0)
      TPRINT("softmax : total %lld cyc, waiting for S %lld, for dP %lld, for P / dS to be released %lld\n",
             clock64() - t_begin, w_s, w_dp, w_pe);
    if (want_db) {
      const float unscale = 1.0f / p.scale;
#pragma unroll
      for (int k = 0; k < NPIECE; ++k) {
        uint32_t rb[32];
        tmem_ld_32x32b_x32(trow + DB_COL0 + 32u * k, rb);
        tmem_ld_wait();
        if (r >= 1 && r < L) {
          float* drow = p.dbias + ((int64_t)h * L + r) * L;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = 32 * k + j;
            if (col >= 1 && col < L) atomicAdd(drow + col, __uint_as_float(rb[j]) * unscale);
          }
        }
      }
    }
  } else if (warp < 8 || warp >= 12) {
    // ===================== epilogue warps: two groups (warps 4-7 and 12-15), each takes every other product and owns
    // one staging tile, so the TMEM drain / convert / TMA store of consecutive products overlap =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 96;" ::: "memory");
    const int q = warp & 3;
    const int grp = warp >= 12 ? 1 : 0;
    const int r = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool leader = threadIdx.x == (grp ? 12 * 32 : 4 * 32);
    const uint32_t stg = sStg + (uint32_t)grp * TILE_B;
    uint32_t n = 0;
    TDECL(w_af = 0, t_begin = clock64());
    for (int it = 0; it < n_my; ++it) {
      const int w0 = (int)blockIdx.x + it * (int)gridDim.x;
#pragma unroll 1
      for (int prod = 0; prod < NPROD; ++prod, ++n) {
        if ((int)(n & 1u) != grp) continue;
        const uint32_t slot = n % NACC;
        TWAIT(w_af, mbar_wait(acc_full(slot), (n / NACC) & 1u, 6));
        tcgen05_fence_after();
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(trow + acc_col(slot), r0);
        tmem_ld_32x32b_x32(trow + acc_col(slot) + 32u, r1);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(slot));
        if (leader) tma_wait_group_read<0>();  // this group's previous store no longer reads its staging tile
        named_bar_sync(1 + grp, 128);
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
        named_bar_sync(1 + grp, 128);
        if (leader) {
          const int c = BWD ? prod / 3 : prod;
          const int which = BWD ? prod - 3 * c : 0;
          tma_store_3d_hint(&tm_out, stg, which * HD + h * DK + 64 * c, 0, w0, L2_EVICT_FIRST);
          tma_commit_group();
        }
      }
    }
    if (leader) tma_wait_group<0>();
    if (leader) TPRINT("epilogue%d: total %lld cyc, waiting for an accumulator %lld\n", grp, clock64() - t_begin, w_af);
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 9) tmem_dealloc(tmem_base, 512);
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
  constexpr int NACC = BWD ? 2 : 4;  // even: slot n % NACC of product n is then private to issuer n % 2
  constexpr int MAXP = LP / 32;
  constexpr int NPROD = BWD ? 3 * NC : NC;
  constexpr uint32_t SUB_B = 8192;  // 64 rows x 128 B
  constexpr float LOG2E = 1.4426950408889634f;
  constexpr uint32_t DP_COL0 = 128u, DB_COL0 = 384u;
  auto acc_col = [](uint32_t a) -> uint32_t { return (BWD ? 256u : 128u) + 64u * a; };

  Params p = tp.p;
  p.offset += rng_step();
  const int NS = tp.n_stages, NSA = tp.n_stages_a, NSB = tp.n_stages_b;
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
  // two operand rings with their own producer and issuer thread each: A = operands of the score products, B = operands of
  // the chunk products; ring B's stages follow ring A's in shared memory
  const uint32_t sRingB = sRing + (uint32_t)NSA * TILE_B;
  auto full_a = [&](int s) { return sBar + 8u * s; };
  auto empty_a = [&](int s) { return sBar + 8u * (MAX_STAGES + s); };
  auto full_b = [&](int s) { return sBar + 8u * (2 * MAX_STAGES + s); };
  auto empty_b = [&](int s) { return sBar + 8u * (3 * MAX_STAGES + s); };
  auto s_full = [&](uint32_t b) { return sBar + 8u * (4 * MAX_STAGES + b); };
  const uint32_t p_full = sBar + 8u * (4 * MAX_STAGES + 2), p_empty = p_full + 8u;
  auto acc_full = [&](int a) { return sBar + 8u * (4 * MAX_STAGES + 4 + a); };
  auto acc_empty = [&](int a) { return sBar + 8u * (4 * MAX_STAGES + 8 + a); };
  const uint32_t tmem_slot = sBar + 8u * (4 * MAX_STAGES + 12);
  float* bias_s = reinterpret_cast<float*>(smem + (sBias - smem_u32(smem)));

  // ---------------- start-up ----------------
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_out);
    if (BWD) tma_prefetch_desc(&tm_do);
    for (int s = 0; s < NSA; ++s) {
      mbar_init(full_a(s), 1);
      mbar_init(empty_a(s), 1);
    }
    for (int s = 0; s < NSB; ++s) {
      mbar_init(full_b(s), 1);
      mbar_init(empty_b(s), 1);
    }
    mbar_init(s_full(0), 1);
    mbar_init(s_full(1), 1);
    mbar_init(p_full, 4);
    mbar_init(p_empty, 2);  // one commit per chunk-product issuer
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
    // ===================== TMA producer: ONE thread feeds both rings, each in its consumers' order; it polls the two
    // "stage free" barriers so that a full ring never holds up the other one =====================
    if (lane == 0) {
      const uint64_t pol_again = BWD ? L2_EVICT_LAST : L2_EVICT_FIRST, pol_once = L2_EVICT_FIRST;
      // backward: q, k, dO are read again by the chunk products of the same tile -> keep them in L2 until then;
      // everything that is read for the last time leaves L2 first
      const int cq = h * DK, ck = HD + h * DK, cv = 2 * HD + h * DK;
      const int na = n_my * (BWD ? 4 * NC : 2 * NC), nb = n_my * (BWD ? 3 * NC : NC);
      int ia = 0, ib = 0, sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      TDECL(t_begin = clock64());
      while (ia < na || ib < nb) {
        if (ia < na && mbar_try_wait(empty_a(sa), pha ^ 1u)) {
          const int per = BWD ? 4 * NC : 2 * NC;          // stages of ring A per tile: per chunk q, k [, dO, v]
          const int it = ia / per, j = ia - it * per;
          const int pc = BWD ? 4 : 2;
          const int c = j / pc, which = j - c * pc;      // 0 q, 1 k, 2 dO, 3 v
          const CUtensorMap* tm = which == 2 ? &tm_do : &tm_qkv;
          const int col = which == 0 ? cq : which == 1 ? ck : which == 2 ? cq : cv;
          const uint64_t pol = which == 3 ? pol_once : pol_again;
          mbar_arrive_expect_tx(full_a(sa), TILE_B);
          tma_load_3d_hint(sRing + (uint32_t)sa * TILE_B, tm, full_a(sa), col + 64 * c, 0, (((int)blockIdx.x + it * (int)gridDim.x) * G), pol);
          ++ia;
          if (++sa == NSA) { sa = 0; pha ^= 1u; }
        }
        if (ib < nb && mbar_try_wait(empty_b(sb), phb ^ 1u)) {
          const int per = BWD ? 3 * NC : NC;              // stages of ring B per tile: per chunk k, q, dO (or v)
          const int it = ib / per, j = ib - it * per;
          const int c = BWD ? j / 3 : j, which = BWD ? j - 3 * c : 0;
          const CUtensorMap* tm = (BWD && which == 2) ? &tm_do : &tm_qkv;
          const int col = BWD ? (which == 0 ? ck : cq) : cv;
          mbar_arrive_expect_tx(full_b(sb), TILE_B);
          tma_load_3d_hint(sRingB + (uint32_t)sb * TILE_B, tm, full_b(sb), col + 64 * c, 0, (((int)blockIdx.x + it * (int)gridDim.x) * G), pol_once);
          ++ib;
          if (++sb == NSB) { sb = 0; phb ^= 1u; }
        }
      }
      TPRINT("producer: total %lld cyc\n", clock64() - t_begin);
    }
  } else if (warp >= 9 && warp <= 11) {
    // ===================== MMA issuers (see attn_tc128_kernel): warp 9 score products from ring A, warp 10 chunk
    // products from ring B =====================
    if (lane == 0) {
      const bool do_a = warp == 9;
      const int n_s = do_a ? NSA : NSB;
      const uint32_t base = do_a ? sRing : sRingB;
      const uint32_t M64 = (64u >> 4) << 24, M128 = (128u >> 4) << 24;
      const uint32_t id_s = (idesc_bf16_m128(64u, false, false) & ~M128) | M64;
      const uint32_t id_pk = (idesc_bf16_m128(64u, false, true) & ~M128) | M64;
      const uint32_t id_ptk = (idesc_bf16_m128(64u, true, true) & ~M128) | M64;
      int s = 0;
      uint32_t ph = 0;
      TDECL(w_full = 0, w_acc = 0, w_p = 0, t_begin = clock64());
      auto acquire = [&]() -> uint32_t {
        TWAIT(w_full, mbar_wait(do_a ? full_a(s) : full_b(s), ph, 2));
        return base + (uint32_t)s * TILE_B;
      };
      auto advance = [&]() { if (++s == n_s) { s = 0; ph ^= 1u; } };
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
        umma_commit(empty_a(sa));
        umma_commit(empty_a(s));
        advance();
      };
      // per sub-tile g: slot_g (64 x 64) = A_g (P or dS, [64 x 64], K- or MN-major) * ring tile rows of g (MN-major B)
      auto chunk_product = [&](uint32_t n, uint32_t me, uint32_t a_base, bool a_mn) {
        if ((n & 1u) != me) return;
        const uint32_t sb = n % (uint32_t)NSB, slot = n % NACC;
        TWAIT(w_full, mbar_wait(full_b(sb), (n / (uint32_t)NSB) & 1u, 2));
        const uint32_t b = sRingB + sb * TILE_B;
        TWAIT(w_acc, mbar_wait(acc_empty(slot), ((n / NACC) & 1u) ^ 1u, 3));
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
        umma_commit(empty_b(sb));
        umma_commit(acc_full(slot));
      };
      if (do_a) {
        // score products of tile `it` into buffer it & 1, published to the softmax warps by the last chunk
        auto phase_a = [&](int it) {
          const uint32_t b = (uint32_t)it & 1u;
          for (int c = 0; c < NC; ++c) {
            score_product(64u * b, c == 0);
            if (BWD) score_product(DP_COL0 + 64u * b, c == 0);
          }
          umma_commit(s_full(b));
        };
        for (int it = 0; it < 2 && it < n_my; ++it) phase_a(it);
        for (int it = 0; it + 2 < n_my; ++it) {
          TWAIT(w_p, mbar_wait(p_full, (uint32_t)it & 1u, 4));  // the softmax has finished reading S / dP buffer it & 1
          tcgen05_fence_after();
          phase_a(it + 2);
        }
      } else {
        // chunk products n = 0, 1, 2, ... of this CTA in order; issuer j takes those with n % 2 == j.  Product n reads
        // stage n % NSB of ring B (one stage per product) and writes accumulator slot n % NACC.
        const uint32_t me = (uint32_t)(warp - 10);
        uint32_t n = 0;
        for (int it = 0; it < n_my; ++it) {
          TWAIT(w_p, mbar_wait(p_full, (uint32_t)it & 1u, 4));  // P / dS of tile `it` written
          tcgen05_fence_after();
          for (int c = 0; c < NC; ++c) {
            if (BWD) {
              chunk_product(n++, me, sDS, false);  // dQ_c = dS K_c
              chunk_product(n++, me, sDS, true);   // dK_c = dS^T Q_c
              chunk_product(n++, me, sP, true);    // dV_c = P^T dO_c
            } else {
              chunk_product(n++, me, sP, false);   // O_c = P V_c
            }
          }
          umma_commit(p_empty);  // (one arrival per issuer) P / dS may be overwritten once these products retire
        }
      }
      TPRINT("mma%d    : total %lld cyc, waiting for operands %lld, for an accumulator slot %lld, for the softmax %lld (%d tiles)\n",
             warp - 9, clock64() - t_begin, w_full, w_acc, w_p, n_my);
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
    TDECL(w_s = 0, w_pe = 0, t_begin = clock64());
    for (int it = 0; it < n_my; ++it) {
      const int t = (int)blockIdx.x + it * (int)gridDim.x;
      const uint32_t b = (uint32_t)it & 1u;
      const int64_t w = (int64_t)t * G + wt;
      const bool valid = (i < L) && (w < p.W);
      const int64_t grow = (w * p.H + h) * (int64_t)L + i;
      TWAIT(w_s, mbar_wait(s_full(b), ((uint32_t)it >> 1) & 1u, 5));
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
      if (it > 0) TWAIT(w_pe, mbar_wait(p_empty, ((uint32_t)it - 1u) & 1u, 7));  // the products of the previous tile have read P (dS)
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
    if (threadIdx.x == 0)
      TPRINT("softmax : total %lld cyc, waiting for S / dP %lld, for P / dS to be released %lld\n", clock64() - t_begin, w_s, w_pe);
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
    TDECL(w_af = 0, t_begin = clock64());
    for (int it = 0; it < n_my; ++it) {
      const int w0 = ((int)blockIdx.x + it * (int)gridDim.x) * G;
#pragma unroll 1
      for (int prod = 0; prod < NPROD; ++prod, ++n) {
        const uint32_t slot = n % NACC;
        TWAIT(w_af, mbar_wait(acc_full(slot), (n / NACC) & 1u, 6));
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
    if (leader) TPRINT("epilogue: total %lld cyc, waiting for an accumulator %lld\n", clock64() - t_begin, w_af);
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
  if constexpr (LP == 128) return attn_tc128_kernel<DK, BWD>;
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
  tp.npiece = (G > 1) ? LP / 32 : 3;
  uint32_t stage_b, fixed;
  if (G > 1) {
    // sub-tile kernel: bias always staged (fp32, LP columns per row, padding columns hold -inf); P / dS 2 x [64 x 64]
    tp.bias_pitch = LP + 1;
    tp.bias_bytes = (uint32_t)(((size_t)p.L * tp.bias_pitch * 4 + 15) / 16 * 16);
    stage_b = TILE_B;
    fixed = 1024u + (BWD ? 4u : 2u) * 8192u + 2u * TILE_B + tp.bias_bytes + BAR_BYTES;
  } else {
    // one window per tile: stages of nkeys rows, bias as bf16 pairs (49 words per row)
    tp.bias_pitch = 49;
    tp.bias_bytes = (uint32_t)(((size_t)p.L * 49 * 4 + 15) / 16 * 16);
    stage_b = (uint32_t)tp.nkeys * 128u;
    fixed = 1024u + (BWD ? 4u : 2u) * PANEL_B + 2u * TILE_B + tp.bias_bytes + BAR_BYTES;
  }
  int ns = (int)((SMEM_LIMIT - fixed) / stage_b);
  if (ns > MAX_STAGES) ns = MAX_STAGES;
  {
    static const int cap = [] {  // LSTC_ATTN_STAGES=<n>: cap the ring depth (experiments on latency hiding)
      const char* e = getenv("LSTC_ATTN_STAGES");
      return e != nullptr ? atoi(e) : 0;
    }();
    if (cap >= 3 && ns > cap) ns = cap;
  }
  if (ns < 4) {
    set_last_error("attention: no shared memory left for the operand ring (L=%d)", p.L);
    return LSTC_ERR_UNSUPPORTED;
  }
  // ring B (chunk-product operands): forward = v (first read, HBM-bound, a third of the loads); backward = k, q, dO
  // re-read through L2 (short latency).  Product n uses stage n % nsb and is issued by thread n % 2: nsb must be EVEN so
  // that every stage (like every accumulator slot) is private to one issuing thread - a parity wait cannot tell two
  // laps of a barrier apart, so consecutive uses of a stage have to be ordered by one thread's program order.
  int nsb = BWD ? 4 : (((ns + 2) / 3 + 1) & ~1);
  {
    static const int nsb_env = [] {  // LSTC_ATTN_STAGES_B=<n>: depth of ring B (experiments)
      const char* e = getenv("LSTC_ATTN_STAGES_B");
      return e != nullptr ? atoi(e) : 0;
    }();
    if (nsb_env > 0) nsb = nsb_env & ~1;
  }
  if (nsb < 2) nsb = 2;
  if (ns - nsb < 3) nsb = 2;
  if (ns - nsb < 2) {
    set_last_error("attention: no shared memory left for the operand rings (L=%d)", p.L);
    return LSTC_ERR_UNSUPPORTED;
  }
  tp.n_stages = ns;
  tp.n_stages_a = ns - nsb;
  tp.n_stages_b = nsb;
  const uint32_t smem = fixed + (uint32_t)ns * stage_b;
  const int box_rows = (G > 1) ? LP : tp.nkeys;
  const int HD = p.H * DK;
  CUtensorMap tq, td, to;
  memset(&td, 0, sizeof(td));
  int rc = make_tmap3d(&tq, p.qkv, 3 * (int64_t)HD, p.L, p.W, p.ld, box_rows, G);
  if (rc != LSTC_OK) return rc;
  if (BWD) {
    rc = make_tmap3d(&td, p.dout, HD, p.L, p.W, p.ld_dout, box_rows, G);
    if (rc != LSTC_OK) return rc;
  }
  rc = make_tmap3d(&to, p.out, (BWD ? 3 : 1) * (int64_t)HD, p.L, p.W, p.ld_out, box_rows, G);
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
  kern<<<dim3((unsigned)gx, (unsigned)p.H), LP == 128 ? NUM_THREADS128 : NUM_THREADS64, smem, stream>>>(tp, tq, td, to);
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
