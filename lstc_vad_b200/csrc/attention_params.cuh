// Argument block shared by the two attention implementations (attention_tc.cu: tcgen05 / TMEM; attention.cu: mma.sync).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lstc {
namespace attn {

struct Params {
  const __nv_bfloat16* qkv;  // [W*L, ld] : q | k | v column blocks of H*dk, head h in columns h*dk..
  int64_t ld;
  const __nv_bfloat16* dout;  // bwd: [W*L, ld_dout]
  int64_t ld_dout;
  int64_t W;
  int L, H;
  const float* bias;  // dense [H,L,L] (zeros in the CLS row / column) or null
  float scale;
  float drop_p, drop_scale;
  uint32_t drop_thr16;
  uint64_t seed, offset;
  __nv_bfloat16* out;  // fwd: O [W*L, ld_out] ; bwd: dqkv [W*L, ld_out]
  int64_t ld_out;
  float* probs;  // fwd optional [W,H,L,L]
  float* dbias;  // bwd optional [H,L,L] (zeroed by the caller; accumulated atomically)
};

}  // namespace attn

namespace attn_tc {
// tcgen05 implementation; returns LSTC_ERR_UNSUPPORTED (with the reason in lstc_last_error) for shapes it does not cover
int run(bool bwd, const attn::Params& p, int dk, cudaStream_t stream);
}  // namespace attn_tc
}  // namespace lstc
